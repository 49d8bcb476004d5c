"""The shuffle engine: map-side hash-partition, exchange, reduce-side merge over
columnar partitions held as torch CUDA tensors.

Mirrors, for the hot path only:
  ShuffleMapTask._run          dpark/task.py:197-255      -> map_side()
  ShuffleFetcher.fetch         dpark/shuffle.py:309-420   -> exchange()  (one NCCL alltoallv)
  DiskHashMerger._merge        dpark/shuffle.py:600-608   -> reduce_side()
  MapOutputTracker             dpark/shuffle.py:809-826   -> the counts matrix

Partition ownership across G ranks: reduce partition r lives on rank
r // ceil(P/G) (contiguous blocks), so the rows a rank sends to one peer are one
contiguous range of its bucket-major buffer.  Map splits are assigned to ranks in
contiguous blocks too, so "source rank order" == "map_id order".
"""
import torch

from . import _native as nv


def owner_blocks(P, G):
    """[first partition of rank g for g in 0..G] (len G+1), contiguous blocks."""
    per = (P + G - 1) // G
    return [min(P, g * per) for g in range(G + 1)]


class MapOutput(object):
    """Bucket-major output of the map side on one rank: the alltoallv send buffer."""
    __slots__ = ("keys", "vals", "offsets", "P")

    def __init__(self, keys, vals, offsets, P):
        self.keys, self.vals, self.offsets, self.P = keys, vals, offsets, P


def map_side(key_chunks, val_chunks, P, thresholds=None, prehashed=False):
    """Hash-partition all local map splits into ONE bucket-major buffer.

    key_chunks/val_chunks: lists of CUDA tensors (the rank's map splits in map_id
    order).  Rows of bucket p are ordered by (map split, position) -- the order
    OrderedGroupByDiskHashMerger produces (dpark/shuffle.py:626-646)."""
    if len(key_chunks) == 1:
        k, v, off = nv.partition(key_chunks[0], val_chunks[0], P, thresholds, prehashed)
        return MapOutput(k, v, off, P)
    dev = key_chunks[0].device
    counts, wss = [], []
    for k in key_chunks:
        c, ws = nv.partition_count(k, P, thresholds, prehashed)
        counts.append(c)
        wss.append(ws)
    cm = torch.stack(counts)                       # [M, P]
    tot = cm.sum(0)                                # rows per bucket
    offsets = torch.zeros(P + 1, dtype=torch.int64, device=dev)
    torch.cumsum(tot, 0, out=offsets[1:])
    # base[m][p] = offsets[p] + rows of bucket p in earlier splits
    base = offsets[:-1].unsqueeze(0) + (torch.cumsum(cm, 0) - cm)
    n = sum(int(k.numel()) for k in key_chunks)
    out_k = torch.empty(n, dtype=key_chunks[0].dtype, device=dev)
    has_v = val_chunks[0] is not None
    out_v = torch.empty(n, dtype=val_chunks[0].dtype, device=dev) if has_v else None
    for m, (k, v) in enumerate(zip(key_chunks, val_chunks)):
        nv.partition_scatter(k, v, P, base[m].contiguous(), out_k, out_v, wss[m], thresholds, prehashed)
    return MapOutput(out_k, out_v, offsets, P)


class Received(object):
    """Rows fetched for the partitions this rank owns.  keys/vals are laid out
    source-rank-major, then bucket-major; seg[s][j] = rows from source s for
    local partition j."""
    __slots__ = ("keys", "vals", "seg", "part_first", "nparts")

    def __init__(self, keys, vals, seg, part_first, nparts):
        self.keys, self.vals, self.seg, self.part_first, self.nparts = keys, vals, seg, part_first, nparts


def exchange(mo, group=None):
    """ShuffleFetcher replacement: one alltoallv of the bucket-major buffers
    (torch.distributed all_to_all_single with split sizes == grouped
    ncclSend/ncclRecv over NVLink).  With one rank it is the identity."""
    import torch.distributed as dist
    G = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    P = mo.P
    if G == 1:
        seg = (mo.offsets[1:] - mo.offsets[:-1]).unsqueeze(0)
        return Received(mo.keys, mo.vals, seg, 0, P)
    rank = dist.get_rank(group)
    blocks = owner_blocks(P, G)
    counts = mo.offsets[1:] - mo.offsets[:-1]                       # [P] rows I hold per bucket
    all_counts = torch.empty(G * P, dtype=torch.int64, device=counts.device)
    dist.all_gather_into_tensor(all_counts, counts.contiguous(), group=group)   # MapOutputTracker
    all_counts = all_counts.view(G, P)
    host_counts = all_counts.cpu()                                  # sizes must be known on the host
    send_splits = [int(host_counts[rank, blocks[d]:blocks[d + 1]].sum()) for d in range(G)]
    p0, p1 = blocks[rank], blocks[rank + 1]
    recv_splits = [int(host_counts[s, p0:p1].sum()) for s in range(G)]
    nrecv = sum(recv_splits)
    rk = torch.empty(nrecv, dtype=mo.keys.dtype, device=mo.keys.device)
    dist.all_to_all_single(rk, mo.keys, recv_splits, send_splits, group=group)
    rv = None
    if mo.vals is not None:
        rv = torch.empty(nrecv, dtype=mo.vals.dtype, device=mo.vals.device)
        dist.all_to_all_single(rv, mo.vals, recv_splits, send_splits, group=group)
    seg = all_counts[:, p0:p1].contiguous()
    return Received(rk, rv, seg, p0, p1 - p0)


def reduce_side(rx, op, P, thresholds=None):
    """DiskHashMerger._merge over everything received.  Returns
    (keys, vals, part_offsets[nparts+1], counts[nparts]): distinct keys of local
    partition j are keys[part_offsets[j] : part_offsets[j] + counts[j]]."""
    dev = rx.keys.device
    rows = rx.seg.sum(0)                                            # rows per local partition
    part_offsets = torch.zeros(rx.nparts + 1, dtype=torch.int64, device=dev)
    if rx.nparts:
        torch.cumsum(rows, 0, out=part_offsets[1:])
    if rx.nparts == 0:
        return rx.keys[:0], rx.vals[:0], part_offsets, rows
    ok, ov, cnt = nv.combine(rx.keys, rx.vals, op, P, part_offsets, rx.part_first, rx.nparts, thresholds)
    return ok, ov, part_offsets, cnt


class HostShuffle(object):
    """End-to-end reduceByKey for HOST-resident columns: the call a user of the
    plugin makes when rows arrive from Python / files (SURVEY.md §8b seam 2).
    Per call: pinned host -> device copy of every map split, map_side,
    exchange, reduce_side, device -> pinned host copy of every partition's
    distinct (key, combined) rows.  Buffers are allocated once and reused."""

    def __init__(self, n_rows, key_dtype, val_dtype, P, op="sum", splits=8, thresholds=None, group=None,
                 device=None):
        self.P, self.op, self.splits, self.thresholds, self.group = P, op, splits, thresholds, group
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.n = n_rows
        self.h_keys = torch.empty(n_rows, dtype=key_dtype).pin_memory()
        self.h_vals = torch.empty(n_rows, dtype=val_dtype).pin_memory()
        self.d_keys = torch.empty(n_rows, dtype=key_dtype, device=self.device)
        self.d_vals = torch.empty(n_rows, dtype=val_dtype, device=self.device)
        self.out_keys = None
        self.out_vals = None
        self.h2d_bytes = n_rows * (self.h_keys.element_size() + self.h_vals.element_size())
        self.d2h_bytes = 0

    def _bounds(self):
        per = (self.n + self.splits - 1) // self.splits
        return [(min(self.n, i * per), min(self.n, (i + 1) * per)) for i in range(self.splits)]

    def run(self):
        kc, vc = [], []
        for a, b in self._bounds():
            self.d_keys[a:b].copy_(self.h_keys[a:b], non_blocking=True)
            self.d_vals[a:b].copy_(self.h_vals[a:b], non_blocking=True)
            kc.append(self.d_keys[a:b])
            vc.append(self.d_vals[a:b])
        mo = map_side(kc, vc, self.P, self.thresholds)
        rx = exchange(mo, self.group)
        ok, ov, po, cnt = reduce_side(rx, self.op, self.P, self.thresholds)
        po_h, cnt_h = po.cpu().tolist(), cnt.cpu().tolist()      # the one host sync: result sizes
        nrx = int(ok.numel())
        if self.out_keys is None or self.out_keys.numel() < nrx:
            self.out_keys = torch.empty(nrx, dtype=ok.dtype).pin_memory()
            self.out_vals = torch.empty(nrx, dtype=ov.dtype).pin_memory()
        res, d2h = [], 0
        for j in range(rx.nparts):
            a, c = po_h[j], cnt_h[j]
            self.out_keys[a:a + c].copy_(ok[a:a + c], non_blocking=True)
            self.out_vals[a:a + c].copy_(ov[a:a + c], non_blocking=True)
            d2h += c * (ok.element_size() + ov.element_size())
            res.append((rx.part_first + j, self.out_keys[a:a + c], self.out_vals[a:a + c]))
        torch.cuda.current_stream().synchronize()
        self.d2h_bytes = d2h
        return res


def reduce_by_key(key_chunks, val_chunks, P, op="sum", thresholds=None, group=None):
    """Whole hot path for this rank's map splits.  Returns a list of
    (partition id, keys, vals) for the partitions this rank owns (device tensors)."""
    mo = map_side(key_chunks, val_chunks, P, thresholds)
    rx = exchange(mo, group)
    ok, ov, po, cnt = reduce_side(rx, op, P, thresholds)
    po_h, cnt_h = po.cpu().tolist(), cnt.cpu().tolist()
    return [(rx.part_first + j, ok[po_h[j]:po_h[j] + cnt_h[j]], ov[po_h[j]:po_h[j] + cnt_h[j]])
            for j in range(rx.nparts)]
