"""Runs one ShuffledRDD on the GPU: the static two-stage plan that replaces the
reference's DAG scheduler for this path (out of scope as a
system; the map stage and the reduce stage are simply run in order).

    stage 1 (map)    every parent split -> columns -> HBM -> dpark_b200.shuffle.map_side
    stage 2 (reduce) exchange + reduce_side / group_side -> per-partition columns

Python rows exist only before stage 1 (ingest of what user lambdas produced) and
after stage 2 (egress to user lambdas); see dpark_b200.columnar.
"""
import numpy as np
import torch

from . import _native as nv
from . import columnar, shuffle


class ShuffleResult(object):
    """Per-partition result columns on the host + lazy conversion to rows."""

    def __init__(self, nparts):
        self.parts = [None] * nparts      # (keys: list, vals: list | (offsets, values)) per partition

    def rows(self, p):
        keys, vals = self.parts[p]
        return list(zip(keys, vals))

    def columns(self, p):
        return self.parts[p]


def _device():
    if not torch.cuda.is_available():
        raise nv.NativeError("the dpark_b200 shuffle needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _gather_parent(srdd, numeric_values, only=None):
    """Stage-1 input: one Columns (host) or a (keys, vals) tensor pair per parent split (`only`: the split indices
    this rank owns under torch.distributed)."""
    from .rdd import ColumnarRDD
    parent = srdd.parent
    out = []
    for i, sp in enumerate(parent.splits):
        if only is not None and i not in only:
            continue
        if isinstance(parent, ColumnarRDD) and numeric_values:
            out.append(parent.columns(sp))
        else:
            out.append(columnar.ingest_pairs(parent.iterator(sp), repr(parent), numeric_values))
    return out


TEXT_INGEST = True            # dpark_b200.textingest: tokenise textFile -> split -> (w, 1) pipelines on the device


def run_shuffle(srdd):
    from . import spmd
    rank, world = spmd.rank_world()
    if world > 1:
        return _run_shuffle_spmd(srdd, rank, world)
    dev = _device()
    P = srdd.partitioner.numPartitions
    thr = srdd.partitioner.thresholds
    if srdd.kind == "reduce":
        from . import textingest
        text = textingest.recognize(srdd.parent) if TEXT_INGEST else None
        if text is not None:     # the word-count shape: tokenise on the device (None back: a non-ASCII split, row-wise path)
            res = textingest.reduce_tokens(text, range(len(text.splits)), P, thr, srdd.op, dev, ShuffleResult(P))
            if res is not None:
                return res
        splits = _gather_parent(srdd, True)
        return _run_reduce(splits, P, thr, srdd.op, dev)
    splits = _gather_parent(srdd, False)
    return _run_group(splits, P, thr, dev)


# ---------------------------------------------------------------------------------------------------------------------
# one driver process per GPU (dpark_b200/spmd.py)
# ---------------------------------------------------------------------------------------------------------------------
ROUTE_EVERYTHING = False      # test hook: numeric reduceByKey also takes the routed path (CPU tests have no NCCL)


def _run_shuffle_spmd(srdd, rank, world):
    """The shuffle with the map stage spread over the ranks: every rank ingests the parent splits it owns.

      numeric keys + numeric values + reduce   device columns through shuffle.reduce_by_key: map_side -> NCCL
                                               alltoallv -> reduce_side, partitions owned in contiguous blocks;
      everything else (group-by: values are    the rows are routed on the HOST to the rank owning their partition
      Python objects; str / bytes keys)        (partition ids come from the CUDA hash / getPartition kernels), as
                                               pickled columns in one all_to_all; the owner then runs the one-GPU path
                                               over what it received, sources in rank order = map split order.

    Every rank ends up with the (small, host) result of every partition, so downstream narrow stages can run anywhere."""
    from . import spmd
    dev = _device()
    P = srdd.partitioner.numPartitions
    thr = srdd.partitioner.thresholds
    nsplits = len(srdd.parent.splits)
    mine = set(spmd.my_indices(nsplits, rank, world))
    numeric = srdd.kind == "reduce"
    splits = None
    if numeric and TEXT_INGEST:
        # the word-count shape: this rank's splits are tokenised AND combined on its GPU (the reference's map-side
        # combine, dpark/task.py:221-226); what is routed to the owners are the distinct (word, partial count) pairs
        from . import textingest
        text = textingest.recognize(srdd.parent)
        if text is not None:
            part = textingest.reduce_tokens(text, sorted(mine), P, thr, srdd.op, dev, ShuffleResult(P))
            if part is not None:
                ks = [k for p in range(P) for k in part.parts[p][0]]
                vs = [v for p in range(P) for v in part.parts[p][1]]
                splits = [columnar.ingest_pairs(zip(ks, vs), "textFile", True)]
    if splits is None:
        splits = _gather_parent(srdd, numeric, only=mine)
    blocks = shuffle.owner_blocks(P, world)
    tensor_in = bool(splits) and not isinstance(splits[0], columnar.Columns)
    # what every rank must agree on before any collective: key kind, value kinds
    if tensor_in:
        desc = ("tensor", str(splits[0][0].dtype), str(splits[0][1].dtype))
    else:
        desc = ("cols", sorted(set(c.key_kind for c in splits if c.n)), sorted(set(c.val_kind for c in splits if c.n)))
    descs = spmd.all_gather_objects(desc)
    res = ShuffleResult(P)
    if any(d[0] == "tensor" for d in descs):
        if not all(d[0] == "tensor" or d[1] == [] for d in descs):
            raise TypeError("mixed columnar and row inputs in one shuffle are not supported on the B200 path")
        kd = next(d for d in descs if d[0] == "tensor")
        kdt, vdt = getattr(torch, kd[1].split(".")[1]), getattr(torch, kd[2].split(".")[1])
        kc = [k.to(dev).contiguous() for k, v in splits] or [torch.empty(0, dtype=kdt, device=dev)]
        vc = [v.to(dev).contiguous() for k, v in splits] or [torch.empty(0, dtype=vdt, device=dev)]
        owned = _device_reduce(kc, vc, P, thr, srdd.op, world)
    else:
        kinds = sorted(set(k for d in descs for k in d[1]))
        vkinds = sorted(set(k for d in descs for k in d[2]))
        if len(kinds) > 1:
            raise TypeError("mixed key types %s in one shuffle are not supported on the B200 path" % kinds)
        kk = kinds[0] if kinds else columnar.KEY_I64
        if numeric and len(vkinds) > 1:
            raise TypeError("reduceByKey values must be all int or all float on the B200 path")
        if numeric and kk in (columnar.KEY_I64, columnar.KEY_F64) and not ROUTE_EVERYTHING:
            kdt = np.int64 if kk == columnar.KEY_I64 else np.float64
            vdt = torch.float64 if vkinds == [columnar.VAL_F64] else torch.int64
            kc = [torch.from_numpy(c.keys.astype(kdt, copy=False)).to(dev) for c in splits] or \
                [torch.empty(0, dtype=torch.int64 if kk == columnar.KEY_I64 else torch.float64, device=dev)]
            vc = [torch.from_numpy(c.vals).to(dev).to(vdt) for c in splits] or [torch.empty(0, dtype=vdt, device=dev)]
            owned = _device_reduce(kc, vc, P, thr, srdd.op, world)
        else:
            owned = _routed_shuffle(splits, kk, numeric, P, thr, srdd.op, dev, rank, world, blocks)
    for part in spmd.all_gather_objects(owned):
        for p, cols in part.items():
            res.parts[p] = cols
    for p in range(P):
        if res.parts[p] is None:
            res.parts[p] = ([], [])
    return res


def _device_reduce(kc, vc, P, thr, op, world):
    """Numeric reduceByKey across the ranks, columns stay on the devices: {partition: (keys, values)} for the
    partitions this rank owns."""
    from . import spmd
    rows = spmd.agree_max(sum(int(k.numel()) for k in kc))
    sb = shuffle.choose_sub_bits(max(rows, 1), P, world)      # every rank must use the same bucket layout
    parts = shuffle.reduce_by_key(kc, vc, P, op, thr, sub_bits=sb)
    return {p: (k.cpu().numpy().tolist(), v.cpu().numpy().tolist()) for p, k, v in parts}


def _routed_shuffle(splits, kk, numeric, P, thr, op, dev, rank, world, blocks):
    """Rows with Python-object values (group-by) or str / bytes keys: every row goes to the rank that owns its
    partition -- the partition id is computed by the CUDA kernels (portable_hash + getPartition), the rows travel as
    pickled host columns in one all_to_all -- and the owner runs the single-GPU shuffle over the received rows, which
    arrive in source-rank order, i.e. in map split order (splits are owned in contiguous blocks)."""
    from . import spmd
    per_dest = [([], []) for _ in range(world)]
    dest_of_part = np.zeros(P, dtype=np.int64)
    for d in range(world):
        dest_of_part[blocks[d]:blocks[d + 1]] = d
    for c in splits:
        if not c.n:
            continue
        keys = columnar.decode_keys(c.key_kind, c.keys, c.key_offsets, c.key_objs)
        vals = c.objs if c.objs is not None else c.vals.tolist()
        h = columnar._hash_column(keys)
        t = None if thr is None else torch.tensor(thr, dtype=torch.int64, device=h.device)
        pid = nv.partition_ids(h, P, t).cpu().numpy()
        dest = dest_of_part[pid]
        for d in np.unique(dest).tolist():
            idx = np.nonzero(dest == d)[0].tolist()
            per_dest[d][0].extend(keys[i] for i in idx)
            per_dest[d][1].extend(vals[i] for i in idx)
    got = spmd.all_to_all_objects(per_dest)
    local = [columnar.ingest_pairs(zip(ks, vs), "shuffle", numeric) for ks, vs in got]
    with shuffle.local_only():
        if numeric:
            r = _run_reduce(local, P, thr, op, dev)
        else:
            r = _run_group(local, P, thr, dev)
    return {p: r.parts[p] for p in range(blocks[rank], blocks[rank + 1])}


def _key_kind_of(splits):
    kinds = set(c.key_kind for c in splits if isinstance(c, columnar.Columns) and c.n)
    if len(kinds) > 1:
        raise TypeError("mixed key types %s in one shuffle are not supported on the B200 path" % sorted(kinds))
    return kinds.pop() if kinds else columnar.KEY_I64


def _run_reduce(splits, P, thr, op, dev):
    res = ShuffleResult(P)
    tensor_in = splits and not isinstance(splits[0], columnar.Columns)
    if tensor_in:
        kc = [k.to(dev).contiguous() for k, v in splits]
        vc = [v.to(dev).contiguous() for k, v in splits]
        parts = shuffle.reduce_by_key(kc, vc, P, op, thr)
        for p, k, v in parts:
            res.parts[p] = (k.cpu().numpy().tolist(), v.cpu().numpy().tolist())
        return res
    kk = _key_kind_of(splits)
    vkinds = set(c.val_kind for c in splits if c.n)
    if len(vkinds) > 1:
        raise TypeError("reduceByKey values must be all int or all float on the B200 path")
    _check_int_sum_range(splits, vkinds, op)
    if kk in (columnar.KEY_I64, columnar.KEY_F64):
        kdt = np.int64 if kk == columnar.KEY_I64 else np.float64
        kc = [torch.from_numpy(c.keys.astype(kdt, copy=False)).to(dev) for c in splits]
        vc = [torch.from_numpy(c.vals).to(dev) for c in splits]
        if vkinds:
            vdt = torch.int64 if vkinds == {columnar.VAL_I64} else torch.float64
            vc = [v.to(vdt) for v in vc]
        parts = shuffle.reduce_by_key(kc, vc, P, op, thr)
        for p, k, v in parts:
            res.parts[p] = (k.cpu().numpy().tolist(), v.cpu().numpy().tolist())
        return res
    from . import strings
    return strings.reduce_by_key_bytes(splits, kk, P, thr, op, dev, res)


def _check_int_sum_range(splits, vkinds, op):
    """The reference adds Python big ints; the device accumulates in int64.  A cheap sufficient check on the ingested
    columns: if the sum of |v| over the whole shuffle stays below 2^63 no key's sum can wrap."""
    if vkinds == {columnar.VAL_I64} and op == "sum":
        bound = sum(float(np.abs(c.vals.astype(np.float64)).sum()) for c in splits if c.n)
        if bound >= 2.0 ** 63:
            raise OverflowError("reduceByKey(add): the values' magnitudes sum to %.3g >= 2^63; int64 accumulation on the "
                                "B200 path could wrap where the reference's big ints do not" % bound)


def _run_group(splits, P, thr, dev):
    from . import grouping
    return grouping.group_by_key(splits, P, thr, dev, ShuffleResult(P))
