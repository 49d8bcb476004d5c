"""ctypes binding of libdpark_b200.so (include/dpark_b200.h).

The CUDA extension is the product: if the library is missing or a tensor is not
on a CUDA device this module raises -- there is no CPU fallback for the shuffle
path.  torch is used only for device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdpark_b200.so")

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA = 0, -1, -2, -3, -4
K_HASHED = -1
K_I64, K_I32, K_F64, K_U64, K_F32, K_ROWID = 0, 1, 2, 3, 4, 5
V_I64, V_F64, V_I32, V_F32 = 0, 1, 2, 3
OPS = {"sum": 0, "min": 1, "max": 2, "prod": 3, "and": 4, "or": 5, "xor": 6}
BYTES_SIGNED, STR_UTF8 = 0, 1
MAX_PARTITIONS = 4096

_KEY_KIND = {torch.int64: K_I64, torch.int32: K_I32, torch.float64: K_F64, torch.float32: K_F32}
if hasattr(torch, "uint64"):
    _KEY_KIND[torch.uint64] = K_U64
_VAL_KIND = {torch.int64: V_I64, torch.float64: V_F64, torch.int32: V_I32, torch.float32: V_F32}

# every symbol include/dpark_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "dpk_abi_version", "dpk_last_error", "dpk_device_info", "dpk_hash_keys", "dpk_hash_bytes",
    "dpk_partition_ids", "dpk_partition_workspace_bytes", "dpk_partition_count",
    "dpk_partition_scatter", "dpk_partition", "dpk_combine_workspace_bytes", "dpk_combine",
    "dpk_launch_count", "dpk_prof_enable", "dpk_prof_count", "dpk_prof_get",
    "dpk_dict_encode_workspace_bytes", "dpk_dict_encode", "dpk_set_option",
    "dpk_key_or", "dpk_radix_pass", "dpk_group_heads_workspace_bytes", "dpk_group_heads", "dpk_gather_i64",
    "dpk_partition_scatter_ptrs", "dpk_copy_segments", "dpk_hash_tuple", "dpk_push_plan", "dpk_push_plan_part", "dpk_pipe_plan", "dpk_fused_plan", "dpk_memcpy_batch",
    "dpk_tokenize_blocks", "dpk_tokenize_count", "dpk_tokenize_emit", "dpk_gather_bytes",
    "dpk_radix_pass_seg_workspace_bytes", "dpk_radix_pass_seg",
]

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "dpark_b200: CUDA extension %s is missing -- build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        i64, i32, vp, ci = C.c_int64, C.c_int32, C.c_void_p, C.c_int
        for name in EXPORTS:
            getattr(L, name).restype = ci
        L.dpk_last_error.restype = C.c_char_p
        L.dpk_partition_workspace_bytes.restype = i64
        L.dpk_combine_workspace_bytes.restype = i64
        L.dpk_launch_count.restype = i64
        L.dpk_device_info.argtypes = [vp]
        L.dpk_hash_keys.argtypes = [vp, ci, i64, vp, vp]
        L.dpk_hash_bytes.argtypes = [vp, vp, i64, ci, vp, vp]
        L.dpk_partition_ids.argtypes = [vp, i64, i32, vp, i32, vp, vp]
        L.dpk_hash_tuple.argtypes = [vp, i64, i32, vp, vp]
        L.dpk_partition_workspace_bytes.argtypes = [i64, i32]
        L.dpk_partition_count.argtypes = [vp, ci, vp, i64, i32, vp, i32, i32, vp, vp, i64, vp]
        L.dpk_partition_scatter.argtypes = [vp, ci, vp, vp, i32, i64, i32, vp, i32, i32, vp, vp, vp, vp, i64, vp]
        L.dpk_partition.argtypes = [vp, ci, vp, vp, i32, i64, i32, vp, i32, i32, vp, vp, vp, vp, i64, vp]
        L.dpk_partition_scatter_ptrs.argtypes = [vp, ci, vp, vp, i32, i64, i32, vp, i32, i32, vp, vp, vp, i64, vp]
        L.dpk_copy_segments.argtypes = [vp, vp, vp, i32, vp]
        L.dpk_push_plan.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, C.c_uint64, C.c_uint64, vp, i32, i32, i64, vp, vp, vp, vp, vp, vp]
        L.dpk_push_plan_part.argtypes = [vp, i32, i32, i32, i32, i32, i32, i64, i32, i32, i32, C.c_uint64, C.c_uint64, vp, i32, i32, i64, vp, vp, vp, vp, vp, vp]
        L.dpk_pipe_plan.argtypes = [vp, i32, i32, i32, i32, i32, i64, i32, i32, i32, C.c_uint64, C.c_uint64, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]
        L.dpk_memcpy_batch.argtypes = [vp, vp, vp, i32, vp]
        L.dpk_fused_plan.argtypes = [vp, i32, i32, i32, i32, i32, vp, i32, i32, i64, C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp]
        L.dpk_tokenize_blocks.restype = i64
        L.dpk_tokenize_blocks.argtypes = [i64]
        L.dpk_tokenize_count.argtypes = [vp, i64, vp, vp, vp]
        L.dpk_tokenize_emit.argtypes = [vp, i64, vp, vp, vp, vp]
        L.dpk_gather_bytes.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp]
        L.dpk_combine_workspace_bytes.argtypes = [i64, i32, i32]
        L.dpk_combine.argtypes = [vp, ci, vp, vp, ci, i64, ci, i32, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp,
                                  vp, vp, i64, vp]
        L.dpk_set_option.argtypes = [C.c_char_p, i64]
        L.dpk_key_or.argtypes = [vp, i64, vp, vp]
        L.dpk_gather_i64.argtypes = [vp, vp, i64, vp, vp]
        L.dpk_radix_pass.argtypes = [vp, vp, i32, i64, i32, i32, vp, vp, vp, i64, vp]
        L.dpk_radix_pass_seg_workspace_bytes.restype = i64
        L.dpk_radix_pass_seg_workspace_bytes.argtypes = [i64, i32, i32, i32]
        L.dpk_radix_pass_seg.argtypes = [vp, vp, i32, i64, i32, i32, i32, i32, vp, vp, vp, vp, vp, i64, vp]
        L.dpk_group_heads_workspace_bytes.restype = i64
        L.dpk_group_heads_workspace_bytes.argtypes = [i64]
        L.dpk_group_heads.argtypes = [vp, i64, vp, vp, vp, vp, i64, vp]
        L.dpk_dict_encode_workspace_bytes.restype = i64
        L.dpk_dict_encode_workspace_bytes.argtypes = [i64]
        L.dpk_dict_encode.argtypes = [vp, vp, vp, i64, vp, vp, i64, vp]
        L.dpk_prof_enable.argtypes = [ci]
        L.dpk_prof_get.argtypes = [ci, C.c_char_p, C.POINTER(C.c_float)]
        if L.dpk_abi_version() != 1:
            raise ImportError("dpark_b200: ABI version mismatch")
        _lib = L
        # DPK_OPTIONS="name=value,..." applies dpk_set_option switches at load (A/B runs of whole test suites)
        for item in filter(None, os.environ.get("DPK_OPTIONS", "").split(",")):
            name, _, value = item.partition("=")
            _check(L.dpk_set_option(name.strip().encode(), int(value)))
    return _lib


def _check(rc):
    if rc == OK:
        return
    msg = lib().dpk_last_error().decode("utf-8", "replace")
    if rc == ERR_UNSUPPORTED:
        raise TypeError(msg)
    if rc == ERR_INVALID:
        raise ValueError(msg)
    raise NativeError("dpark_b200 native error %d: %s" % (rc, msg))


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise NativeError("dpark_b200 shuffle kernels need CUDA tensors (no CPU fallback); got %s" % t.device)
        if t is not None and not t.is_contiguous():
            raise ValueError("columns must be contiguous")


def key_kind(t, prehashed=False):
    if prehashed:
        if t.dtype != torch.int64:
            raise TypeError("prehashed keys must be int64")
        return K_HASHED
    try:
        return _KEY_KIND[t.dtype]
    except KeyError:
        raise TypeError("%s is unhashable by portable_hash" % t.dtype)


def val_kind(t):
    try:
        return _VAL_KIND[t.dtype]
    except KeyError:
        raise TypeError("unsupported value dtype %s" % t.dtype)


def _thr(thresholds, device):
    if thresholds is None:
        return None, 0
    if not torch.is_tensor(thresholds):
        thresholds = torch.tensor(list(thresholds), dtype=torch.int64, device=device)
    thresholds = thresholds.to(device=device, dtype=torch.int64).contiguous()
    return thresholds, int(thresholds.numel())


def device_info():
    import numpy as np
    info = np.zeros(4, dtype=np.int32)
    _check(lib().dpk_device_info(info.ctypes.data_as(C.c_void_p)))
    return {"sm_count": int(info[0]), "cc": (int(info[1]), int(info[2])), "l2_mb": int(info[3])}


# ---- a1 / a2 -------------------------------------------------------------------
def hash_keys(keys):
    """portable_hash of a key column (dpark/portable_hash.pyx:51-70)."""
    _need_cuda(keys)
    out = torch.empty(keys.numel(), dtype=torch.int64, device=keys.device)
    _check(lib().dpk_hash_keys(_ptr(keys), key_kind(keys), keys.numel(), _ptr(out), _stream()))
    return out


def hash_bytes(data, offsets, mode):
    _need_cuda(data, offsets)
    n = offsets.numel() - 1
    out = torch.empty(n, dtype=torch.int64, device=offsets.device)
    _check(lib().dpk_hash_bytes(_ptr(data), _ptr(offsets), n, mode, _ptr(out), _stream()))
    return out


def hash_tuple(item_hashes):
    """tuple_hash (dpark/portable_hash.pyx:3-15) of n rows from their items' hashes: int64 tensor [arity, n]."""
    _need_cuda(item_hashes)
    arity, n = int(item_hashes.shape[0]), int(item_hashes.shape[1])
    out = torch.empty(n, dtype=torch.int64, device=item_hashes.device)
    _check(lib().dpk_hash_tuple(_ptr(item_hashes), n, arity, _ptr(out), _stream()))
    return out


def partition_ids(hashes, P, thresholds=None):
    """HashPartitioner.getPartition over a hash column (dpark/dependency.py:229-233)."""
    _need_cuda(hashes)
    thr, nthr = _thr(thresholds, hashes.device)
    out = torch.empty(hashes.numel(), dtype=torch.int32, device=hashes.device)
    _check(lib().dpk_partition_ids(_ptr(hashes), hashes.numel(), P, _ptr(thr), nthr, _ptr(out), _stream()))
    return out


# ---- a4: map side --------------------------------------------------------------
def partition_workspace(nbuckets, device):
    nbytes = lib().dpk_partition_workspace_bytes(0, nbuckets)
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


K_UNORDERED = 0x100   # DPK_K_UNORDERED: rows of a bucket may come out in any order


def _kk(keys, prehashed, row_hash, unordered=False):
    kk = K_ROWID if row_hash is not None else key_kind(keys, prehashed)
    return kk | K_UNORDERED if (unordered and kk >= 0) else kk


def partition_count(keys, P, thresholds=None, prehashed=False, sub_bits=0, ws=None, row_hash=None, unordered=False):
    """Rows per bucket of one chunk; returns (counts[P << sub_bits] int64 device, ws)."""
    _need_cuda(keys, row_hash)
    F = P << sub_bits
    thr, nthr = _thr(thresholds, keys.device)
    if ws is None:
        ws = partition_workspace(F, keys.device)
    counts = torch.empty(F, dtype=torch.int64, device=keys.device)
    _check(lib().dpk_partition_count(_ptr(keys), _kk(keys, prehashed, row_hash, unordered), _ptr(row_hash), keys.numel(), P,
                                     _ptr(thr), nthr,
                                     sub_bits, _ptr(counts), _ptr(ws), ws.numel(), _stream()))
    return counts, ws


def partition_scatter(keys, vals, P, bucket_base, out_keys, out_vals, ws, thresholds=None, prehashed=False,
                      sub_bits=0, row_hash=None, unordered=False):
    _need_cuda(keys, vals, bucket_base, out_keys, out_vals, ws, row_hash)
    thr, nthr = _thr(thresholds, keys.device)
    vb = 0 if vals is None else vals.element_size()
    _check(lib().dpk_partition_scatter(_ptr(keys), _kk(keys, prehashed, row_hash, unordered), _ptr(row_hash), _ptr(vals), vb,
                                       keys.numel(), P,
                                       _ptr(thr), nthr, sub_bits, _ptr(bucket_base), _ptr(out_keys),
                                       _ptr(out_vals), _ptr(ws), ws.numel(), _stream()))


def partition_scatter_ptrs(keys, vals, P, key_ptrs, val_ptrs, ws, thresholds=None, prehashed=False, sub_bits=0,
                           row_hash=None, unordered=False):
    """Fused scatter + exchange: bucket b of this chunk is written through key_ptrs[b] / val_ptrs[b]
    (device int64 tensors holding absolute device addresses, possibly peer-GPU memory)."""
    _need_cuda(keys, vals, key_ptrs, val_ptrs, ws, row_hash)
    thr, nthr = _thr(thresholds, keys.device)
    vb = 0 if vals is None else vals.element_size()
    _check(lib().dpk_partition_scatter_ptrs(_ptr(keys), _kk(keys, prehashed, row_hash, unordered), _ptr(row_hash), _ptr(vals),
                                            vb, keys.numel(), P, _ptr(thr), nthr, sub_bits, _ptr(key_ptrs),
                                            _ptr(val_ptrs), _ptr(ws), ws.numel(), _stream()))


def copy_segments(src_ptrs, dst_ptrs, nbytes, sms=0):
    """One launch copying nbytes[s] bytes from device address src_ptrs[s] to dst_ptrs[s] (int64 device
    tensors; destinations may be peer-GPU memory): the exchange as block pushes over NVLink.
    sms > 0: the copy runs on that many whole SMs only (dpk_set_option "copy_sms") -- for a push that overlaps
    another kernel."""
    _need_cuda(src_ptrs, dst_ptrs, nbytes)
    if not (src_ptrs.numel() == dst_ptrs.numel() == nbytes.numel()):
        raise ValueError("segment table columns differ in length")
    if sms:
        set_option("copy_sms", int(sms))
    try:
        _check(lib().dpk_copy_segments(_ptr(src_ptrs), _ptr(dst_ptrs), _ptr(nbytes), nbytes.numel(), _stream()))
    finally:
        if sms:
            set_option("copy_sms", 0)


def memcpy_batch(dst_ptrs, src_ptrs, nbytes):
    """Block copies by the copy engines: one cudaMemcpyBatchAsync over HOST lists of device addresses and sizes."""
    n = len(nbytes)
    if not (len(dst_ptrs) == len(src_ptrs) == n):
        raise ValueError("segment table columns differ in length")
    if n == 0:
        return
    U, I = C.c_uint64 * n, C.c_int64 * n
    _check(lib().dpk_memcpy_batch(U(*dst_ptrs), U(*src_ptrs), I(*nbytes), n, _stream()))


def push_plan(all_counts, nranks, per_block, my_src, my_rank, keys, vals, dst_base, capacity, need_over, want_seg=True,
              part=None, dst_row0=0):
    """dpk_push_plan(_part): (src_ptrs, dst_ptrs, nbytes [ncols * nranks], seg [nsrc, own buckets] | None), one launch.
    keys / vals: my bucket-major columns (vals may be None); dst_base: device int64 [ncols * nranks] receive-buffer
    addresses (column-major).  part = (blk_lo, blk_hi): only those buckets of every destination's block, into the
    region of `capacity` rows starting at row dst_row0 of the receive buffers."""
    _need_cuda(all_counts, dst_base, need_over)
    nsrc, F = int(all_counts.shape[0]), int(all_counts.shape[1])
    ncols = 1 if vals is None else 2
    dev = all_counts.device
    lo, hi = (0, per_block) if part is None else part
    src = torch.empty(ncols * nranks, dtype=torch.int64, device=dev)
    dst = torch.empty(ncols * nranks, dtype=torch.int64, device=dev)
    nby = torch.empty(ncols * nranks, dtype=torch.int64, device=dev)
    b0, b1 = min(F, my_rank * per_block + lo), min(F, (my_rank + 1) * per_block, my_rank * per_block + hi)
    seg = torch.empty((nsrc, max(0, b1 - b0)), dtype=torch.int64, device=dev) if want_seg else None
    _check(lib().dpk_push_plan_part(_ptr(all_counts), nsrc, nranks, F, per_block, lo, hi, dst_row0, my_src, my_rank, ncols,
                                    C.c_uint64(keys.data_ptr()), C.c_uint64(0 if vals is None else vals.data_ptr()),
                                    _ptr(dst_base), keys.element_size(), 0 if vals is None else vals.element_size(),
                                    capacity, _ptr(src), _ptr(dst), _ptr(nby), _ptr(need_over), _ptr(seg), _stream()))
    return src, dst, nby, seg


def pipe_plan(all_counts, nranks, per_block, nparts, region_rows, my_src, my_rank, keys, vals, dst_base, need_over,
              want_seg=True):
    """dpk_pipe_plan: (bucket_base[F], src_ptrs, dst_ptrs, nbytes [nparts, ncols * nranks], seg [nparts, nsrc, per_block /
    nparts] | None) for one group of map splits; keys / vals: the group's SEND buffers (allocated with pipe_pad_rows()
    spare rows)."""
    _need_cuda(all_counts, dst_base, need_over, keys, vals)
    nsrc, F = int(all_counts.shape[0]), int(all_counts.shape[1])
    ncols = 1 if vals is None else 2
    dev = all_counts.device
    base = torch.empty(F, dtype=torch.int64, device=dev)
    src = torch.empty((nparts, ncols * nranks), dtype=torch.int64, device=dev)
    dst = torch.empty((nparts, ncols * nranks), dtype=torch.int64, device=dev)
    nby = torch.empty((nparts, ncols * nranks), dtype=torch.int64, device=dev)
    seg = torch.empty((nparts, nsrc, per_block // nparts), dtype=torch.int64, device=dev) if want_seg else None
    _check(lib().dpk_pipe_plan(_ptr(all_counts), nsrc, nranks, F, per_block, nparts, region_rows, my_src, my_rank, ncols,
                               C.c_uint64(keys.data_ptr()), C.c_uint64(0 if vals is None else vals.data_ptr()),
                               _ptr(dst_base), keys.element_size(), 0 if vals is None else vals.element_size(),
                               _ptr(base), _ptr(src), _ptr(dst), _ptr(nby), _ptr(need_over), _ptr(seg), _stream()))
    return base, src, dst, nby, seg


def pipe_pad_rows(nranks, nparts, key_bytes, val_bytes=None):
    """Spare rows a pipelined step's send buffer needs for the congruence pads (dpk_pipe_plan)."""
    return nranks * nparts * (16 // min(key_bytes, val_bytes or key_bytes))


def fused_plan(all_counts, nranks, per_block, my_rank, dst_base, key_bytes, val_bytes, capacity, dump_keys, dump_vals,
               need_over, want_seg=True):
    """dpk_fused_plan: (key_ptrs[F], val_ptrs[F] | None, seg [nranks, own buckets] | None), one launch: where every
    bucket of this rank's map output goes in its owner's receive buffer (see include/dpark_b200.h)."""
    _need_cuda(all_counts, dst_base, need_over, dump_keys, dump_vals)
    G, F = int(all_counts.shape[0]), int(all_counts.shape[1])
    if G != nranks:
        raise ValueError("all_counts has %d source rows for %d ranks" % (G, nranks))
    ncols = 1 if dump_vals is None else 2
    dev = all_counts.device
    kp = torch.empty(F, dtype=torch.int64, device=dev)
    vp_ = torch.empty(F, dtype=torch.int64, device=dev) if ncols == 2 else None
    b0, b1 = min(F, my_rank * per_block), min(F, (my_rank + 1) * per_block)
    seg = torch.empty((G, b1 - b0), dtype=torch.int64, device=dev) if want_seg else None
    _check(lib().dpk_fused_plan(_ptr(all_counts), nranks, F, per_block, my_rank, ncols, _ptr(dst_base), key_bytes,
                                val_bytes if ncols == 2 else 0, capacity, C.c_uint64(dump_keys.data_ptr()),
                                C.c_uint64(0 if dump_vals is None else dump_vals.data_ptr()), _ptr(kp), _ptr(vp_),
                                _ptr(need_over), _ptr(seg), _stream()))
    return kp, vp_, seg


def partition(keys, vals, P, thresholds=None, prehashed=False, sub_bits=0, row_hash=None, unordered=False):
    """Stable hash-partition of one chunk (ShuffleMapTask._run, dpark/task.py:209-226).
    Returns (out_keys, out_vals, offsets[(P << sub_bits) + 1] int64 device)."""
    _need_cuda(keys, vals, row_hash)
    if vals is not None and vals.numel() != keys.numel():
        from .errors import DparkUserFatalError
        raise DparkUserFatalError("ragged pair columns: %d keys, %d values" % (keys.numel(), vals.numel()))
    F = P << sub_bits
    thr, nthr = _thr(thresholds, keys.device)
    ws = partition_workspace(F, keys.device)
    out_keys = torch.empty_like(keys)
    out_vals = None if vals is None else torch.empty_like(vals)
    offsets = torch.empty(F + 1, dtype=torch.int64, device=keys.device)
    vb = 0 if vals is None else vals.element_size()
    _check(lib().dpk_partition(_ptr(keys), _kk(keys, prehashed, row_hash, unordered), _ptr(row_hash), _ptr(vals), vb,
                               keys.numel(), P,
                               _ptr(thr), nthr, sub_bits, _ptr(out_keys), _ptr(out_vals), _ptr(offsets),
                               _ptr(ws), ws.numel(), _stream()))
    return out_keys, out_vals, offsets


# ---- a9: reduce side -----------------------------------------------------------
def acc_dtype(vals_dtype):
    """Accumulator/output dtype of combine: ints -> int64, floats -> float64
    (the reference adds Python ints / Python floats)."""
    return torch.float64 if vals_dtype in (torch.float32, torch.float64) else torch.int64


def combine(keys, vals, op, P, seg_rows, part_first=0, nparts=None, thresholds=None, sub_bits=0,
            row_hash=None):
    """Reduce-side merge (DiskHashMerger._merge, dpark/shuffle.py:600-608) of the
    rows of partitions [part_first, part_first+nparts).  Rows are laid out
    source-major, bucket-major inside; seg_rows: device int64 [nsrc, nparts <<
    sub_bits] rows of local fine bucket b from source s.  Returns (out_keys,
    out_vals, out_offsets[nparts+1], out_counts[nparts]); partition j's distinct
    keys are out[out_offsets[j] : out_offsets[j] + out_counts[j]].  With row_hash
    (the per-row portable_hash column) the keys are representative row ids from
    dict_encode (DPK_K_ROWID)."""
    _need_cuda(keys, vals, seg_rows, row_hash)
    if nparts is None:
        nparts = P
    n = keys.numel()
    F = nparts << sub_bits
    if seg_rows.dim() == 1:
        seg_rows = seg_rows.unsqueeze(0)
    nsrc = int(seg_rows.shape[0])
    if seg_rows.shape[1] != F or seg_rows.dtype != torch.int64:
        raise ValueError("seg_rows must be int64[nsrc, %d]" % F)
    bucket_rows = seg_rows
    thr, nthr = _thr(thresholds, keys.device)
    ws_bytes = lib().dpk_combine_workspace_bytes(n, F, nsrc)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=keys.device)
    out_keys = torch.empty_like(keys)
    out_vals = torch.empty(n, dtype=acc_dtype(vals.dtype), device=keys.device)
    out_offsets = torch.empty(nparts + 1, dtype=torch.int64, device=keys.device)
    out_counts = torch.empty(nparts, dtype=torch.int64, device=keys.device)
    kk = key_kind(keys) if row_hash is None else K_ROWID
    _check(lib().dpk_combine(_ptr(keys), kk, _ptr(row_hash), _ptr(vals), val_kind(vals), n, OPS[op], P,
                             _ptr(thr), nthr, sub_bits, part_first, nparts, nsrc, _ptr(bucket_rows), _ptr(out_keys),
                             _ptr(out_vals), _ptr(out_offsets), _ptr(out_counts), _ptr(ws), ws_bytes,
                             _stream()))
    return out_keys, out_vals, out_offsets, out_counts


# ---- variable-length keys ----------------------------------------------------------
def dict_encode(data, offsets, hashes):
    """Representative row id per row: rep[i] == rep[j] <=> the byte strings are equal."""
    _need_cuda(data, offsets, hashes)
    n = offsets.numel() - 1
    ws_bytes = lib().dpk_dict_encode_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=offsets.device)
    rep = torch.empty(n, dtype=torch.int64, device=offsets.device)
    _check(lib().dpk_dict_encode(_ptr(data), _ptr(offsets), _ptr(hashes), n, _ptr(rep), _ptr(ws), ws_bytes,
                                 _stream()))
    return rep


# ---- f4: device text ingest --------------------------------------------------------
def tokenize(data):
    """Tokens (str.split() without arguments) of an ASCII byte range on the device: (starts, lens, ascii) -- int64
    device tensors in text order; ascii False = the range holds a byte >= 0x80 and must be tokenised by Python
    (starts / lens are None then).  One host read (the token count sizes the outputs)."""
    _need_cuda(data)
    n = int(data.numel())
    dev = data.device
    if n == 0:
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        return z, z.clone(), True
    nb = int(lib().dpk_tokenize_blocks(n))
    counts = torch.empty(nb + 1, dtype=torch.int64, device=dev)   # [nb] = the high-byte flag
    counts[nb] = 0
    _check(lib().dpk_tokenize_count(_ptr(data), n, _ptr(counts), C.c_void_p(counts.data_ptr() + 8 * nb), _stream()))
    incl = torch.cumsum(counts[:nb], 0)
    total, flag = int(incl[-1].item()), int(counts[nb].item())
    if flag & 1:
        return None, None, False
    base = (incl - counts[:nb]).contiguous()
    starts = torch.empty(total, dtype=torch.int64, device=dev)
    lens = torch.empty(total, dtype=torch.int64, device=dev)
    if total:
        _check(lib().dpk_tokenize_emit(_ptr(data), n, _ptr(base), _ptr(starts), _ptr(lens), _stream()))
    return starts, lens, True


def gather_bytes(data, starts, lens, idx=None):
    """Selected rows made contiguous: (bytes uint8, offsets int64 [m + 1]) with row i = data[starts[r] : starts[r] +
    lens[r]], r = idx[i] (idx None: every row)."""
    _need_cuda(data, starts, lens, idx)
    sel = lens if idx is None else lens[idx]
    m = int(sel.numel())
    off = torch.zeros(m + 1, dtype=torch.int64, device=data.device)
    if m:
        torch.cumsum(sel, 0, out=off[1:])
    out = torch.empty(int(off[-1].item()) if m else 0, dtype=torch.uint8, device=data.device)
    if m and out.numel():
        _check(lib().dpk_gather_bytes(_ptr(data), _ptr(starts), _ptr(lens), _ptr(idx), m, _ptr(off), _ptr(out), _stream()))
    return out, off


# ---- a10: groupByKey reduce side ---------------------------------------------------
def key_or(keys):
    """Device uint64 (as int64 tensor[1]): OR over i of keys[i] ^ keys[0]."""
    _need_cuda(keys)
    out = torch.empty(1, dtype=torch.int64, device=keys.device)
    _check(lib().dpk_key_or(_ptr(keys), keys.numel(), _ptr(out), _stream()))
    return out


def gather_i64(src, idx):
    _need_cuda(src, idx)
    out = torch.empty(idx.numel(), dtype=torch.int64, device=idx.device)
    _check(lib().dpk_gather_i64(_ptr(src), _ptr(idx), idx.numel(), _ptr(out), _stream()))
    return out


def radix_pass(keys, vals, shift, bits, out_keys=None, out_vals=None, ws=None):
    """One stable LSD radix pass over int64 key bits (the multisplit with digit buckets)."""
    _need_cuda(keys, vals)
    if keys.dtype != torch.int64:
        raise TypeError("radix_pass sorts int64 key bits")
    if out_keys is None:
        out_keys = torch.empty_like(keys)
    if out_vals is None and vals is not None:
        out_vals = torch.empty_like(vals)
    if ws is None:
        ws = partition_workspace(1 << bits, keys.device)
    vb = 0 if vals is None else vals.element_size()
    _check(lib().dpk_radix_pass(_ptr(keys), _ptr(vals), vb, keys.numel(), shift, bits, _ptr(out_keys),
                                _ptr(out_vals), _ptr(ws), ws.numel(), _stream()))
    return out_keys, out_vals


def radix_pass_seg(keys, vals, shift, bits, seg_rows, out_keys=None, out_vals=None):
    """One stable radix pass inside every first-level bucket (dpk_radix_pass_seg).  seg_rows: device int64
    [nsrc, nbuckets].  Returns (out_keys, out_vals)."""
    _need_cuda(keys, vals, seg_rows)
    if keys.dtype != torch.int64:
        raise TypeError("radix_pass_seg sorts int64 key bits")
    nsrc, F = int(seg_rows.shape[0]), int(seg_rows.shape[1])
    n = keys.numel()
    if out_keys is None:
        out_keys = torch.empty_like(keys)
    if out_vals is None and vals is not None:
        out_vals = torch.empty_like(vals)
    ws_bytes = lib().dpk_radix_pass_seg_workspace_bytes(n, F, nsrc, bits)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=keys.device)
    fine = torch.empty((F << bits) + 1, dtype=torch.int64, device=keys.device)
    vb = 0 if vals is None else vals.element_size()
    _check(lib().dpk_radix_pass_seg(_ptr(keys), _ptr(vals), vb, n, shift, bits, F, nsrc, _ptr(seg_rows.contiguous()),
                                    _ptr(out_keys), _ptr(out_vals), _ptr(fine), _ptr(ws), ws_bytes, _stream()))
    return out_keys, out_vals


def group_heads(sorted_keys):
    """CSR heads of a key-sorted column: (group_keys[n], starts[n+1], ngroups[1]) device tensors;
    only the first ngroups (+1) entries are meaningful."""
    _need_cuda(sorted_keys)
    n = sorted_keys.numel()
    ws_bytes = lib().dpk_group_heads_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=sorted_keys.device)
    out_keys = torch.empty(n, dtype=torch.int64, device=sorted_keys.device)
    out_starts = torch.empty(n + 1, dtype=torch.int64, device=sorted_keys.device)
    ng = torch.empty(1, dtype=torch.int64, device=sorted_keys.device)
    _check(lib().dpk_group_heads(_ptr(sorted_keys), n, _ptr(out_keys), _ptr(out_starts), _ptr(ng), _ptr(ws),
                                 ws_bytes, _stream()))
    return out_keys, out_starts, ng


def set_option(name, value):
    _check(lib().dpk_set_option(name.encode(), int(value)))


# ---- measurement hooks ---------------------------------------------------------
def launch_count():
    """Kernels launched by the library since load (exact, counted in C)."""
    return int(lib().dpk_launch_count())


def prof_enable(on=True):
    _check(lib().dpk_prof_enable(1 if on else 0))


def prof_collect():
    """[(kernel label, device ms)] for every launch since prof_enable(True)."""
    L = lib()
    out = []
    name = C.create_string_buffer(64)
    ms = C.c_float()
    for i in range(L.dpk_prof_count()):
        _check(L.dpk_prof_get(i, name, C.byref(ms)))
        out.append((name.value.decode(), float(ms.value)))
    return out
