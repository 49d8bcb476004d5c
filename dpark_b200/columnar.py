"""Columnar partition container: the boundary between Python rows and HBM.

The reference moves Python tuples through dict loops.  Here a partition is a
pair of columns (struct of arrays):

    keys : int64 | float64 array,  or  (uint8 data, int64 offsets) for str/bytes
    vals : int64 | float64 array,  or  row ids into a host-side object list

Python objects exist only at ingest (rows produced by user lambdas upstream of
the shuffle) and egress (rows handed to user lambdas downstream): the shuffle
itself never touches them -- it is not a CPU shuffle.
"""
import numpy as np
import torch

from .errors import DparkUserFatalError

KEY_I64, KEY_F64, KEY_STR, KEY_BYTES = "i64", "f64", "str", "bytes"
VAL_I64, VAL_F64, VAL_OBJ = "i64", "f64", "obj"


class Columns(object):
    """One split's rows in columnar form (host side, numpy)."""
    __slots__ = ("n", "key_kind", "keys", "key_offsets", "val_kind", "vals", "objs")

    def __init__(self, n, key_kind, keys, key_offsets, val_kind, vals, objs=None):
        self.n, self.key_kind, self.keys, self.key_offsets = n, key_kind, keys, key_offsets
        self.val_kind, self.vals, self.objs = val_kind, vals, objs


def _unhashable(t):
    return TypeError("%s is unhashable by portable_hash" % t)


def _key_column(keys):
    kinds = set(map(type, keys))
    if not kinds:
        return KEY_I64, np.empty(0, np.int64), None
    if len(kinds) > 1:
        if kinds <= {int, float} or any(issubclass(t, np.number) for t in kinds):
            raise TypeError("mixed key types %s in one shuffle are not supported on the B200 path"
                            % sorted(t.__name__ for t in kinds))
        raise TypeError("mixed key types %s in one shuffle are not supported on the B200 path"
                        % sorted(t.__name__ for t in kinds))
    t = kinds.pop()
    if t is int:
        try:
            return KEY_I64, np.array(keys, dtype=np.int64), None
        except OverflowError:
            raise TypeError("int keys beyond int64 are not supported on the B200 path")
    if t is float:
        arr = np.array(keys, dtype=np.float64)
        if np.isnan(arr).any():
            raise TypeError("NaN keys are not supported (CPython hashes NaN by identity)")
        return KEY_F64, arr + 0.0, None      # -0.0 and 0.0 are ONE dict key in Python: canonical spelling 0.0 on every path
    if t is str or t is bytes:
        blobs = [k.encode("utf-8", "surrogatepass") for k in keys] if t is str else keys
        offs = np.zeros(len(blobs) + 1, dtype=np.int64)
        np.cumsum(np.fromiter(map(len, blobs), dtype=np.int64, count=len(blobs)), out=offs[1:])
        data = np.frombuffer(b"".join(blobs), dtype=np.uint8)
        return (KEY_STR if t is str else KEY_BYTES), data, offs
    if issubclass(t, np.integer):
        return KEY_I64, np.array(keys, dtype=np.int64), None
    if issubclass(t, np.floating):
        return KEY_F64, np.array(keys, dtype=np.float64) + 0.0, None
    if t is bool or t in (list, dict, set, complex):
        raise _unhashable(t)                      # dpark/portable_hash.pyx:70
    if t is tuple or t is type(None):
        raise TypeError("%s keys are hashable in the reference but not yet supported on the B200 path"
                        % t.__name__)
    raise _unhashable(t)


def _val_column(vals, want_numeric):
    if not want_numeric:
        return VAL_OBJ, np.arange(len(vals), dtype=np.int64), vals
    kinds = set(map(type, vals))
    if not kinds:
        return VAL_I64, np.empty(0, np.int64), None
    if kinds <= {int} or all(issubclass(t, np.integer) for t in kinds):
        try:
            return VAL_I64, np.array(vals, dtype=np.int64), None
        except OverflowError:
            raise TypeError("int values beyond int64 are not supported on the B200 path")
    if kinds <= {float} or all(issubclass(t, np.floating) for t in kinds):
        return VAL_F64, np.array(vals, dtype=np.float64), None
    raise TypeError("reduceByKey values must be all int or all float on the B200 path, got %s"
                    % sorted(t.__name__ for t in kinds))


def ingest_pairs(rows, scope="rdd", numeric_values=True):
    """Python (k, v) rows -> Columns.  Non-pair rows raise DparkUserFatalError
    like dpark/task.py:216-219."""
    keys, vals = [], []
    ka, va = keys.append, vals.append
    for item in rows:
        try:
            k, v = item
        except (TypeError, ValueError) as e:
            raise DparkUserFatalError("item of %s should be (k, v) pair, got: %r, exception: %s"
                                      % (scope, item, e))
        ka(k)
        va(v)
    kk, kd, ko = _key_column(keys)
    vk, vd, objs = _val_column(vals, numeric_values)
    return Columns(len(keys), kk, kd, ko, vk, vd, objs)


def decode_keys(kind, data, offsets=None):
    """Key column (numpy) -> list of Python keys."""
    if kind in (KEY_I64, KEY_F64):
        return data.tolist()
    raw = data.tobytes()
    offs = offsets.tolist()
    if kind == KEY_BYTES:
        return [raw[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    return [raw[offs[i]:offs[i + 1]].decode("utf-8", "surrogatepass") for i in range(len(offs) - 1)]


def _hash_column(keys):
    """portable_hash of every key of a Python list, evaluated by the CUDA kernels
    (dpk_hash_keys / dpk_hash_bytes).  Returns an int64 device tensor."""
    from . import _native as nv
    kk, kd, ko = _key_column(keys)
    if not torch.cuda.is_available():
        raise nv.NativeError("portable_hash needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    if kk in (KEY_I64, KEY_F64):
        return nv.hash_keys(torch.from_numpy(kd).to(dev))
    data = torch.from_numpy(np.ascontiguousarray(kd) if kd.size else np.zeros(1, np.uint8)).to(dev)
    return nv.hash_bytes(data, torch.from_numpy(ko).to(dev), nv.STR_UTF8 if kk == KEY_STR else nv.BYTES_SIGNED)


def hashes_of_keys(keys):
    """[portable_hash(k) for k in keys] as a list of Python ints (device-evaluated)."""
    if not keys:
        return []
    return _hash_column(list(keys)).cpu().tolist()


def partition_of_key(key, P, thresholds=None):
    """HashPartitioner.getPartition(key) for one Python key, evaluated by the CUDA
    kernels (1-row launch)."""
    from . import _native as nv
    h = _hash_column([key])
    thr = None if thresholds is None else torch.tensor(thresholds, dtype=torch.int64, device=h.device)
    return int(nv.partition_ids(h, P, thr)[0])
