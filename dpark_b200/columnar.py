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

KEY_I64, KEY_F64, KEY_STR, KEY_BYTES, KEY_TUPLE = "i64", "f64", "str", "bytes", "tuple"
VAL_I64, VAL_F64, VAL_OBJ = "i64", "f64", "obj"


class Columns(object):
    """One split's rows in columnar form (host side, numpy)."""
    __slots__ = ("n", "key_kind", "keys", "key_offsets", "val_kind", "vals", "objs", "key_objs")

    def __init__(self, n, key_kind, keys, key_offsets, val_kind, vals, objs=None, key_objs=None):
        self.n, self.key_kind, self.keys, self.key_offsets = n, key_kind, keys, key_offsets
        self.val_kind, self.vals, self.objs = val_kind, vals, objs
        # KEY_TUPLE: the Python keys themselves (the byte column only carries their IDENTITY; the hash is computed
        # on the device from the leaf columns, and a distinct key is handed back as the object it came in as)
        self.key_objs = key_objs


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
        data, offs = _tuple_identity_bytes(keys)
        return KEY_TUPLE, data, offs
    raise _unhashable(t)


# ---- tuple / None keys (dpark/portable_hash.pyx:3-15, 53-54) -------------------------------------------------------
# Identity: a canonical byte string per key (ints 8 bytes, floats their canonical 8 bytes, str/bytes length-prefixed,
# nested tuples concatenated) compared by VALUE on the device (dpk_dict_encode), exactly like str keys.  All keys of
# one shuffle must have the same shape (arity and leaf types position by position): then the encoding is injective and
# Python's cross-type equalities (1 == 1.0) cannot occur between two keys.
# Hash: computed on the device from the leaf columns (dpk_hash_keys / dpk_hash_bytes per leaf, dpk_hash_tuple per
# tuple node), never on the host.
def _shape_of(k):
    t = type(k)
    if t is tuple:
        return ("t",) + tuple(_shape_of(x) for x in k)
    if k is None:
        return ("n",)
    if t is int or (isinstance(k, np.integer)):
        return ("i",)
    if t is float or isinstance(k, np.floating):
        return ("f",)
    if t is str:
        return ("s",)
    if t is bytes:
        return ("b",)
    raise _unhashable(t)


def _leaves(shape, keys, out):
    """Column-wise leaves of same-shaped keys, depth first: out gets (kind, python list) per leaf."""
    if shape[0] == "t":
        for a, sub in enumerate(shape[1:]):
            _leaves(sub, [k[a] for k in keys], out)
    else:
        out.append((shape[0], keys))


def _check_shapes(keys):
    shape = _shape_of(keys[0])
    for k in keys:
        if _shape_of(k) != shape:
            raise TypeError("tuple / None keys of one shuffle must all have the same shape on the B200 path "
                            "(%r vs %r)" % (keys[0], k))
    return shape


def _tuple_identity_bytes(keys):
    import struct
    shape = _check_shapes(keys)
    leaves = []
    _leaves(shape, keys, leaves)
    parts = [[] for _ in keys]
    for kind, col in leaves:
        if kind == "i":
            try:
                raw = np.array(col, dtype=np.int64).tobytes()
            except OverflowError:
                raise TypeError("int keys beyond int64 are not supported on the B200 path")
            for i in range(len(col)):
                parts[i].append(raw[8 * i:8 * i + 8])
        elif kind == "f":
            arr = np.array(col, dtype=np.float64) + 0.0
            if np.isnan(arr).any():
                raise TypeError("NaN keys are not supported (CPython hashes NaN by identity)")
            raw = arr.tobytes()
            for i in range(len(col)):
                parts[i].append(raw[8 * i:8 * i + 8])
        elif kind in ("s", "b"):
            for i, x in enumerate(col):
                b = x.encode("utf-8", "surrogatepass") if kind == "s" else x
                parts[i].append(struct.pack("<q", len(b)) + b)
    blobs = [b"".join(p) for p in parts]
    offs = np.zeros(len(blobs) + 1, dtype=np.int64)
    np.cumsum(np.fromiter(map(len, blobs), dtype=np.int64, count=len(blobs)), out=offs[1:])
    return np.frombuffer(b"".join(blobs), dtype=np.uint8), offs


def tuple_hashes_on_device(keys, dev):
    """portable_hash of same-shaped tuple / None keys, evaluated by the CUDA kernels: int64 device tensor."""
    from . import _native as nv
    shape = _check_shapes(keys)

    def node(shape, col):
        n = len(col)
        kind = shape[0]
        if kind == "t":
            items = [node(sub, [k[a] for k in col]) for a, sub in enumerate(shape[1:])]
            mat = torch.stack(items) if items else torch.empty((0, n), dtype=torch.int64, device=dev)
            return nv.hash_tuple(mat.contiguous())
        if kind == "n":
            return torch.full((n,), 1315925605, dtype=torch.int64, device=dev)      # portable_hash.pyx:53-54
        if kind == "i":
            return nv.hash_keys(torch.from_numpy(np.array(col, dtype=np.int64)).to(dev))
        if kind == "f":
            return nv.hash_keys(torch.from_numpy(np.array(col, dtype=np.float64)).to(dev))
        blobs = [x.encode("utf-8", "surrogatepass") for x in col] if kind == "s" else col
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.fromiter(map(len, blobs), dtype=np.int64, count=n), out=offs[1:])
        data = np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8)
        return nv.hash_bytes(torch.from_numpy(data.copy()).to(dev), torch.from_numpy(offs).to(dev),
                             nv.STR_UTF8 if kind == "s" else nv.BYTES_SIGNED)
    return node(shape, keys)


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
    return Columns(len(keys), kk, kd, ko, vk, vd, objs, keys if kk == KEY_TUPLE else None)


def decode_keys(kind, data, offsets=None, key_objs=None):
    """Key column (numpy) -> list of Python keys."""
    if kind == KEY_TUPLE:
        return list(key_objs)
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
    if kk == KEY_TUPLE:
        return tuple_hashes_on_device(keys, dev)
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
