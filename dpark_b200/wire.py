"""Legacy shuffle wire / on-disk format: the segments BucketDumper writes and
RemoteFile.unsorted_batches reads (dpark/task.py:332-353, dpark/shuffle.py:35-65, 247-289), so that the GPU map side
can feed the reference's CPU reducers and the GPU reduce side can ingest buckets written by the reference's CPU map
tasks.

A bucket file is a sequence of segments

    flag (1 byte: b'M' marshal+sorted, b'P' pickle+sorted, b'm' marshal+unsorted, b'p' pickle+unsorted)
    length (uint32, native byte order, of the compressed blob)
    blob = compress(marshal.dumps(items) | pickle.dumps(items, -1)),  items = list of (key, combiner) pairs

`compress` in the reference is lz4framed, else python-snappy (dpark/utils/__init__.py:44-58).  Neither library is in
this image, so the codec is a PARAMETER here: `codec=("lz4"|"snappy")` uses the library when importable and raises
otherwise; `codec="zlib"` (zlib level 1 -- the stand-in the reference is run with in this repository's golden
generators) is what the tests can verify byte for byte against segments written by the reference itself.  This is host
(de)serialisation by nature -- marshal and pickle are CPython codecs -- and sits OUTSIDE the hot path: the GPU shuffle
never serialises (the bucket-major HBM buffer is the send buffer).
"""
import marshal
import pickle
import struct
import zlib

F_MAPPING = {(True, True): b"M", (False, True): b"P", (True, False): b"m", (False, False): b"p"}
F_MAPPING_R = dict((v, k) for k, v in F_MAPPING.items())


def _codec(name):
    if name == "zlib":
        return (lambda s: zlib.compress(s, 1)), zlib.decompress
    if name == "lz4":
        import lz4framed                      # noqa: F401  (absent in this image)
        return lz4framed.compress, lz4framed.decompress
    if name == "snappy":
        import snappy                         # noqa: F401  (absent in this image)
        return snappy.compress, snappy.decompress
    raise ValueError("codec must be 'lz4', 'snappy' or 'zlib'")


def marshalable(o, _depth=0):
    """dpark/serialize.py:193-209 (first 100 elements of containers are inspected, like the reference)."""
    import itertools
    if o is None:
        return True
    t = type(o)
    if t in (bytes, str, bool, int, float, complex):
        return True
    if t in (tuple, list, set):
        return all(marshalable(i) for i in itertools.islice(o, 100))
    if t is dict:
        return all(marshalable(k) and marshalable(v) for k, v in itertools.islice(o.items(), 100))
    return False


def pack_segment(items, codec="zlib", is_sorted=False):
    """One segment for the (key, combiner) pairs of a bucket, as BucketDumper._prepare + _dump_bucket produce it."""
    items = list(items)
    comp, _ = _codec(codec)
    try:
        if marshalable(items):
            is_marshal, d = True, marshal.dumps(items)
        else:
            is_marshal, d = False, pickle.dumps(items, -1)
    except ValueError:
        is_marshal, d = False, pickle.dumps(items, -1)
    blob = comp(d)
    return F_MAPPING[(is_marshal, is_sorted)] + struct.pack("I", len(blob)) + blob


def unpack_segments(data, codec="zlib"):
    """All segments of a bucket file -> list of item lists (RemoteFile.unsorted_batches, dpark/shuffle.py:247-289,
    including its length checks)."""
    _, decomp = _codec(codec)
    out, at = [], 0
    while at < len(data):
        head = data[at:at + 5]
        if len(head) != 5:
            raise IOError("fetch bad head length %d" % len(head))
        is_marshal, is_sorted = F_MAPPING_R[head[:1]]
        length, = struct.unpack("I", head[1:5])
        blob = data[at + 5:at + 5 + length]
        if len(blob) != length:
            raise IOError("length not match: expected %d, but got %d" % (length, len(blob)))
        d = decomp(blob)
        out.append(marshal.loads(d) if is_marshal else pickle.loads(d))
        at += 5 + length
    return out


def dump_partition_columns(keys, vals, codec="zlib"):
    """Bucket of the GPU map side (two host columns: numpy arrays or lists, distinct keys after a map-side combine or
    raw rows) -> the bytes of the reference's bucket file for it.  An empty bucket is one segment holding []
    (dpark/task.py:301-304)."""
    ks = keys.tolist() if hasattr(keys, "tolist") else list(keys)
    vs = vals.tolist() if hasattr(vals, "tolist") else list(vals)
    return pack_segment(list(zip(ks, vs)), codec)


def load_partition_rows(data, codec="zlib"):
    """The reference's bucket file -> (keys, combiners) lists in file order (segments concatenated): what
    dpark_b200.columnar.ingest_pairs takes on the GPU reduce side."""
    keys, vals = [], []
    for items in unpack_segments(data, codec):
        for k, v in items:
            keys.append(k)
            vals.append(v)
    return keys, vals
