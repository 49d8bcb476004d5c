"""Legacy shuffle wire format, against bucket files written by the reference itself
(tests/golden/make_wire_golden.py: BucketDumper._prepare + pack_header, lz4framed stubbed by zlib level 1 -- the image
has neither lz4framed nor snappy, so the zlib-flagged form is what can be verified byte for byte)."""
import pytest

from dpark_b200 import wire
from tests.golden_util import dec, load

CASES = load("wire_cases.json")["cases"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reads_what_the_reference_wrote_and_writes_the_same_bytes(case):
    data = bytes.fromhex(case["bytes"])
    segs = wire.unpack_segments(data, "zlib")
    if case["name"] == "unmarshalable":
        assert segs == [[(1, frozenset([1, 2]))]] and data[:1] == b"p"
        assert wire.pack_segment([(1, frozenset([1, 2]))], "zlib")[:1] == b"p"
        return
    want = [(dec(k), dec(v)) for k, v in case["items"]]
    got = [tuple(x) for s in segs for x in s]
    assert got == want
    assert data[:1] == (b"m" if case["is_marshal"] else b"p")
    if case["name"] != "two_segments":
        mine = wire.pack_segment(want, "zlib")
        assert wire.unpack_segments(mine, "zlib") == segs and mine[:1] == data[:1]
        if case["name"] in ("ints", "empty"):
            # byte-identical to the reference's file (with str / tuple objects marshal sets its FLAG_REF bits from the
            # objects' reference counts, so two dumps of equal values need not agree byte for byte -- also inside the
            # reference itself; the decoded content is what the reader sees)
            assert mine == data
        k, v = wire.load_partition_rows(data, "zlib")
        assert list(zip(k, v)) == want


def test_truncated_and_unknown_segments_fail_like_the_reference():
    data = wire.pack_segment([(1, 2)], "zlib")
    with pytest.raises(IOError):
        wire.unpack_segments(data[:-1], "zlib")
    with pytest.raises(IOError):
        wire.unpack_segments(data + b"m\x01", "zlib")
    with pytest.raises(KeyError):
        wire.unpack_segments(b"x" + data[1:], "zlib")
    for codec in ("lz4", "snappy"):                               # absent here: must say so, not fall back
        with pytest.raises(ImportError):
            wire.pack_segment([(1, 2)], codec)


def test_columns_round_trip():
    import numpy as np
    k = np.array([3, -1, 2 ** 40], dtype=np.int64)
    v = np.array([1.5, 2.5, -0.0])
    data = wire.dump_partition_columns(k, v)
    kk, vv = wire.load_partition_rows(data)
    assert kk == k.tolist() and vv == v.tolist()
    assert wire.unpack_segments(wire.dump_partition_columns([], [])) == [[]]
