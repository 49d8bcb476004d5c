#!/usr/bin/env python
"""Golden bucket files of the legacy wire format, written BY THE REFERENCE (BucketDumper._prepare / pack_header,
dpark/task.py:332-353, dpark/shuffle.py:35-65) in this container with the same out-of-tree bootstrap as make_golden.py
(lz4framed stubbed by zlib level 1).  Rewrites tests/golden/wire_cases.json."""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.make_golden import bootstrap, build_reference, enc  # noqa: E402


def main():
    scratch = tempfile.mkdtemp(prefix="dpark_ref_")
    try:
        build_reference(scratch)
        bootstrap(scratch)
        from dpark.shuffle import pack_header, unpack_header
        from dpark.task import BucketDumper
        from dpark.utils import compress  # noqa: F401
        cases = []
        buckets = {
            "ints": [(-4, 1), (0, 1), (2 ** 40, -7)],
            "empty": [],
            "strs": [("w1", 3), ("你好", 1)],
            "floats": [(1.5, 2.25), (0.0, -1.0)],
            "lists": [(7, [1, 2, 3]), (8, [])],
            "tuples": [((1, "a"), 5), ((2, "b"), 6)],
            "unmarshalable": [(1, frozenset([1, 2]))],
        }
        d = BucketDumper.__new__(BucketDumper)
        for name, items in buckets.items():
            (is_marshal, data), size = d._prepare(iter(items))
            blob = pack_header(len(data), is_marshal, False) + data
            assert unpack_header(blob[:5])[0] == len(data)
            cases.append({"name": name, "items": [[enc(k), enc(v)] if name != "unmarshalable" else None for k, v in items],
                          "is_marshal": is_marshal, "bytes": blob.hex()})
        two = bytes.fromhex(cases[0]["bytes"]) + bytes.fromhex(cases[2]["bytes"])
        cases.append({"name": "two_segments", "items": cases[0]["items"] + cases[2]["items"], "is_marshal": True,
                      "bytes": two.hex()})
        with open(os.path.join(HERE, "wire_cases.json"), "w") as f:
            json.dump({"codec": "zlib level 1 (the lz4framed stub of the golden bootstrap)", "cases": cases}, f, indent=1)
        print("wrote", len(cases), "cases")
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
