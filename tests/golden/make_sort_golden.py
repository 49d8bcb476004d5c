#!/usr/bin/env python
"""Golden vectors for RDD.sort FROM THE REAL REFERENCE (dpark/rdd.py:273-287, RangePartitioner
dpark/dependency.py:242-258); same out-of-tree build as make_golden.py.

    python tests/golden/make_sort_golden.py         # writes tests/golden/sort_cases.json

Stored: the exact content AND order of every output partition (the sample-derived range bounds decide the layout)."""
import json
import logging
import os
import random
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import bootstrap, build_reference, enc  # noqa: E402

KEYS = {"id": lambda x: x, "neg": lambda x: -x, "second": lambda x: x[1], "mod": lambda x: (x % 10, x)}


def generate():
    from dpark import DparkContext
    logging.getLogger("dpark").setLevel(logging.ERROR)
    dc = DparkContext("local")
    dc.init()
    rnd = random.Random(31)
    cases = []

    def case(name, xs, M, key, reverse, P):
        out = dc.parallelize(xs, M).sort(key=KEYS[key], reverse=reverse, numSplits=P).glom().collect()
        cases.append({"name": name, "xs": [enc(x) for x in xs], "M": M, "key": key, "reverse": reverse, "P": P,
                      "parts": [[enc(x) for x in part] for part in out]})

    xs = [rnd.randrange(-1000, 1000) for _ in range(500)]
    for reverse in (False, True):
        case("ints_%s" % ("rev" if reverse else "fwd"), xs, 5, "id", reverse, 4)
        case("ints_modkey_%s" % ("rev" if reverse else "fwd"), xs, 3, "mod", reverse, None)
    case("ints_negkey", xs, 4, "neg", False, 3)
    pairs = [("w%d" % rnd.randrange(50), rnd.randrange(100)) for _ in range(300)]
    case("pairs_by_second", pairs, 6, "second", False, 5)
    case("single_partition", xs[:50], 1, "id", True, 4)
    case("more_splits_than_samples", xs[:30], 2, "id", False, 8)
    words = ["k%03d" % rnd.randrange(400) for _ in range(300)]
    case("strings", words, 4, "id", False, 3)
    json.dump({"cases": cases}, open(os.path.join(HERE, "sort_cases.json"), "w"), separators=(",", ":"))
    dc.stop()
    print("wrote", len(cases), "sort cases")


def main():
    scratch = tempfile.mkdtemp(prefix="dpark_ref_")
    try:
        build_reference(scratch)
        bootstrap(scratch)
        generate()
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
