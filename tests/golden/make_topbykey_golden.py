#!/usr/bin/env python
"""Golden vectors for topByKey FROM THE REAL REFERENCE (dpark/rdd.py:552-594, HeapAggregator
dpark/dependency.py:164-193); same out-of-tree build as make_golden.py.

    python tests/golden/make_topbykey_golden.py     # writes tests/golden/topbykey_cases.json

The inputs of the reference's own tests (tests/test_rdd.py:353-374) plus seeded random rows with many ties;
order functions are named so that the tests can rebuild them."""
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

ORDER = {"none": None, "first": lambda x: x[0], "mod7": lambda x: x % 7, "neg": lambda x: -x}


def generate():
    from dpark import DparkContext
    logging.getLogger("dpark").setLevel(logging.ERROR)
    dc = DparkContext("local")
    dc.init()
    rnd = random.Random(11)
    cases = []

    def case(name, rows, M, top_n, order, reverse, P):
        out = dc.makeRDD(rows, M).topByKey(top_n=top_n, order_func=ORDER[order], reverse=reverse, num_splits=P)
        parts = [sorted(([enc(k), enc(list(v))] for k, v in part), key=json.dumps) for part in out.glom().collect()]
        cases.append({"name": name, "rows": [[enc(k), enc(v)] for k, v in rows], "M": M, "top_n": top_n,
                      "order": order, "reverse": reverse, "P": P, "parts": parts})

    ks = [1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6]
    ds = list(zip(ks, range(5, 26)))
    case("test_rdd_top2", ds, 2, 2, "none", False, 2)
    case("test_rdd_top3_rev", ds, 2, 3, "none", True, 2)
    stable = [(rnd.randrange(1, 4), (rnd.randrange(1, 4), i)) for i in range(30)]
    case("stable_first_rev", stable, 2, 3, "first", True, 2)
    case("stable_first", stable, 2, 2, "first", False, 2)
    rows = [(rnd.randrange(12), rnd.randrange(-40, 40)) for _ in range(600)]
    for order in ("none", "mod7", "neg"):
        for reverse in (False, True):
            case("rand_%s_%s" % (order, "rev" if reverse else "fwd"), rows, 5, 4, order, reverse, 3)
    case("top_n_larger_than_groups", rows[:40], 3, 50, "none", False, 2)
    json.dump({"cases": cases}, open(os.path.join(HERE, "topbykey_cases.json"), "w"), separators=(",", ":"))
    dc.stop()
    print("wrote", len(cases), "topByKey cases")


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
