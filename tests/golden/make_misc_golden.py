#!/usr/bin/env python
"""Golden vectors FROM THE REAL REFERENCE for the smaller shuffle-backed operators: groupBy, update,
innerJoin, percentilesByKey (dpark/rdd.py:298-301, 599-647, 815-850); same out-of-tree build as make_golden.py.

    python tests/golden/make_misc_golden.py         # writes tests/golden/misc_cases.json

`update` inputs have at most one row per key and side (with more the reference's fold depends on the random
fetch order); percentilesByKey runs with ONE map task for the bit-exact cases (a single digest per key) and with
several map tasks for the tolerance cases (merge order is not fixed by the reference)."""
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


def generate():
    import dpark.conf
    from dpark import DparkContext
    logging.getLogger("dpark").setLevel(logging.ERROR)
    dc = DparkContext("local")
    dc.init()
    rnd = random.Random(23)
    conf = dpark.conf.rddconf(ordered_group=True)
    out = {}

    def canon(parts, val=lambda v: enc(v)):
        return [sorted(([enc(k), val(v)] for k, v in part), key=json.dumps) for part in parts]

    xs = [rnd.randrange(-100, 100) for _ in range(400)]
    out["groupBy_mod5"] = {"xs": xs, "M": 4, "P": 3,
                           "parts": canon(dc.parallelize(xs, 4).groupBy(lambda x: x % 5, 3, rddconf=conf).glom().collect(),
                                          lambda v: enc(list(v)))}
    old = [(k, rnd.randrange(1000)) for k in rnd.sample(range(200), 80)]
    new = [(k, -rnd.randrange(1000)) for k in rnd.sample(range(200), 60)]
    for replace_only in (False, True):
        r = dc.parallelize(old, 3).update(dc.parallelize(new, 2), replace_only=replace_only, numSplits=4)
        out["update_%s" % ("replace_only" if replace_only else "all")] = {
            "old": old, "new": new, "replace_only": replace_only, "P": 4, "parts": canon(r.glom().collect())}
    # the reference's own test (tests/test_rdd.py:376-381)
    r4 = [("foo", 1), ("wtf", 233)]
    r5 = [("foo", 2), ("bar", 3), ("wtf", None)]
    out["update_test_rdd"] = {"old": [[enc(k), enc(v)] for k, v in r4], "new": [[enc(k), enc(v)] for k, v in r5],
                              "all": sorted(json.dumps([enc(k), enc(v)]) for k, v in
                                            dc.makeRDD(r4).update(dc.makeRDD(r5)).collect()),
                              "replace_only": sorted(json.dumps([enc(k), enc(v)]) for k, v in
                                                     dc.makeRDD(r4).update(dc.makeRDD(r5), replace_only=True).collect())}
    big = [(rnd.randrange(30), rnd.randrange(100)) for _ in range(300)]
    small = [(rnd.randrange(10, 40), "s%d" % i) for i in range(25)]
    out["innerJoin"] = {"big": big, "small": [[k, enc(v)] for k, v in small], "M": 3,
                        "rows": sorted(json.dumps([k, enc(tuple(v))]) for k, v in
                                       dc.parallelize(big, 3).innerJoin(dc.parallelize(small, 2)).collect())}
    rows = [(rnd.randrange(6), rnd.gauss(0, 100)) for _ in range(3000)]
    pcs = [10, 50, 90, 99]
    for name, M in (("percentilesByKey_one_map", 1), ("percentilesByKey_four_maps", 4)):
        r = dc.parallelize(rows, M).percentilesByKey(pcs, numSplits=3).glom().collect()
        out[name] = {"rows": [[k, v.hex()] for k, v in rows], "M": M, "P": 3, "p": pcs,
                     "parts": [sorted([k, [q.hex() for q in qs]] for k, qs in part) for part in r]}
    json.dump(out, open(os.path.join(HERE, "misc_cases.json"), "w"), separators=(",", ":"))
    dc.stop()
    print("wrote", len(out), "cases")


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
