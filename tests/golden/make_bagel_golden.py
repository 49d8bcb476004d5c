#!/usr/bin/env python
"""Golden vectors for Bagel FROM THE REAL REFERENCE (same out-of-tree build as make_golden.py).

    python tests/golden/make_bagel_golden.py        # writes tests/golden/bagel_cases.json

PageRank in the shape of examples/pagerank.py (vertex ids are strings, messages are floats combined with the
default BasicCombiner(operator.add), termination by epsilon after >= 10 supersteps) on seeded random graphs, plus a
max-propagation job (int ids, BasicCombiner(max), an Aggregator over the vertices).  The compute functions live in
tests/bagel_jobs.py and are shared with the tests; only the RESULTS of running them on the reference are stored."""
import json
import logging
import os
import random
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from make_golden import bootstrap, build_reference  # noqa: E402
import bagel_jobs  # noqa: E402


def generate():
    from dpark import DparkContext
    from dpark import bagel as ref_bagel
    logging.getLogger("dpark").setLevel(logging.ERROR)
    dc = DparkContext("local")
    dc.init()
    rnd = random.Random(5)
    cases = []
    for name, nv, deg, parts in [("tiny_5", 5, 2, 2), ("ring_12", 12, 1, 3), ("random_40", 40, 3, 4)]:
        graph = bagel_jobs.random_graph(rnd, nv, deg, ring=name.startswith("ring"))
        res = bagel_jobs.run_pagerank(dc, ref_bagel, graph, parts)
        cases.append({"job": "pagerank", "name": name, "graph": graph, "parts": parts,
                      "values": {k: v.hex() for k, v in sorted(res.items())}})
    for name, nv, deg, parts in [("maxprop_30", 30, 2, 3)]:
        graph = bagel_jobs.random_graph(rnd, nv, deg, int_ids=True)
        res = bagel_jobs.run_maxprop(dc, ref_bagel, graph, parts)
        cases.append({"job": "maxprop", "name": name, "graph": graph, "parts": parts,
                      "values": {str(k): v for k, v in sorted(res.items())}})
    json.dump({"cases": cases}, open(os.path.join(HERE, "bagel_cases.json"), "w"), separators=(",", ":"))
    dc.stop()
    print("wrote", len(cases), "bagel cases")


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
