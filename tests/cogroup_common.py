"""Shared by the CPU (stand-in engine) and GPU runs of the cogroup / join golden cases
(tests/golden/cogroup_cases.json, captured from the reference by make_cogroup_golden.py)."""
import json
import sys

from tests.golden_util import dec, load

CG = load("cogroup_cases.json")
COGROUP_CASES = [c for c in CG["cases"] if c["op"] in ("cogroup", "cogroup_prepartitioned")]
JOIN_CASES = [c for c in CG["cases"] if c["op"] in ("join", "leftOuterJoin", "rightOuterJoin", "outerJoin")]


def ctx():
    sys.argv = [sys.argv[0]]
    from dpark_b200 import DparkContext
    return DparkContext("local")


def _enc(o):
    from tests.golden.make_golden import enc
    return enc(o)


def inputs_of(dc, case):
    return [dc.parallelize([(dec(k), dec(v)) for k, v in inp["rows"]], inp["M"]) for inp in case["inputs"]]


def check_cogroup(case):
    dc = ctx()
    rdds = inputs_of(dc, case)
    if case["op"] == "cogroup_prepartitioned":      # as the generator: the left input already has the partitioner
        from dpark_b200 import HashPartitioner
        rdds[0] = rdds[0].groupByKey(case["P"]).flatMapValue(lambda x: x)
        assert rdds[0].partitioner == HashPartitioner(case["P"])
    out = rdds[0].groupWith(rdds[1:], case["P"], fixSkew=case.get("fixSkew", -1))
    if case["op"] == "cogroup_prepartitioned":
        assert out.narrow == [0]
    if "thresholds" in case:
        assert out.partitioner.thresholds == case["thresholds"]
    parts = out.glom().collect()
    ordered = case["op"] == "cogroup"          # the pre-partitioned case is stored as multisets (see the generator)
    got = [sorted(([_enc(k), [_enc(list(g) if ordered else sorted(g)) for g in groups]] for k, groups in part),
                  key=json.dumps) for part in parts]
    assert got == case["parts"]
    assert all(isinstance(groups, tuple) and len(groups) == len(rdds) for part in parts for _, groups in part)


def check_join(case):
    dc = ctx()
    a, b = inputs_of(dc, case)
    parts = getattr(a, case["op"])(b, case["P"]).glom().collect()
    got = [sorted(([_enc(k), _enc(tuple(v))] for k, v in part), key=json.dumps) for part in parts]
    assert got == case["parts"]
