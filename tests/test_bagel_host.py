"""Bagel (dpark_b200/bagel.py) against results captured from the reference's Bagel on the same jobs
(tests/golden/make_bagel_golden.py, tests/bagel_jobs.py).  CPU: the two shuffles of every superstep
(combineByKey of the messages, groupWith of vertices and messages) go through a stand-in engine built from the
oracle's hash/partition functions; the loop, the accumulators, the aggregator plumbing and the termination rule
are the code under test.  PageRank values are float sums whose order the reference does not fix: tolerance
1e-12 relative."""
import sys

import pytest

from tests import bagel_jobs
from tests.standin import standin_engine  # noqa: F401  (fixture)
from tests.golden_util import load

CASES = load("bagel_cases.json")["cases"]


def _ctx():
    sys.argv = [sys.argv[0]]
    from dpark_b200 import DparkContext
    return DparkContext("local")


@pytest.mark.parametrize("case", [c for c in CASES if c["job"] == "pagerank"], ids=lambda c: c["name"])
def test_pagerank_matches_the_reference(case, standin_engine):
    from dpark_b200 import bagel
    got = bagel_jobs.run_pagerank(_ctx(), bagel, case["graph"], case["parts"])
    want = {k: float.fromhex(v) for k, v in case["values"].items()}
    assert sorted(got) == sorted(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-12 * abs(want[k])
    assert abs(sum(got.values()) - sum(want.values())) < 1e-12


@pytest.mark.parametrize("case", [c for c in CASES if c["job"] == "maxprop"], ids=lambda c: c["name"])
def test_max_propagation_with_aggregator_matches_the_reference(case, standin_engine):
    from dpark_b200 import bagel
    got = bagel_jobs.run_maxprop(_ctx(), bagel, [[v, ts] for v, ts in case["graph"]], case["parts"])
    assert {str(k): v for k, v in got.items()} == case["values"]


def test_bagel_surface_and_termination(standin_engine):
    from dpark_b200 import bagel
    from dpark_b200.accumulator import Accumulator, listAcc
    dc = _ctx()
    calls = []

    def compute(vert, inbox):                      # two-argument form through addAggregatorArg
        calls.append(vert.id)
        return bagel.Vertex(vert.id, vert.value + 1, [], False), []

    verts = dc.parallelize([(i, bagel.Vertex(i, 0, [], True)) for i in range(4)], 2)
    out = bagel.Bagel.run(dc, verts, dc.parallelize([], 2), bagel.Bagel.addAggregatorArg(compute))
    assert sorted((k, v.value, v.active) for k, v in out.collect()) == [(i, 1, False) for i in range(4)]
    # maxSuperstep bounds the loop even when vertices stay active
    def forever(vert, inbox, agg, step):
        return bagel.Vertex(vert.id, step, [], True), []
    out = bagel.Bagel.run(dc, verts, dc.parallelize([], 2), forever, maxSuperstep=3)
    assert set(v.value for _, v in out.collect()) == {2}
    # combiners the GPU shuffle cannot express are refused when the job is declared
    class SetCombiner(bagel.Combiner):
        def createCombiner(self, msg):
            return {msg}

        def mergeValue(self, combiner, msg):
            return combiner | {msg}

        def mergeCombiners(self, a, b):
            return a | b

    with pytest.raises(NotImplementedError):
        bagel.Bagel.run(dc, verts, dc.parallelize([(0, 1)], 2), forever, combiner=SetCombiner(), maxSuperstep=1)
    acc = Accumulator([], listAcc)
    acc.add([1])
    acc.add([2])
    assert acc.value == [1, 2] and acc.reset() == [1, 2] and acc.value == []
    assert dc.accumulator(5).value == 5


def test_list_collecting_combiners_run_as_group_by(standin_engine):
    """An aggregator that builds lists by hand is recognised as a group-by (trace._builds_lists): Bagel's
    DefaultListCombiner and the usual three-lambda combineByKey give every key its values in (split, position)
    order; an aggregator that builds something else is still refused."""
    from dpark_b200 import bagel, trace
    from dpark_b200.dependency import Aggregator
    assert trace.recognize_aggregator(bagel.DefaultListCombiner()) == ("group", None)
    lists = Aggregator(lambda v: [v], lambda c, v: c + [v], lambda a, b: a + b)
    assert trace.recognize_aggregator(lists) == ("group", None)
    dc = _ctx()
    rows = [(i % 4, i) for i in range(20)]
    got = dict(dc.parallelize(rows, 3).combineByKey(lists, 2).collect())
    assert got == {k: [i for i in range(20) if i % 4 == k] for k in range(4)}
    for bad in (Aggregator(lambda v: [v, v], lambda c, v: c + [v], lambda a, b: a + b),      # not [v]
                Aggregator(lambda v: [v], lambda c, v: [v] + c, lambda a, b: a + b),          # prepends
                Aggregator(lambda v: {v}, lambda c, v: c | {v}, lambda a, b: a | b)):         # sets
        with pytest.raises(NotImplementedError):
            trace.recognize_aggregator(bad)

    def compute(vert, inbox, agg, step):          # inbox = [list of messages] with the list combiner
        seen = sorted(inbox[0]) if inbox else []
        out = [(e.target_id, vert.id) for e in vert.outEdges] if step == 0 else []
        return bagel.Vertex(vert.id, seen, vert.outEdges, step == 0), out

    verts = dc.parallelize([(i, bagel.Vertex(i, None, [bagel.Edge((i + 1) % 3), bagel.Edge((i + 2) % 3)], True))
                            for i in range(3)], 2)
    out = bagel.Bagel.run(dc, verts, dc.parallelize([], 2), compute, combiner=bagel.DefaultListCombiner())
    assert dict((k, v.value) for k, v in out.collect()) == {0: [1, 2], 1: [0, 2], 2: [0, 1]}


def test_compute_runs_once_per_vertex_and_superstep(standin_engine):
    """without caching a superstep's (vertex, outbox) rows every later superstep re-ran compute() for all
    earlier ones through the co-partitioned cogroup's narrow dependency (2000 calls instead of 200 for 20 vertices
    and 10 supersteps, plus 200 more on the final collect)."""
    from dpark_b200 import bagel
    dc = _ctx()
    calls = []

    def compute(vert, inbox, agg, step):
        calls.append((step, vert.id))
        return bagel.Vertex(vert.id, vert.value + 1, [], step < 9), [(vert.id, 1)]

    verts = dc.parallelize([(i, bagel.Vertex(i, 0, [], True)) for i in range(20)], 4)
    out = bagel.Bagel.run(dc, verts, dc.parallelize([], 4), compute, maxSuperstep=10)
    assert sorted(v.value for _, v in out.collect()) == [10] * 20
    assert len(calls) == 200 and len(set(calls)) == 200
