"""cogroup / join host logic (tagging, stable split by tag, outer-join None padding, fixSkew over the union)
against the reference's golden outputs, on CPU: the GPU group-by engine is replaced by a stand-in built from
the oracle's hash/partition functions that honours the same contract (per partition: keys with their values in
(map split, position) order).  The same cases run through the real engine in tests/test_gpu_rdd.py."""
import pytest

from tests import cogroup_common as cc
from tests.standin import standin_engine  # noqa: F401  (fixture)


@pytest.mark.parametrize("case", cc.COGROUP_CASES, ids=[c["name"] for c in cc.COGROUP_CASES])
def test_cogroup_matches_reference(case, standin_engine):
    cc.check_cogroup(case)


@pytest.mark.parametrize("case", cc.JOIN_CASES, ids=[c["name"] for c in cc.JOIN_CASES])
def test_joins_match_reference(case, standin_engine):
    cc.check_join(case)


def test_group_with_accepts_one_rdd_or_a_list_and_defaults(standin_engine):
    dc = cc.ctx()
    a = dc.parallelize([(1, "a"), (2, "b")], 2)
    b = dc.parallelize([(2, "x")], 1)
    one = dict(a.groupWith(b).collect())
    many = dict(a.cogroup([b]).collect())
    assert one == many == {1: (["a"], []), 2: (["b"], ["x"])}
    assert len(a.groupWith(b).splits) == dc.defaultParallelism
    assert dict(a.join(b).collect()) == {2: ("b", "x")}
    assert sorted(a.leftOuterJoin(b).collect()) == [(1, ("a", None)), (2, ("b", "x"))]


def test_co_partitioned_inputs_are_not_shuffled_again(standin_engine):
    """dpark/rdd.py:1280-1293: inputs that already carry the partitioner are read through narrow dependencies.
    When every input does, the cogroup is a per-partition merge with no shuffle at all (Bagel's steady state)."""
    import random
    dc = cc.ctx()
    rnd = random.Random(9)
    ra = [(rnd.randrange(40), rnd.randrange(100)) for _ in range(300)]
    rb = [(rnd.randrange(20, 60), rnd.randrange(100)) for _ in range(200)]
    a = dc.parallelize(ra, 4).groupByKey(3).flatMapValue(lambda vs: vs)
    b = dc.parallelize(rb, 2).reduceByKey(lambda x, y: x + y, 3)
    out = a.groupWith(b)                                  # numSplits defaults to a's partitioner
    assert out.narrow == [0, 1] and out._grouped is None and len(out.splits) == 3
    plain = dc.parallelize(ra, 4).groupWith(dc.parallelize(rb, 2).reduceByKey(lambda x, y: x + y, 3), 3)
    assert plain.narrow == [1] and plain._grouped is not None
    canon = lambda rdd: [sorted((k, sorted(g0), sorted(g1)) for k, (g0, g1) in part) for part in rdd.glom().collect()]
    assert canon(out) == canon(plain)
    want = {}
    for k, v in ra:
        want.setdefault(k, ([], []))[0].append(v)
    sums = {}
    for k, v in rb:
        sums[k] = sums.get(k, 0) + v
    for k, v in sums.items():
        want.setdefault(k, ([], []))[1].append(v)
    got = dict((k, (sorted(g0), sorted(g1))) for k, (g0, g1) in out.collect())
    assert got == dict((k, (sorted(g0), sorted(g1))) for k, (g0, g1) in want.items())


def test_uniq_top_hot(standin_engine):
    """dpark/rdd.py:383-398 over the shuffle: uniq keeps one of each element in the partition its hash selects;
    hot = counts + top."""
    import collections
    import random
    from oracle import oracle as orc
    dc = cc.ctx()
    rnd = random.Random(1)
    xs = [rnd.randrange(-30, 30) for _ in range(500)] + ["w%d" % rnd.randrange(9) for _ in range(0)]
    parts = dc.parallelize(xs, 5).uniq(4).glom().collect()
    assert sorted(x for p in parts for x in p) == sorted(set(xs))
    assert all(orc.get_partition(x, 4) == i for i, p in enumerate(parts) for x in p)
    words = ["w%d" % int(rnd.paretovariate(1.2)) for _ in range(800)]
    hot = dc.parallelize(words, 3).hot(5, 2)
    want = collections.Counter(words).most_common()
    assert [c for _, c in hot] == [c for _, c in want[:5]]
    assert all(dict(want)[w] == c for w, c in hot)
    nums = dc.parallelize(list(range(100)), 7)
    assert nums.top(3) == [99, 98, 97] and nums.top(2, reverse=True) == [0, 1]
    assert nums.top(2, key=lambda x: -abs(x - 50)) == [50, 49]


TOPK = __import__("tests.golden_util", fromlist=["load"]).load("topbykey_cases.json")["cases"]
TOPK_ORDER = {"none": None, "first": lambda x: x[0], "mod7": lambda x: x % 7, "neg": lambda x: -x}


@pytest.mark.parametrize("case", TOPK, ids=[c["name"] for c in TOPK])
def test_top_by_key_matches_reference(case, standin_engine):
    """topByKey against the reference's outputs (its own test inputs, tests/test_rdd.py:353-374, and seeded rows
    with many ties): bounded-heap semantics = stable sort of the (partition, position)-ordered group, cut."""
    import json
    from tests.golden.make_golden import enc
    from tests.golden_util import dec
    dc = cc.ctx()
    rows = [(dec(k), dec(v)) for k, v in case["rows"]]
    out = dc.makeRDD(rows, case["M"]).topByKey(case["top_n"], TOPK_ORDER[case["order"]], case["reverse"], case["P"])
    got = [sorted(([enc(k), enc(list(v))] for k, v in part), key=json.dumps) for part in out.glom().collect()]
    assert got == case["parts"]
