"""cogroup / join host logic (tagging, stable split by tag, outer-join None padding, fixSkew over the union)
against the reference's golden outputs, on CPU: the GPU group-by engine is replaced by a stand-in built from
the oracle's hash/partition functions that honours the same contract (per partition: keys with their values in
(map split, position) order).  The same cases run through the real engine in tests/test_gpu_rdd.py."""
import pytest

from oracle import oracle as orc
from tests import cogroup_common as cc


@pytest.fixture
def standin_engine(monkeypatch):
    from dpark_b200 import columnar, engine

    def run_shuffle(srdd):
        assert srdd.kind == "group"
        P, thr = srdd.partitioner.numPartitions, srdd.partitioner.thresholds
        buckets = [dict() for _ in range(P)]
        for sp in srdd.parent.splits:
            for k, v in srdd.parent.iterator(sp):
                buckets[orc.get_partition(k, P, thr)].setdefault(k, []).append(v)
        res = engine.ShuffleResult(P)
        for p, b in enumerate(buckets):
            res.parts[p] = (list(b.keys()), list(b.values()))
        return res

    monkeypatch.setattr(engine, "run_shuffle", run_shuffle)
    monkeypatch.setattr(columnar, "hashes_of_keys", lambda keys: [orc.portable_hash(k) for k in keys])


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
