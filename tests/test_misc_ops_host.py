"""groupBy / update / innerJoin / percentilesByKey / fold / aggregate / enumerate against outputs captured from the
reference (tests/golden/make_misc_golden.py).  CPU: shuffles go through the stand-in engine (tests/standin.py);
the operator compositions are the code under test."""
import json

import pytest

from tests import cogroup_common as cc
from tests.golden.make_golden import enc
from tests.golden_util import dec, load
from tests.standin import standin_engine  # noqa: F401  (fixture)

G = load("misc_cases.json")


def _canon(parts, val=enc):
    return [sorted(([enc(k), val(v)] for k, v in part), key=json.dumps) for part in parts]


def test_group_by(standin_engine):
    c = G["groupBy_mod5"]
    got = cc.ctx().parallelize(c["xs"], c["M"]).groupBy(lambda x: x % 5, c["P"]).glom().collect()
    assert _canon(got, lambda v: enc(list(v))) == c["parts"]


@pytest.mark.parametrize("name", ["update_all", "update_replace_only"])
def test_update(name, standin_engine):
    c = G[name]
    dc = cc.ctx()
    old, new = [tuple(r) for r in c["old"]], [tuple(r) for r in c["new"]]
    got = dc.parallelize(old, 3).update(dc.parallelize(new, 2), replace_only=c["replace_only"], numSplits=c["P"])
    assert _canon(got.glom().collect()) == c["parts"]


def test_update_reference_test_vectors(standin_engine):
    c = G["update_test_rdd"]                       # tests/test_rdd.py:376-381, None as a new value included
    dc = cc.ctx()
    old = [(dec(k), dec(v)) for k, v in c["old"]]
    new = [(dec(k), dec(v)) for k, v in c["new"]]
    for key, ro in (("all", False), ("replace_only", True)):
        got = dc.makeRDD(old).update(dc.makeRDD(new), replace_only=ro).collect()
        assert sorted(json.dumps([enc(k), enc(v)]) for k, v in got) == c[key]


def test_inner_join(standin_engine):
    c = G["innerJoin"]
    dc = cc.ctx()
    big = [tuple(r) for r in c["big"]]
    small = [(k, dec(v)) for k, v in c["small"]]
    got = dc.parallelize(big, c["M"]).innerJoin(dc.parallelize(small, 2)).collect()
    assert sorted(json.dumps([k, enc(tuple(v))]) for k, v in got) == c["rows"]


def test_percentiles_by_key_single_map_is_bit_identical(standin_engine):
    c = G["percentilesByKey_one_map"]              # one digest per key: nothing depends on a merge order
    rows = [(k, float.fromhex(v)) for k, v in c["rows"]]
    got = cc.ctx().parallelize(rows, c["M"]).percentilesByKey(c["p"], numSplits=c["P"]).glom().collect()
    assert [sorted([k, [q.hex() for q in qs]] for k, qs in part) for part in got] == c["parts"]


def test_percentiles_by_key_across_map_tasks_tracks_the_order_statistics(standin_engine):
    """With several map tasks the reference merges per-task digests with `d1 + d2`, which ASSIGNS the right
    operand's weight to `_unmerge_weight` (dpark/utils/tdigest.py:69) and so forgets the left operand's still
    buffered points: its answers are biased low (by 10-200 on N(0, 100) data in the captured run, see the golden
    file) and depend on the random fetch order.  This implementation compresses before merging; it is checked
    against the exact order statistics instead, and only the key -> partition layout against the reference."""
    import numpy as np
    c = G["percentilesByKey_four_maps"]
    rows = [(k, float.fromhex(v)) for k, v in c["rows"]]
    got = cc.ctx().parallelize(rows, c["M"]).percentilesByKey(c["p"], numSplits=c["P"]).glom().collect()
    assert [sorted(k for k, _ in part) for part in got] == [[k for k, _ in part] for part in c["parts"]]
    ref_err = ours_err = 0.0
    for gp, wp in zip(got, c["parts"]):
        want = dict((k, [float.fromhex(q) for q in qs]) for k, qs in wp)
        for k, qs in gp:
            vals = np.sort([v for kk, v in rows if kk == k])
            for p, q, w in zip(c["p"], qs, want[k]):
                rank = np.searchsorted(vals, q) / len(vals)
                assert abs(rank - p / 100.) <= 0.02              # within 2 % of the requested rank
                exact = np.percentile(vals, p)
                ours_err = max(ours_err, abs(q - exact))
                ref_err = max(ref_err, abs(w - exact))
    assert ours_err < 30 < ref_err                                # and far closer than the reference's own answer


def test_fold_aggregate_enumerate():
    dc = cc.ctx()
    r = dc.parallelize(list(range(20)), 3)
    assert r.fold(0, lambda a, b: a + b) == 190
    assert r.aggregate([], lambda acc, x: acc + [x * x], lambda a, b: a + b) == [x * x for x in range(20)]
    assert r.toList() == list(range(20))
    assert dc.parallelize(list("abcd"), 3).enumerate().collect() == [(0, "a"), (1, "b"), (2, "c"), (3, "d")]
    assert [i for i, _ in dc.parallelize(list("abcd"), 2).enumeratePartition().collect()] == [0, 0, 1, 1]
    seen = []
    r.foreachPartition(lambda it: seen.append(sum(it)))
    assert sum(seen) == 190


SORT = load("sort_cases.json")["cases"]
SORT_KEYS = {"id": lambda x: x, "neg": lambda x: -x, "second": lambda x: x[1], "mod": lambda x: (x % 10, x)}


@pytest.mark.parametrize("case", SORT, ids=[c["name"] for c in SORT])
def test_sort_partitions_equal_the_reference(case, standin_engine):
    """RDD.sort: same sample-derived range bounds, hence the same partitions with the same rows in the same order
    (equal keys may be ordered differently inside a partition: the reference does not fix the order in which a
    reducer meets them, so rows are compared after a stable re-sort on (key, repr))."""
    xs = [dec(x) for x in case["xs"]]
    xs = [tuple(x) if isinstance(x, list) else x for x in xs]
    key = SORT_KEYS[case["key"]]
    got = cc.ctx().parallelize(xs, case["M"]).sort(key=key, reverse=case["reverse"], numSplits=case["P"]).glom().collect()
    want = [[dec(x) for x in part] for part in case["parts"]]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert [key(x) for x in g] == [key(x) for x in w]            # same keys in the same positions
        assert sorted(map(repr, g)) == sorted(map(repr, w))          # same rows
    flat = [key(x) for part in got for x in part]
    assert flat == sorted(flat, reverse=case["reverse"])             # globally sorted across partitions
