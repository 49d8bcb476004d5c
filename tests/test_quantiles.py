"""dpark_b200/quantiles.py (merging t-digest + fixSkew thresholds) against vectors captured from the
reference's own class (tests/golden/make_tdigest_golden.py -> tdigest_vectors.json).  Bit-exact:
floats are compared through float.hex().  CPU-only."""
import json
import math
import os

import pytest

from dpark_b200 import quantiles

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "tdigest_vectors.json")) as f:
    GOLD = json.load(f)


def _merged(parts):
    merged = None
    for part in parts:
        d = quantiles.MergingDigest().update(part)
        d.compress()
        merged = d if merged is None else merged + d
    merged.compress()
    return merged


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_digest_is_bit_identical_to_the_reference(case):
    d = _merged(case["parts"])
    assert len(d) == case["len"]
    assert [(m.hex(), w.hex()) for m, w in zip(d.means, d.weights)] == [tuple(c) for c in case["centroids"]]
    got = [d.quantile(p / 100.) for p in case["percents"]]
    assert [g.hex() for g in got] == case["quantiles"]
    assert [g.hex() for g in quantiles.percentiles_of_partitions(case["parts"], case["percents"])] == case["quantiles"]


@pytest.mark.parametrize("case", GOLD["skew"], ids=[c["name"] for c in GOLD["skew"]])
def test_fix_skew_thresholds_match_the_reference(case):
    thr, splits = quantiles.skew_thresholds(case["parts"], case["splits"])
    assert thr == case["thresholds"]
    assert splits == len(case["thresholds"]) + 1
    assert all(a < b for a, b in zip(thr, thr[1:]))


def test_digest_edge_cases():
    d = quantiles.MergingDigest()
    assert math.isnan(d.quantile(0.5)) and len(d) == 0
    with pytest.raises(ValueError):
        d.add(float("nan"))
    with pytest.raises(ValueError):
        d.quantile(1.5)
    with pytest.raises(TypeError):
        d + 3
    d.add(5)
    assert d.quantile(0.0) == 5.0 and d.quantile(1.0) == 5.0
    assert quantiles.percentiles_of_partitions([], [50]) != quantiles.percentiles_of_partitions([], [50])  # NaN
    assert quantiles.skew_thresholds([[], []], 4) == ([], 1)        # nothing sampled: a single partition
