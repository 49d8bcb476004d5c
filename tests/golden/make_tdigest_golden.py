#!/usr/bin/env python
"""Capture golden vectors for dpark_b200/quantiles.py from the REFERENCE's t-digest
(/root/reference/dpark/utils/tdigest.py, loaded by path: it is pure Python and needs only `six`).

    python tests/golden/make_tdigest_golden.py      # writes tests/golden/tdigest_vectors.json

Each case = partitions of numbers -> one digest per partition (add, compress), merged left to right with
`+` exactly as RDD.percentiles does (dpark/rdd.py:791-814), queried at the listed percentiles.  Floats are
stored as hex strings (float.hex) so the comparison is bit-exact.  The `skew` cases also record the
thresholds combineByKey(fixSkew) would derive from them (dpark/rdd.py:516-537).
This script only runs in the build container (the reference is not on the GPU box); the JSON travels."""
import importlib.util
import json
import math
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_tdigest", "/root/reference/dpark/utils/tdigest.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def run(parts, percents):
    merged = None
    for part in parts:
        d = ref.TDigest()
        for x in part:
            d.add(x)
        d.compress()
        merged = d if merged is None else merged + d
    merged.compress()
    qs = [merged.quantile(p / 100.) for p in percents]
    cents = [(c.mean.hex(), c.count.hex()) for c in merged.centroids]
    return qs, cents, len(merged)


def thresholds(parts, splits):
    step = 100. / splits
    offs = [step * i for i in range(1, splits)]
    qs, _, _ = run(parts, offs)
    thr = []
    for p in qs:
        if math.isnan(p):
            continue
        p = int(math.ceil(p))
        if not thr or p > thr[-1]:
            thr.append(p)
    return thr


def split(seq, n):
    per = (len(seq) + n - 1) // n
    return [seq[i * per:(i + 1) * per] for i in range(n)]


def main():
    rnd = random.Random(20240611)
    cases = []
    pcs = [0, 0.1, 1, 5, 10, 25, 33.333, 50, 66.6, 75, 90, 99, 99.9, 100]

    def case(name, parts, percents=pcs):
        qs, cents, n = run(parts, percents)
        cases.append({"name": name, "parts": parts, "percents": percents, "quantiles": [q.hex() for q in qs],
                      "centroids": cents, "len": n})

    case("single_value", [[42]])
    case("two_values", [[1, 2]])
    case("small_ints", split(list(range(15)), 4))
    case("with_empty_partitions", [[], [3, 1, 2], [], [10, 9], []])
    case("first_partition_empty", [[], [5, 6, 7, 8]])
    case("duplicates", split([7] * 300 + [9] * 50 + [1] * 5, 3))
    case("sorted_5k", split(list(range(5000)), 5))
    case("reverse_5k", split(list(range(5000, 0, -1)), 5))
    case("uniform_6k", split([rnd.randrange(-2 ** 62, 2 ** 62) for _ in range(6000)], 8))
    case("gauss_floats_8k", split([rnd.gauss(0, 1e6) for _ in range(8000)], 3))
    case("zipf_like_6k", split([int(1 / (rnd.random() + 1e-4)) for _ in range(6000)], 7))
    case("ragged", [[rnd.randrange(1000) for _ in range(n)] for n in (1, 0, 250, 3, 1000, 17)])

    skew = []
    # the reference's own fixSkew test input (tests/test_rdd.py:259-266): keys 0..9 once, key 10 five times,
    # makeRDD(dsk, 10): hash(int k) == k
    dsk = list(range(10)) + [10] * 5

    slice_like_parallelize = split           # ParallelCollection.slice: ceil(m/n) rows per slice (dpark/rdd.py:1576-1596)

    for name, keys, nparts, splits in [
        ("test_rdd_fixSkew", dsk, 10, 3),
        ("hot_key", [5] * 4000 + [rnd.randrange(-10 ** 9, 10 ** 9) for _ in range(6000)], 6, 8),
        ("uniform_hashes", [rnd.randrange(-2 ** 61, 2 ** 61) for _ in range(8000)], 4, 16),
        ("all_equal", [3] * 100, 2, 4),
        ("tiny", [1, 2, 3], 3, 5),
    ]:
        parts = slice_like_parallelize(keys, nparts)
        skew.append({"name": name, "parts": parts, "splits": splits, "thresholds": thresholds(parts, splits)})

    json.dump({"cases": cases, "skew": skew}, open(os.path.join(HERE, "tdigest_vectors.json"), "w"),
              separators=(",", ":"))
    print("wrote %d digest cases, %d skew cases" % (len(cases), len(skew)))
    for s in skew:
        print(" ", s["name"], s["splits"], s["thresholds"][:6])


if __name__ == "__main__":
    main()
