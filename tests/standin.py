"""A stand-in for dpark_b200.engine.run_shuffle so that HOST-side operator logic (cogroup tagging, joins, Bagel's
loop, fixSkew thresholds) can be checked on CPU against the reference's golden outputs.  It honours the engine's
contract -- per reduce partition: the keys HashPartitioner assigns to it, with the values combined by the
aggregator (reduce) or listed in (map split, position) order (group) -- using the oracle's hash / partition
functions.  Test infrastructure only; the product has no CPU shuffle."""
import pytest

from oracle import oracle as orc


def run_shuffle(srdd):
    from dpark_b200 import engine
    P, thr = srdd.partitioner.numPartitions, srdd.partitioner.thresholds
    agg = srdd.aggregator
    buckets = [dict() for _ in range(P)]
    for sp in srdd.parent.splits:
        for k, v in srdd.parent.iterator(sp):
            b = buckets[orc.get_partition(k, P, thr)]
            if srdd.kind == "group":
                b.setdefault(k, []).append(v)
            else:
                b[k] = agg.mergeValue(b[k], v) if k in b else agg.createCombiner(v)
    res = engine.ShuffleResult(P)
    for p, b in enumerate(buckets):
        res.parts[p] = (list(b.keys()), list(b.values()))
    return res


@pytest.fixture
def standin_engine(monkeypatch):
    from dpark_b200 import columnar, engine
    monkeypatch.setattr(engine, "run_shuffle", run_shuffle)
    monkeypatch.setattr(columnar, "hashes_of_keys", lambda keys: [orc.portable_hash(k) for k in keys])
