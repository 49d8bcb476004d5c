"""Runs one ShuffledRDD on the GPU: the static two-stage plan that replaces the
reference's DAG scheduler for this path (out of scope as a
system; the map stage and the reduce stage are simply run in order).

    stage 1 (map)    every parent split -> columns -> HBM -> dpark_b200.shuffle.map_side
    stage 2 (reduce) exchange + reduce_side / group_side -> per-partition columns

Python rows exist only before stage 1 (ingest of what user lambdas produced) and
after stage 2 (egress to user lambdas); see dpark_b200.columnar.
"""
import numpy as np
import torch

from . import _native as nv
from . import columnar, shuffle


class ShuffleResult(object):
    """Per-partition result columns on the host + lazy conversion to rows."""

    def __init__(self, nparts):
        self.parts = [None] * nparts      # (keys: list, vals: list | (offsets, values)) per partition

    def rows(self, p):
        keys, vals = self.parts[p]
        return list(zip(keys, vals))

    def columns(self, p):
        return self.parts[p]


def _device():
    if not torch.cuda.is_available():
        raise nv.NativeError("the dpark_b200 shuffle needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _gather_parent(srdd, numeric_values):
    """Stage-1 input: one Columns (host) or a (keys, vals) tensor pair per parent split."""
    from .rdd import ColumnarRDD
    parent = srdd.parent
    out = []
    for sp in parent.splits:
        if isinstance(parent, ColumnarRDD) and numeric_values:
            out.append(parent.columns(sp))
        else:
            out.append(columnar.ingest_pairs(parent.iterator(sp), repr(parent), numeric_values))
    return out


def run_shuffle(srdd):
    dev = _device()
    P = srdd.partitioner.numPartitions
    thr = srdd.partitioner.thresholds
    if srdd.kind == "reduce":
        splits = _gather_parent(srdd, True)
        return _run_reduce(splits, P, thr, srdd.op, dev)
    splits = _gather_parent(srdd, False)
    return _run_group(splits, P, thr, dev)


def _key_kind_of(splits):
    kinds = set(c.key_kind for c in splits if isinstance(c, columnar.Columns) and c.n)
    if len(kinds) > 1:
        raise TypeError("mixed key types %s in one shuffle are not supported on the B200 path" % sorted(kinds))
    return kinds.pop() if kinds else columnar.KEY_I64


def _run_reduce(splits, P, thr, op, dev):
    res = ShuffleResult(P)
    tensor_in = splits and not isinstance(splits[0], columnar.Columns)
    if tensor_in:
        kc = [k.to(dev).contiguous() for k, v in splits]
        vc = [v.to(dev).contiguous() for k, v in splits]
        parts = shuffle.reduce_by_key(kc, vc, P, op, thr)
        for p, k, v in parts:
            res.parts[p] = (k.cpu().numpy().tolist(), v.cpu().numpy().tolist())
        return res
    kk = _key_kind_of(splits)
    vkinds = set(c.val_kind for c in splits if c.n)
    if len(vkinds) > 1:
        raise TypeError("reduceByKey values must be all int or all float on the B200 path")
    if kk in (columnar.KEY_I64, columnar.KEY_F64):
        kdt = np.int64 if kk == columnar.KEY_I64 else np.float64
        kc = [torch.from_numpy(c.keys.astype(kdt, copy=False)).to(dev) for c in splits]
        vc = [torch.from_numpy(c.vals).to(dev) for c in splits]
        if vkinds:
            vdt = torch.int64 if vkinds == {columnar.VAL_I64} else torch.float64
            vc = [v.to(vdt) for v in vc]
        if vkinds == {columnar.VAL_I64} and op == "sum":
            # the reference adds Python big ints; the device accumulates in int64.  A cheap sufficient check: if the
            # sum of |v| over the whole shuffle stays below 2^63 no key's sum can wrap
            bound = sum(float(np.abs(c.vals.astype(np.float64)).sum()) for c in splits if c.n)
            if bound >= 2.0 ** 63:
                raise OverflowError("reduceByKey(add): the values' magnitudes sum to %.3g >= 2^63; int64 accumulation on the "
                                    "B200 path could wrap where the reference's big ints do not" % bound)
        parts = shuffle.reduce_by_key(kc, vc, P, op, thr)
        for p, k, v in parts:
            res.parts[p] = (k.cpu().numpy().tolist(), v.cpu().numpy().tolist())
        return res
    from . import strings
    return strings.reduce_by_key_bytes(splits, kk, P, thr, op, dev, res)


def _run_group(splits, P, thr, dev):
    from . import grouping
    return grouping.group_by_key(splits, P, thr, dev, ShuffleResult(P))
