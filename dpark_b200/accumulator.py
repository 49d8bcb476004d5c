"""Accumulators (dpark/accumulator.py): write-only counters a job adds to and the driver reads.

The reference ships per-task deltas from executors back to the driver; the B200 path runs every Python
stage in the driver process, so an accumulator is simply a value with the reference's `add` / `value` /
`reset` surface and its four stock parameter sets.  Under torch.distributed (one driver process per GPU, all running
the same script: dpark_b200/spmd.py) each rank also keeps what it added since the last job, and the ranks exchange
those deltas after every job, so `value` is the job-wide total on every rank."""
import copy
from operator import add


class AccumulatorParam(object):
    def __init__(self, zero, addInPlace):
        self.zero, self.addInPlace = zero, addInPlace


numAcc = AccumulatorParam(0, add)
listAcc = AccumulatorParam([], lambda x, y: x.extend(y) or x)
mapAcc = AccumulatorParam({}, lambda x, y: x.update(y) or x)
setAcc = AccumulatorParam(set(), lambda x, y: x.update(y) or x)


class Accumulator(object):
    _next_id = 0

    def __init__(self, initialValue=0, param=numAcc):
        Accumulator._next_id += 1
        self.id = Accumulator._next_id
        self.param = numAcc if param is None else param
        self.value = initialValue
        self._delta = copy.copy(self.param.zero)
        from . import spmd
        spmd.register_accumulator(self)

    def add(self, v):
        self.value = self.param.addInPlace(self.value, v)
        self._delta = self.param.addInPlace(self._delta, copy.copy(v) if isinstance(v, (list, dict, set)) else v)

    def _take_delta(self):
        d, self._delta = self._delta, copy.copy(self.param.zero)
        return d

    def _absorb(self, d):
        self.value = self.param.addInPlace(self.value, d)

    def reset(self):
        v, self.value = self.value, copy.copy(self.param.zero)
        self._delta = copy.copy(self.param.zero)
        return v

    def __repr__(self):
        return "<Accumulator %d: %r>" % (self.id, self.value)
