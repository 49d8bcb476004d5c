"""Bagel -- the Pregel-style loop of dpark/bagel.py over this package's shuffles.

Every superstep is two shuffles on the GPU: the messages are combined per target vertex
(`msgs.combineByKey(combiner)` -- a reduceByKey when the combiner is a recognised binary op, e.g. the
default `BasicCombiner(operator.add)`), and the vertices are co-grouped with the combined messages
(`verts.groupWith(...)`, dpark_b200.rdd.CoGroupedRDD).  The user's `compute(vertex, message, aggregated,
superstep) -> (vertex', [(target_id, value), ...])` stays a host-side Python function, as in the reference.
Same classes, same argument meaning, same termination rule (no messages and no active vertices) as
dpark/bagel.py:83-128.
"""
import operator
import sys


class Vertex(object):
    def __init__(self, id_, value, outEdges, active):
        self.id, self.value, self.outEdges, self.active = id_, value, outEdges, active

    def __repr__(self):
        return "<Vertex(%s, %s, %s)>" % (self.id, self.value, self.active)


class Edge(object):
    def __init__(self, target_id, value=0):
        self.target_id, self.value = target_id, value

    def __repr__(self):
        return "<Edge(%s, %s)>" % (self.target_id, self.value)


class Message(object):
    def __init__(self, target_id, value):
        self.target_id, self.value = target_id, value

    def __repr__(self):
        return "<Message(%s, %s)>" % (self.target_id, self.value)


class Combiner(object):
    """How the messages bound for one vertex are merged (an Aggregator in dpark.dependency terms)."""

    def createCombiner(self, msg):
        raise NotImplementedError

    def mergeValue(self, combiner, msg):
        raise NotImplementedError

    def mergeCombiners(self, a, b):
        raise NotImplementedError


class Aggregator(object):
    """A global reduction over the vertices, handed to every compute() of the superstep."""

    def createAggregator(self, vert):
        raise NotImplementedError

    def mergeAggregator(self, a, b):
        raise NotImplementedError


class BasicCombiner(Combiner):
    def __init__(self, op):
        self.op = op

    def createCombiner(self, msg):
        return msg

    def mergeValue(self, combiner, msg):
        return self.op(combiner, msg)

    def mergeCombiners(self, a, b):
        return self.op(a, b)


DefaultValueCombiner = BasicCombiner(operator.add)


class DefaultListCombiner(Combiner):
    def createCombiner(self, msg):
        return [msg]

    def mergeValue(self, combiner, msg):
        return combiner + [msg]

    def mergeCombiners(self, a, b):
        return a + b


class Bagel(object):
    @classmethod
    def run(cls, ctx, verts, msgs, compute, combiner=DefaultValueCombiner, aggregator=None,
            maxSuperstep=sys.maxsize, numSplits=None, checkpointDir=None):
        # checkpointDir is accepted for signature compatibility and ignored: nothing is written to disk, every
        # superstep's (vertex, outbox) rows are cached in memory for exactly one superstep (see comp()).
        superstep = 0
        previous = None
        while superstep < maxSuperstep:
            aggregated = cls.agg(verts, aggregator) if aggregator else None
            inbox = msgs.combineByKey(combiner, numSplits)
            grouped = verts.groupWith(inbox, numSplits=numSplits)

            def step(vert, inbox_values, _agg=aggregated, _n=superstep):
                return compute(vert, inbox_values, _agg, _n)

            verts, msgs, sent, active, moved = cls.comp(ctx, grouped, step, checkpointDir)
            if previous is not None:
                previous.uncache()            # the superstep before last is no longer reachable through a cache miss
            previous = moved
            superstep += 1
            if sent == 0 and active == 0:
                break
        return verts

    @classmethod
    def agg(cls, verts, aggregator):
        # the reference calls `mergeAggregators` here although its base class spells it `mergeAggregator`
        # (dpark/bagel.py:48, 109): accept either spelling
        merge = getattr(aggregator, "mergeAggregators", None) or aggregator.mergeAggregator
        return verts.map(lambda kv: aggregator.createAggregator(kv[1])).reduce(merge)

    @classmethod
    def comp(cls, ctx, grouped, compute, checkpointDir=None):
        sent, active = ctx.accumulator(0), ctx.accumulator(0)

        def advance(groups):
            mine, inbox = groups
            if not mine:                      # messages for a vertex that does not exist are dropped
                return []
            vert, outbox = compute(mine[0], inbox)
            sent.add(len(outbox))
            if vert.active:
                active.add(1)
            return [(vert, outbox)]

        # cached: `moved` is read three times (the count below, the next superstep's message shuffle and its
        # groupWith), and through the narrow dependency of a co-partitioned cogroup every later superstep would
        # otherwise re-run compute() for all earlier ones (quadratic, and user side effects would re-fire)
        moved = grouped.flatMapValue(advance).cache()
        verts = moved.mapValue(lambda vert_outbox: vert_outbox[0])
        msgs = moved.flatMap(lambda kv: kv[1][1])
        verts.count()                         # one evaluation of the superstep; the counters are read after it
        return verts, msgs, sent.value, active.value, moved

    @classmethod
    def addAggregatorArg(cls, compute):
        def with_unused_arguments(vert, messages, aggregator, superstep):
            return compute(vert, messages)
        return with_unused_arguments
