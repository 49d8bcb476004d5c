"""Partitioner / Aggregator records of the shuffle -- the reference's names and
semantics (dpark/dependency.py:67-75, 107-161, 209-239), written for a columnar
GPU engine: they describe WHAT to compute; the kernels in csrc/ do it.
"""


class Aggregator(object):
    """Three-callback record (dpark/dependency.py:121-137).  The callbacks are
    never run per row here: `dpark_b200.trace` maps them onto a kernel op."""

    def __init__(self, createCombiner, mergeValue, mergeCombiners):
        self.createCombiner = createCombiner
        self.mergeValue = mergeValue
        self.mergeCombiners = mergeCombiners


class AddAggregator(object):
    """dpark/dependency.py:140-148 -- sum."""

    def createCombiner(self, x):
        return x

    def mergeValue(self, s, x):
        return s + x

    def mergeCombiners(self, x, y):
        return x + y


class GroupByAggregator(object):
    """dpark/dependency.py:107-118 -- groupByKey: list of values per key."""

    def createCombiner(self, x):
        return [x]

    def mergeValue(self, c, x):
        c.append(x)
        return c

    def mergeCombiners(self, x, y):
        x.extend(y)
        return x


class MergeAggregator(GroupByAggregator):
    """dpark/dependency.py:151-161 -- same list-building contract."""


class Partitioner(object):
    @property
    def numPartitions(self):
        raise NotImplementedError

    def getPartition(self, key):
        raise NotImplementedError


class HashPartitioner(Partitioner):
    """portable_hash(key) floor-mod n, or bisect over `thresholds` (n-1 ascending
    hash values) -- dpark/dependency.py:218-239."""

    def __init__(self, partitions, thresholds=None):
        self.partitions = max(1, int(partitions))
        self.thresholds = None if thresholds is None else [int(t) for t in thresholds]
        if self.thresholds is not None and len(self.thresholds) != self.partitions - 1:
            raise AssertionError("thresholds must have partitions-1 entries")

    @property
    def numPartitions(self):
        return self.partitions

    def getPartition(self, key):
        """Partition of ONE key.  Evaluated by the same CUDA kernels as the
        bulk path (a 1-row launch); used by lookup()."""
        from .columnar import partition_of_key
        return partition_of_key(key, self.partitions, self.thresholds)

    def __eq__(self, other):
        return isinstance(other, HashPartitioner) and other.partitions == self.partitions and \
            other.thresholds == self.thresholds

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((self.partitions, None if self.thresholds is None else tuple(self.thresholds)))


class RangePartitioner(Partitioner):
    """dpark/dependency.py:242-258: partition = number of bounds <= key (mirrored when reverse).  Evaluated on the
    host: `RDD.sort` routes every element to its range first and shuffles on the range index."""

    def __init__(self, keys, reverse=False):
        self.keys = sorted(keys)
        self.reverse = reverse

    @property
    def numPartitions(self):
        return len(self.keys) + 1

    def getPartition(self, key):
        import bisect
        idx = bisect.bisect(self.keys, key)
        return len(self.keys) - idx if self.reverse else idx

    def __eq__(self, other):
        return isinstance(other, RangePartitioner) and other.keys == self.keys and other.reverse == self.reverse

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((tuple(self.keys), self.reverse))


class ShuffleDependency(object):
    """dpark/dependency.py:67-75."""

    def __init__(self, shuffleId, rdd, aggregator, partitioner, rddconf=None):
        self.shuffleId = shuffleId
        self.rdd = rdd
        self.aggregator = aggregator
        self.partitioner = partitioner
        self.rddconf = rddconf
        self.isShuffle = True
