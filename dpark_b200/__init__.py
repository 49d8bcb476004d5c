"""dpark_b200 -- the DPark shuffle hot path (reduceByKey / groupByKey) rebuilt
B200-native: hand-written sm_100a kernels behind a C ABI (include/dpark_b200.h),
torch CUDA tensors as the columnar partition container, one NCCL alltoallv as
the exchange.  The user-facing names are the reference's:

    from dpark_b200 import DparkContext, optParser      # or, unchanged scripts:
    from dpark import DparkContext, optParser           # via the `dpark` alias package

See DESIGN.md and INTEGRATION.md.
"""
from .errors import DparkUserFatalError  # noqa: F401
from .context import DparkContext, parser as optParser  # noqa: F401
from .dependency import (Aggregator, AddAggregator, GroupByAggregator, HashPartitioner,  # noqa: F401
                         MergeAggregator, RangePartitioner)
from . import conf  # noqa: F401

__all__ = ["DparkContext", "optParser", "DparkUserFatalError"]
__version__ = "0.1.0"
