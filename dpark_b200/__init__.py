"""dpark_b200 -- the DPark shuffle hot path (reduceByKey / groupByKey) rebuilt
B200-native: hand-written sm_100a kernels behind a C ABI (include/dpark_b200.h),
torch CUDA tensors as the columnar partition container, one NCCL alltoallv as
the exchange.  See DESIGN.md."""
from .errors import DparkUserFatalError  # noqa: F401

__version__ = "0.1.0"
