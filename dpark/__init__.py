"""Alias so that unmodified DPark scripts (`from dpark import DparkContext,
optParser`, `import dpark.conf`) run on the B200 shuffle: importing `dpark`
hands back the dpark_b200 package itself."""
import sys as _sys

import dpark_b200 as _impl

_sys.modules[__name__] = _impl
for _name in ("conf", "rdd", "context", "dependency", "bagel", "accumulator"):
    _sys.modules[__name__ + "." + _name] = __import__("dpark_b200." + _name, fromlist=["_"])
