#!/usr/bin/env python
"""Run the operator-surface golden cases that the CPU suite checks through the stand-in engine
(tests/test_cogroup_host.py, test_misc_ops_host.py, test_bagel_host.py) through the REAL engine on a GPU.

    gpurun -- python scripts/surface_on_gpu.py

Not part of `pytest -m gpu` yet: the cogroup / join / fixSkew cases already are (tests/test_gpu_rdd.py); the
operators added after the round's GPU budget was spent (topByKey, uniq, hot, update, groupBy, percentilesByKey,
sort, Bagel) are listed here so that their first GPU run is one command.  Prints one line per case and exits
non-zero on the first failure class."""
import inspect
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import pytest  # noqa: F401  (the test modules import it)
    from tests import test_bagel_host, test_cogroup_host, test_misc_ops_host
    failed = 0
    for mod in (test_cogroup_host, test_misc_ops_host, test_bagel_host):
        for name, fn in sorted(vars(mod).items()):
            if not name.startswith("test_") or not callable(fn):
                continue
            params = [()]
            for mark in getattr(fn, "pytestmark", []):
                if mark.name == "parametrize":
                    names = [n.strip() for n in mark.args[0].split(",")]
                    params = [v if isinstance(v, (tuple, list)) and len(names) > 1 else (v,) for v in mark.args[1]]
            wants_engine = "standin_engine" in inspect.signature(fn).parameters
            for p in params:
                label = "%s.%s%s" % (mod.__name__.split(".")[-1], name,
                                     "[%s]" % (p[0]["name"] if p and isinstance(p[0], dict) and "name" in p[0] else
                                               ",".join(map(str, p))) if p else "")
                try:
                    fn(*p, **({"standin_engine": None} if wants_engine else {}))
                    print("ok    ", label)
                except Exception:
                    failed += 1
                    print("FAILED", label)
                    traceback.print_exc(limit=3)
    print("%d failed" % failed)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
