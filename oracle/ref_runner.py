#!/usr/bin/env python
"""Runs the UNMODIFIED reference (baseline/_ref/dpark, built by oracle/build_reference.py) on the host cores:
the CPU arm of bench.py (`--impl reference`, `cpu_baseline.kind == "reference"`).

Measurement infrastructure only -- executed by bench.py as a SUBPROCESS (the reference registers itself as the
`dpark` package, which in this repository's own processes is the alias of dpark_b200).

    python oracle/ref_runner.py --config c2 --rows 4000000 --splits 64 --parts 8 --procs 64

Prints one JSON line: rows, seconds of the whole job, seconds of `source.count()` alone, shuffle-only seconds
("shuffle-only time = job time - time of src.count() on the same source") and a
checksum the caller compares with its own generator.

What runs: DparkContext('process') with `-p procs` -> MultiProcessScheduler (dpark/schedule.py:841-910):
ShuffleMapTask._run per split (dpark/task.py:197-255: dict upsert per bucket, marshal + compress + file per
bucket), ShuffleFetcher + DiskHashMerger per reduce partition (dpark/shuffle.py:309-420, 524-623).  The four
modules this image lacks are stubbed and the forked workers get SURVEY Appendix B's tracker
shim, without which `-m process` returns empty shuffles at this commit; no reference file is edited.
"""
import argparse
import inspect
import json
import logging
import operator
import os
import sys
import time
import types
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def bootstrap():
    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m

    class Dict(dict):
        pass

    mod("addict", Dict=Dict)                                                     # schedule.py:16, executor.py:20
    mod("pymesos", MesosSchedulerDriver=object, MesosExecutorDriver=object, Executor=object,
        encode_data=lambda x: x, decode_data=lambda x: x)                        # schedule.py:17, executor.py:21
    mod("dpark.utils.recursion", get_recursion_depth=lambda: len(inspect.stack()))   # utils/__init__.py:119-125
    mod("lz4framed", compress=lambda s, *a, **k: zlib.compress(s, 1), decompress=zlib.decompress)
    # the repository root must NOT shadow the reference: drop it (and '') from sys.path
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    sys.path.insert(0, REF_DIR)


_ORIG_RUN_TASK = None


def _run_task_with_tracker(task, tid, environ):
    """Module-level (picklable by reference from the forked pool workers): SURVEY.md Appendix B."""
    from dpark.env import env
    if env.trackerServer is not None:
        env.trackerServer = None
        env.slave_started = False
        env.environ.update(environ)
        env.start_slave()
    return _ORIG_RUN_TASK(task, tid, environ)


def tracker_shim():
    """a forked pool worker still holds the master's TrackerServer object; make it ask the
    master's tracker like a Mesos executor would (dpark/shuffle.py:821-826, dpark/env.py:258-278)."""
    global _ORIG_RUN_TASK
    import dpark.schedule as S
    _ORIG_RUN_TASK = S.run_task_in_process
    S.run_task_in_process = _run_task_with_tracker


def gen_rows(config, rows, seed):
    """Same generators as bench.py's GPU arm (numpy twins of the torch calls; the row VALUES differ from the
    GPU run's, the distribution and types are the same)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    if config == "c4":
        keys = rng.integers(0, 2 ** 24, rows, dtype=np.int64).tolist()
        vals = rng.random(rows, dtype=np.float32).astype(np.float64).tolist()
    elif config == "c3":
        ranks = zipf_ranks(rng, rows, 1.1, 10 ** 9)
        keys = ((ranks * 0x9E3779B1) % (2 ** 31)).tolist()
        vals = list(range(rows))
    else:
        keys = rng.integers(0, 2 ** 31, rows, dtype=np.int64).tolist()
        vals = rng.integers(0, 2 ** 16, rows, dtype=np.int64).tolist()
    return list(zip(keys, vals))


def zipf_ranks(rng, n, s, support):
    """Inverse-CDF sample of Zipf(s) over ranks 1..support (continuous approximation of the CDF, exact enough
    for a skew benchmark): rank = ((1 - u * (1 - support^(1-s)))^(1/(1-s)))."""
    import numpy as np
    u = rng.random(n)
    a = 1.0 - s
    r = np.power(1.0 - u * (1.0 - float(support) ** a), 1.0 / a)
    return np.minimum(np.maximum(r.astype(np.int64), 1), support)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4"])
    ap.add_argument("--rows", type=int, default=1000000)
    ap.add_argument("--splits", type=int, default=8)
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--master", default="process", choices=["process", "local"])
    ap.add_argument("--seed", type=int, default=1000)
    args = ap.parse_args()
    procs = args.procs or (os.cpu_count() or 1)

    bootstrap()
    sys.argv = [sys.argv[0]] + (["-m", "process", "-p", str(procs)] if args.master == "process" else ["-m", "local"])
    import dpark                                                        # the REFERENCE
    assert os.path.abspath(dpark.__file__).startswith(REF_DIR), dpark.__file__
    from dpark import DparkContext
    logging.getLogger("dpark").setLevel(logging.ERROR)
    if args.master == "process":
        tracker_shim()
    loaded = sorted(set(os.path.relpath(m.__file__, ROOT) for m in list(sys.modules.values())
                        if getattr(m, "__file__", None) and m.__file__.endswith(".so") and
                        os.path.abspath(m.__file__).startswith(REF_DIR)))

    rows = gen_rows(args.config, args.rows, args.seed)
    ctx = DparkContext()
    ctx.start()
    src = ctx.parallelize(rows, args.splits)
    src.count()                                                         # warm the pool (fork, imports)
    t0 = time.perf_counter()
    n_src = src.count()
    t_src = time.perf_counter() - t0
    t0 = time.perf_counter()
    if args.config == "c3":
        out = src.groupByKey(args.parts)
        distinct = out.count()
    else:
        out = src.reduceByKey(operator.add, args.parts)
        distinct = out.count()
    t_job = time.perf_counter() - t0
    # result check on a second, collected run kept OUT of the timing: value checksum survives the shuffle
    check = None
    if args.config == "c2" and args.rows <= 2000000:
        got = src.reduceByKey(operator.add, args.parts).collect()
        check = (sum(v for _, v in got) == sum(v for _, v in rows)) and len(got) == distinct
    ctx.stop()
    print(json.dumps({"rows": n_src, "distinct": distinct, "job_s": t_job, "src_count_s": t_src,
                      "shuffle_s": max(t_job - t_src, 1e-9), "procs": procs, "master": args.master,
                      "splits": args.splits, "parts": args.parts, "config": args.config, "checksum_ok": check,
                      "native_so_loaded": loaded, "dpark_file": os.path.relpath(dpark.__file__, ROOT)}))
    sys.stdout.flush()
    os._exit(0)                                                         # the reference's atexit hooks can hang on pool teardown


if __name__ == "__main__":
    main()
