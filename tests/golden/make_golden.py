#!/usr/bin/env python
"""Generate golden vectors for the shuffle hot path FROM THE REAL REFERENCE.

Runs only in the build container (needs /root/reference, read-only).  It builds
an out-of-tree scratch copy of the reference under a temp dir (SURVEY.md §8(c)
recipe: cythonize -2 portable_hash.pyx, gcc crc32c, four stub modules), drives
the reference's own `DparkContext('local')` and writes the JSON fixtures that
sit next to this script.  The fixtures travel to the GPU box; this script and
the reference do not need to.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.json

Nothing under /root/reference is modified; nothing from it is copied into the
repository (only the *outputs* of running it are stored).
"""
import inspect
import json
import logging
import os
import random
import shutil
import subprocess
import sys
import sysconfig
import tempfile
import types
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DPARK_REFERENCE", "/root/reference")


# ----------------------------------------------------------------------------
# build + bootstrap the reference (scratch dir only)
# ----------------------------------------------------------------------------
def build_reference(scratch):
    shutil.copytree(os.path.join(REF, "dpark"), os.path.join(scratch, "dpark"))
    subprocess.check_call(
        ["cythonize", "-i", "-2", "dpark/portable_hash.pyx"], cwd=scratch,
        stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    inc = sysconfig.get_paths()["include"]
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    subprocess.check_call(
        ["gcc", "-O2", "-fPIC", "-shared", "-msse4.2", "-I" + inc,
         "dpark/utils/crc32c.c", "dpark/utils/crc32c_mod.c",
         "-o", "dpark/utils/crc32c" + ext], cwd=scratch)


def bootstrap(scratch):
    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m

    class Dict(dict):
        pass

    mod("addict", Dict=Dict)
    mod("pymesos", MesosSchedulerDriver=object, MesosExecutorDriver=object,
        Executor=object, encode_data=lambda x: x, decode_data=lambda x: x)
    mod("dpark.utils.recursion",
        get_recursion_depth=lambda: len(inspect.stack()))
    mod("lz4framed", compress=lambda s, *a, **k: zlib.compress(s, 1),
        decompress=zlib.decompress)
    sys.path.insert(0, scratch)
    sys.argv = [sys.argv[0]]


# ----------------------------------------------------------------------------
# typed JSON encoding of keys/values (ints can exceed 2**53, bytes, tuples …)
# ----------------------------------------------------------------------------
def enc(o):
    """Compact typed encoding: int -> JSON int (Python's json is exact for big
    ints), None -> null, everything else a one-key object."""
    import numpy as np
    if o is None:
        return None
    if isinstance(o, np.generic):
        return {"np": o.dtype.name, "v": repr(o.item())}
    if type(o) is bool:
        return {"bool": int(o)}
    if type(o) is int:
        return o
    if type(o) is float:
        return {"f": o.hex()}
    if type(o) is bytes:
        return {"b": o.hex()}
    if type(o) is str:
        # code points, so lone surrogates / non-BMP survive JSON untouched
        return {"s": [ord(c) for c in o]}
    if type(o) is tuple:
        return {"tu": [enc(x) for x in o]}
    if type(o) is list:
        return {"l": [enc(x) for x in o]}
    raise TypeError(type(o))


def main():
    scratch = tempfile.mkdtemp(prefix="dpark_ref_")
    try:
        build_reference(scratch)
        bootstrap(scratch)
        generate()
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


def generate():
    import numpy as np
    import dpark  # the REFERENCE (scratch copy)
    import dpark.conf
    from dpark import DparkContext
    from dpark.utils import portable_hash
    from dpark.dependency import HashPartitioner, Aggregator

    logging.getLogger("dpark").setLevel(logging.ERROR)
    rnd = random.Random(20260922)

    # ------------------------------------------------------------------ a1/a2
    keys = []
    # survey §8(c) table
    keys += [0, 1, 7, -1, -2, -5, 2 ** 31, 2 ** 61 - 1, 2 ** 61, 2 ** 63 - 1,
             -2 ** 63, 2 ** 64 + 5, 1.5, 0.0, None, b"", b"a", b"hello",
             b"\xff\x80", "hello", "你好", "w1", "w2", (), (1, 2),
             (b"a", 1), ("a", (1, 2.0)), (1, "a"), np.int64(7), np.int32(-1),
             np.float32(2.5), np.uint64(2 ** 63 + 1)]
    # integer edges around the Mersenne modulus and the int64 range
    M61 = 2 ** 61 - 1
    for base in (0, M61, 2 * M61, 3 * M61, 4 * M61, 2 ** 62, 2 ** 63 - 1):
        for d in (-2, -1, 0, 1, 2):
            v = base + d
            if -2 ** 63 <= v <= 2 ** 63 - 1:
                keys.append(v)
            if -2 ** 63 <= -v <= 2 ** 63 - 1:
                keys.append(-v)
    keys += [rnd.randint(-2 ** 63, 2 ** 63 - 1) for _ in range(1500)]
    keys += [rnd.randint(-2 ** 31, 2 ** 31 - 1) for _ in range(500)]
    keys += [rnd.randint(0, 1000) for _ in range(200)]
    # bytes: every byte value, signed-char path, lengths 0..48
    keys += [bytes([b]) for b in range(256)]
    keys += [bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 49)))
             for _ in range(600)]
    keys += [("w%d" % i).encode() for i in range(0, 3000, 37)]
    # str: ASCII, latin-1, BMP, astral
    alph = [chr(c) for c in list(range(32, 127)) + [0xe9, 0xff, 0x100, 0x4f60,
                                                    0x597d, 0xffff, 0x10000,
                                                    0x1f600, 0x10ffff]]
    keys += ["".join(rnd.choice(alph) for _ in range(rnd.randrange(0, 24)))
             for _ in range(600)]
    keys += ["w%d" % i for i in range(0, 3000, 37)]
    # floats
    keys += [float(i) for i in (-3, -1, 0, 1, 2, 2 ** 31, 2 ** 53, 2 ** 61 - 1,
                                2 ** 61, 2 ** 62, 2 ** 64)]
    keys += [float("inf"), float("-inf"), -0.0, 5e-324, -5e-324, 2.2250738585072014e-308,
             1.7976931348623157e308, 0.1, -0.1, 1e-10, 1e300, 3.141592653589793]
    keys += [rnd.uniform(-1e6, 1e6) for _ in range(300)]
    keys += [rnd.random() * 10 ** rnd.randint(-300, 300) * rnd.choice((1, -1))
             for _ in range(300)]
    keys += [np.float32(rnd.uniform(-100, 100)) for _ in range(50)]
    keys += [np.int32(rnd.randint(-2 ** 31, 2 ** 31 - 1)) for _ in range(50)]
    keys += [np.uint64(rnd.randint(0, 2 ** 64 - 1)) for _ in range(50)]
    # tuples (nested, mixed)
    def rnd_atom():
        c = rnd.randrange(5)
        if c == 0:
            return rnd.randint(-2 ** 63, 2 ** 63 - 1)
        if c == 1:
            return rnd.randint(-5, 5)
        if c == 2:
            return bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 6)))
        if c == 3:
            return "".join(rnd.choice(alph) for _ in range(rnd.randrange(0, 6)))
        return rnd.uniform(-10, 10)
    for _ in range(400):
        n = rnd.randrange(0, 6)
        t = tuple(rnd_atom() if rnd.random() < 0.8 else
                  tuple(rnd_atom() for _ in range(rnd.randrange(0, 3)))
                  for _ in range(n))
        keys.append(t)

    PS = [1, 2, 3, 4, 6, 8, 64, 1000]
    hash_rows = []
    for k in keys:
        h = portable_hash(k)
        hash_rows.append({"key": enc(k), "hash": h,
                          "part": [HashPartitioner(p).getPartition(k) for p in PS]})
    unhashable = []
    for k in (True, [1, 2], {1: 2}, 1 + 2j):
        try:
            portable_hash(k)
            unhashable.append({"repr": repr(k), "raises": None})
        except TypeError as e:
            unhashable.append({"repr": repr(k), "raises": "TypeError"})

    thr = sorted(rnd.randint(-2 ** 62, 2 ** 62) for _ in range(7))
    thr_rows = []
    hp = HashPartitioner(8, thresholds=thr)
    for h in [-5, 9, 10, 11, 99, 100, 101] + thr + [t - 1 for t in thr] + \
             [rnd.randint(-2 ** 63, 2 ** 63 - 1) for _ in range(200)]:
        # keys whose hash is the key itself: |h| < 2**61-1 not guaranteed, so
        # store the key and let the test hash it
        thr_rows.append({"key": h, "part": hp.getPartition(h)})
    small = HashPartitioner(3, thresholds=[10, 100])
    thr_small = [{"key": h, "part": small.getPartition(h)}
                 for h in (-5, 9, 10, 11, 99, 100, 101)]
    json.dump({"partitions": PS, "rows": hash_rows, "unhashable": unhashable,
               "thresholds": thr, "threshold_rows": thr_rows,
               "thresholds_small": [10, 100], "threshold_rows_small": thr_small},
              open(os.path.join(HERE, "hash_vectors.json"), "w"), separators=(",", ":"))

    # --------------------------------------------------------- a3..a11 shuffles
    dc = DparkContext("local")
    cases = []

    def run_reduce(name, rows, M, P, func_name, ordered=False, thresholds=None):
        funcs = {
            "add": lambda x, y: x + y,
            "min": lambda x, y: min(x, y),
            "max": lambda x, y: max(x, y),
            "mul": lambda x, y: x * y,
            "or": lambda x, y: x | y,
            "and": lambda x, y: x & y,
            "xor": lambda x, y: x ^ y,
        }
        rdd = dc.parallelize(rows, M)
        if thresholds is None:
            out = rdd.reduceByKey(funcs[func_name], P).glom().collect()
        else:
            f = funcs[func_name]
            agg = Aggregator(lambda x: x, f, f)
            out = rdd.combineByKey(agg, HashPartitioner(P, thresholds=thresholds)).glom().collect()
        # per-partition multiset → sorted for a canonical form
        canon = [sorted(([enc(k), enc(v)] for k, v in part), key=json.dumps) for part in out]
        splits = [len(x) for x in dc.parallelize(rows, M).glom().collect()]
        cases.append({"name": name, "op": "reduceByKey", "func": func_name,
                      "M": M, "P": P, "split_sizes": splits,
                      "thresholds": thresholds,
                      "rows": [[enc(k), enc(v)] for k, v in rows], "parts": canon})

    def run_group(name, rows, M, P):
        conf = dpark.conf.rddconf(ordered_group=True)
        out = dc.parallelize(rows, M).groupByKey(P, rddconf=conf).glom().collect()
        canon = [sorted(([enc(k), enc(list(v))] for k, v in part), key=json.dumps) for part in out]
        splits = [len(x) for x in dc.parallelize(rows, M).glom().collect()]
        cases.append({"name": name, "op": "groupByKey", "M": M, "P": P,
                      "split_sizes": splits,
                      "rows": [[enc(k), enc(v)] for k, v in rows], "parts": canon})

    # survey layout vector
    lay = [(k, 1) for k in [-1, -2, -3, -4, -5, 0, 1, 2, 3, 4, 5, 2 ** 61 - 1, 2 ** 61]] * 2
    run_reduce("layout_survey", lay, 3, 4, "add")
    # reference test_basic
    d = list(zip([1, 2, 3, 3], list(range(4, 8))))
    run_reduce("test_basic_reduce", d, 2, 2, "add")
    run_group("test_basic_group", d, 2, 2)
    run_group("ordering_vector", [(1, i) for i in range(12)], 4, 2)
    # random int64 keys / int64 values
    for i, (n, M, P, lo, hi) in enumerate([
            (1000, 4, 8, 0, 2 ** 31), (2000, 8, 8, -50, 50), (1500, 3, 5, -2 ** 63, 2 ** 63 - 1),
            (777, 7, 64, 0, 300), (1, 1, 4, 5, 6), (64, 8, 1, 0, 10)]):
        rows = [(rnd.randint(lo, hi - 1 if hi > lo else lo), rnd.randint(-2 ** 16, 2 ** 16))
                for _ in range(n)]
        for fn in ("add", "min", "max"):
            run_reduce("rand_i64_%d_%s" % (i, fn), rows, M, P, fn)
        run_group("rand_i64_group_%d" % i, rows, M, P)
    rows = [(rnd.randint(0, 40), rnd.randint(0, 2 ** 20)) for _ in range(600)]
    for fn in ("or", "and", "xor"):
        run_reduce("bitwise_%s" % fn, rows, 4, 4, fn)
    rows = [(rnd.randint(0, 20), rnd.randint(-3, 3) or 1) for _ in range(200)]
    run_reduce("mul_small", rows, 4, 4, "mul")
    # Zipf-ish skew
    zipf = [(min(int(1.0 / (1e-9 + rnd.random()) ** 1.1), 10 ** 6), i) for i in range(3000)]
    run_reduce("zipf_add", zipf, 8, 8, "add")
    run_group("zipf_group", zipf, 8, 8)
    # empty input, empty partitions
    run_reduce("empty", [], 2, 4, "add")
    run_group("empty_group", [], 2, 4)
    # thresholds (bisect path)
    rows = [(rnd.randint(-1000, 1000), 1) for _ in range(800)]
    run_reduce("thresholds_bisect", rows, 4, 4, "add", thresholds=[-500, 0, 500])
    # float32-like values: reference adds Python floats = float64
    rows = [(rnd.randint(0, 99), float(np.float32(rnd.random()))) for _ in range(4000)]
    run_reduce("f32_vals_add", rows, 8, 8, "add")
    run_reduce("f32_vals_min", rows, 8, 8, "min")
    run_reduce("f32_vals_max", rows, 8, 8, "max")
    # int32-range keys incl. -1 (hash -2)
    rows = [(rnd.choice([-1, -2, 0, 1, 2 ** 31 - 1, -2 ** 31]) if rnd.random() < .2
             else rnd.randint(-2 ** 31, 2 ** 31 - 1), rnd.random()) for _ in range(1000)]
    run_reduce("i32_keys_f_vals", rows, 4, 8, "add")
    # bytes / str keys (wc shape)
    vocab = ["w%d" % i for i in range(200)] + ["你好", "naïve", "\U0001f600x", ""]
    words = [rnd.choice(vocab) for _ in range(3000)]
    run_reduce("wc_str", [(w, 1) for w in words], 4, 6, "add")
    run_reduce("wc_bytes", [(w.encode("utf-8"), 1) for w in words], 4, 6, "add")
    run_group("group_str", [(w, i) for i, w in enumerate(words[:500])], 4, 3)
    # fixSkew layout pinned by the reference test (tests/test_rdd.py:259-266)
    dsk = list(zip(range(10), range(10))) + [(10, 10)] * 5
    out = dc.makeRDD(dsk, 10).groupByKey(3, fixSkew=1).map(lambda kv: (kv[0], list(kv[1]))).glom().collect()
    fix_skew = [sorted([[k, sorted(v)] for k, v in part]) for part in out]

    # wc.py end-to-end through textFile → files
    tmp = tempfile.mkdtemp(prefix="dpark_wc_")
    try:
        lines = [" ".join(rnd.choice(vocab[:60]) for _ in range(rnd.randrange(0, 12)))
                 for _ in range(400)]
        inp = os.path.join(tmp, "in.txt")
        with open(inp, "w", encoding="utf-8") as f:
            f.write("\n".join(lines) + "\n")
        outdir = os.path.join(tmp, "out")

        def fm(x):
            for w in x.strip().split():
                yield (w, 1)
        (dc.textFile(inp).flatMap(fm).reduceByKey(lambda x, y: x + y, numSplits=6)
           .map(lambda x: " ".join(list(map(str, x)))).saveAsTextFile(outdir, overwrite=False))
        files = sorted(os.listdir(outdir))
        per_file = {fn: sorted(open(os.path.join(outdir, fn), encoding="utf-8").read().splitlines())
                    for fn in files}
        wc = {"lines": lines, "files": per_file}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

    json.dump({"cases": cases, "fix_skew_test_basic": fix_skew, "wc": wc},
              open(os.path.join(HERE, "shuffle_cases.json"), "w"), separators=(",", ":"))
    dc.stop()
    print("wrote", len(hash_rows), "hash rows,", len(cases), "shuffle cases")


if __name__ == "__main__":
    main()
