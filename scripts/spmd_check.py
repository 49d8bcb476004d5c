#!/usr/bin/env python
"""The operator surface with one driver process per GPU (dpark_b200/spmd.py), on real GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 \
        scripts/spmd_check.py [lines] [vertices edges]

Runs (1) the wc.py pipeline (BASELINE config 1 shape: Zipf words, flatMap -> reduceByKey(+, 6) -> saveAsTextFile) and
(2) Bagel PageRank (config 5 shape, 10 supersteps), both through DparkContext on every rank, and checks on rank 0:
word counts == a Python Counter, the part files written by their owning ranks hold exactly those counts, PageRank ==
a numpy power iteration (1e-9).  Prints timings."""
import collections
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    lines_n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    nv_ = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
    ne = int(sys.argv[3]) if len(sys.argv) > 3 else 500_000
    sys.argv = sys.argv[:1]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    tmp = os.path.join(tempfile.gettempdir(), "dpk_spmd_check")
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp)
    rng = np.random.default_rng(1)
    vocab = 50_000
    w = 1.0 / np.arange(1, vocab + 1)
    ids = np.searchsorted(np.cumsum(w / w.sum()), rng.random(lines_n * 10)).reshape(lines_n, 10)
    inp = os.path.join(tmp, "in.txt")
    from dpark_b200 import DparkContext, bagel
    from dpark_b200 import _native as nvl
    dc = DparkContext("local")
    dc.init()                                  # joins the process group under torchrun
    import torch.distributed as dist
    if rank == 0:
        with open(inp, "w") as f:
            for row in ids:
                f.write(" ".join("w%d" % i for i in row) + "\n")
    if world > 1:
        dist.barrier()

    def fm(x):
        for wd in x.strip().split():
            yield (wd, 1)
    out = os.path.join(tmp, "out")
    l0 = nvl.launch_count()
    t0 = time.perf_counter()
    counts = dc.textFile(inp, numSplits=4 * world).flatMap(fm).reduceByKey(lambda x, y: x + y, numSplits=6)
    counts.map(lambda x: " ".join(list(map(str, x)))).saveAsTextFile(out, overwrite=False)
    got = dict(counts.collect())
    t_wc = time.perf_counter() - t0
    want = dict(collections.Counter("w%d" % i for i in ids.ravel()))
    assert got == want, "rank %d: word counts differ" % rank
    if world > 1:
        dist.barrier()
    if rank == 0:
        files = {}
        for fn in sorted(os.listdir(out)):
            for line in open(os.path.join(out, fn)):
                k, c = line.split()
                files[k] = int(c)
        assert files == want, "part files differ"
        print("wc: %d lines, %d tokens, %d words on %d GPU(s): %.2f s end to end (%.3e tokens/s), counts identical on every "
              "rank and in the %d part files; %d kernel launches on rank 0"
              % (lines_n, lines_n * 10, len(got), world, t_wc, lines_n * 10 / t_wc, len(os.listdir(out)),
                 nvl.launch_count() - l0))
    # ---- PageRank
    steps = 10
    rng = np.random.default_rng(5)
    pw = 1.0 / np.arange(1, nv_ + 1) ** 0.8
    src = rng.choice(nv_, size=ne, p=pw / pw.sum())
    dst = rng.integers(0, nv_, ne)
    order = np.argsort(src, kind="stable")
    src, dst = src[order], dst[order]
    starts = np.searchsorted(src, np.arange(nv_ + 1))
    n = float(nv_)

    def compute(self, msg_sum, agg, superstep):
        new = 0.15 / n + 0.85 * msg_sum[0] if msg_sum and msg_sum[0] else self.value
        done = superstep >= steps - 1
        outbox = [] if done or not self.outEdges else [(e.target_id, new / len(self.outEdges)) for e in self.outEdges]
        return bagel.Vertex(self.id, new, self.outEdges, not done), outbox
    verts = dc.parallelize([(int(v), bagel.Vertex(int(v), 1.0 / n, [bagel.Edge(int(t)) for t in dst[starts[v]:starts[v + 1]]],
                                                   True)) for v in range(nv_)], 4 * world)
    t0 = time.perf_counter()
    res = bagel.Bagel.run(dc, verts, dc.parallelize([], 4 * world), compute, maxSuperstep=steps, numSplits=4 * world)
    ranks = dict((k, v.value) for k, v in res.collect())
    t_pr = time.perf_counter() - t0
    r = np.full(nv_, 1.0 / n)
    deg = (starts[1:] - starts[:-1]).astype(np.float64)
    for s in range(steps):
        if s > 0:
            r = np.where(inbox > 0, 0.15 / n + 0.85 * inbox, r)
        if s < steps - 1:
            inbox = np.zeros(nv_)
            np.add.at(inbox, dst, (r / np.maximum(deg, 1.0))[src])
    g = np.array([ranks[i] for i in range(nv_)])
    err = float(np.max(np.abs(g - r) / np.maximum(np.abs(r), 1e-300)))
    assert err < 1e-9, err
    if rank == 0:
        print("pagerank: %d vertices, %d edges, %d supersteps on %d GPU(s): %.2f s (%.2f s per superstep, %.3e "
              "edge-messages/s), max relative deviation from the numpy power iteration %.1e"
              % (nv_, ne, steps, world, t_pr, t_pr / steps, ne * (steps - 1) / t_pr, err))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
