#!/usr/bin/env python
"""BASELINE config 5 (examples/pagerank.py shape) through the operator surface on ONE GPU: a synthetic graph
(default 1e5 vertices / 1e6 edges; `pagerank_e2e.py 1000000 10000000` is the config's size), out-degrees Zipf-skewed,
10 supersteps of Bagel -- per superstep one combineByKey(sum) of the messages and one groupWith of vertices and
combined messages, both on the GPU shuffle (int vertex ids, float64 message sums).

Reports wall time per superstep and how much of it the GPU shuffles take (the vertex programs and the Vertex/Edge
objects are Python, as in the reference: the end-to-end number is host-bound), and checks the ranks against a plain
numpy power iteration of the same update rule (1e-9 relative)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    nv_ = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    ne = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    steps = 10
    sys.argv = sys.argv[:1]
    rng = np.random.default_rng(5)
    w = 1.0 / np.arange(1, nv_ + 1) ** 0.8
    src = rng.choice(nv_, size=ne, p=w / w.sum())          # out-degree skew
    dst = rng.integers(0, nv_, ne)
    order = np.argsort(src, kind="stable")
    src, dst = src[order], dst[order]
    starts = np.searchsorted(src, np.arange(nv_ + 1))
    from dpark_b200 import DparkContext, bagel
    from dpark_b200 import _native as nvl
    dc = DparkContext("local")
    n = float(nv_)

    def compute(self, msg_sum, agg, superstep):
        if msg_sum and msg_sum[0]:
            new = 0.15 / n + 0.85 * msg_sum[0]
        else:
            new = self.value
        done = superstep >= steps - 1
        out = [] if done or not self.outEdges else [(e.target_id, new / len(self.outEdges)) for e in self.outEdges]
        return bagel.Vertex(self.id, new, self.outEdges, not done), out

    t0 = time.perf_counter()
    verts = dc.parallelize([(int(v), bagel.Vertex(int(v), 1.0 / n, [bagel.Edge(int(t)) for t in dst[starts[v]:starts[v + 1]]],
                                                   True)) for v in range(nv_)], 8)
    t_build = time.perf_counter() - t0
    l0 = nvl.launch_count()
    t0 = time.perf_counter()
    out = bagel.Bagel.run(dc, verts, dc.parallelize([], 8), compute, maxSuperstep=steps, numSplits=8)
    got = dict((k, v.value) for k, v in out.collect())
    t_run = time.perf_counter() - t0
    launches = nvl.launch_count() - l0
    # numpy restatement of the same update rule
    r = np.full(nv_, 1.0 / n)
    deg = (starts[1:] - starts[:-1]).astype(np.float64)
    for s in range(steps):
        if s > 0:
            r = np.where(inbox > 0, 0.15 / n + 0.85 * inbox, r)
        if s < steps - 1:
            inbox = np.zeros(nv_)
            np.add.at(inbox, dst, (r / np.maximum(deg, 1.0))[src])
    g = np.array([got[i] for i in range(nv_)])
    err = float(np.max(np.abs(g - r) / np.maximum(np.abs(r), 1e-300)))
    assert err < 1e-9, err
    print("pagerank: %d vertices, %d edges, %d supersteps on one GPU: %.2f s (graph objects built in %.2f s), "
          "%.2f s per superstep, %d CUDA kernel launches, max relative deviation from the numpy power iteration %.2e"
          % (nv_, ne, steps, t_run, t_build, t_run / steps, launches, err))
    print("  %.3e edge-messages/s end to end (host-bound: the vertex programs are Python objects, as in the reference)"
          % (ne * (steps - 1) / t_run))


if __name__ == "__main__":
    main()
