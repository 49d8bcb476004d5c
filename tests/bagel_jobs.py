"""Two Bagel jobs used by the golden generator (run on the reference) and by the tests (run on this package):
the module that provides Vertex / Edge / Bagel / BasicCombiner / Aggregator is passed in, so the same job
text drives both implementations."""


def random_graph(rnd, nv, deg, ring=False, int_ids=False):
    """[[id, [target ids]], ...]: every vertex has >= 1 out-edge (PageRank divides by the out-degree)."""
    ids = list(range(nv)) if int_ids else ["v%02d" % i for i in range(nv)]
    g = []
    for i, v in enumerate(ids):
        if ring:
            targets = [ids[(i + 1) % nv]]
        else:
            targets = sorted(set(rnd.sample(ids, deg)) - {v}) or [ids[(i + 1) % nv]]
        g.append([v, targets])
    return g


def run_pagerank(dc, bagel, graph, parts):
    n = len(graph)
    eps = 0.01 / n

    def compute(self, message_sum, agg, superstep):
        if message_sum and message_sum[0]:
            new = 0.15 / n + 0.85 * message_sum[0]
        else:
            new = self.value
        done = (superstep >= 10 and abs(new - self.value) < eps) or superstep > 30
        out = [] if done else [(e.target_id, new / len(self.outEdges)) for e in self.outEdges]
        return bagel.Vertex(self.id, new, self.outEdges, not done), out

    verts = dc.parallelize([(v, bagel.Vertex(v, 1.0 / n, [bagel.Edge(t) for t in ts], True)) for v, ts in graph], parts)
    msgs = dc.parallelize([], parts)
    out = bagel.Bagel.run(dc, verts, msgs, compute, numSplits=parts)
    return dict((k, v.value) for k, v in out.collect())


class ActiveCount(object):
    """Duck-typed Bagel aggregator (module level: the reference pickles what it ships to tasks)."""

    def createAggregator(self, vert):
        return 1 if vert.active else 0

    def mergeAggregators(self, a, b):
        return a + b

    mergeAggregator = mergeAggregators


def run_maxprop(dc, bagel, graph, parts):
    """Every vertex learns the largest id that can reach it (BasicCombiner(max)); an Aggregator counts the
    active vertices before each superstep and every vertex remembers the last count it was shown."""

    def compute(self, inbox, agg, superstep):
        best = max([self.value] + [m for m in inbox if m is not None])
        changed = best > self.value or superstep == 0
        out = [(e.target_id, best) for e in self.outEdges] if changed else []
        vert = bagel.Vertex(self.id, best, self.outEdges, changed)
        vert.last_agg = agg
        return vert, out

    verts = dc.parallelize([(v, bagel.Vertex(v, v, [bagel.Edge(t) for t in ts], True)) for v, ts in graph], parts)
    msgs = dc.parallelize([], parts)
    out = bagel.Bagel.run(dc, verts, msgs, compute, combiner=bagel.BasicCombiner(max), aggregator=ActiveCount(),
                          numSplits=parts)
    return dict((k, [v.value, v.last_agg]) for k, v in out.collect())
