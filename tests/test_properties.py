"""Property tests (hypothesis) of the host-side planning code: ownership blocks, sub-bucket choice, the segment
table of the block-push exchange, the merging digest.  CPU-only."""
import bisect
import math

import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from dpark_b200 import peer, quantiles, shuffle

# the same examples on every run (this suite gates a build; the search for new counter-examples belongs to development)
settings.register_profile("dpk", derandomize=True, deadline=None, database=None)
settings.load_profile("dpk")


@given(st.integers(1, 5000), st.integers(1, 64))
def test_owner_blocks_are_a_contiguous_cover(P, G):
    b = shuffle.owner_blocks(P, G)
    assert len(b) == G + 1 and b[0] == 0 and b[-1] == P
    assert all(x <= y for x, y in zip(b, b[1:]))
    sizes = [y - x for x, y in zip(b, b[1:])]
    assert max(sizes) == -(-P // G)                      # ceil(P / G) partitions per owning rank ...
    nz = [s for s in sizes if s]
    assert all(s == max(sizes) for s in nz[:-1])         # ... every owner but the last one is full


@given(st.integers(0, 10 ** 11), st.integers(1, 4096))
def test_choose_sub_bits_respects_the_kernel_limits(n, P):
    sb = shuffle.choose_sub_bits(n, P)
    F = P << sb
    assert 0 <= sb <= 12 and F <= max(P, 1024) and F <= 4096
    if sb:                                               # sub-buckets only when a partition is too big
        assert n / float(P << (sb - 1)) > shuffle.TARGET_BUCKET_ROWS or (P << sb) <= 1024


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 8), st.integers(1, 12), st.integers(0, 3), st.data())
def test_push_plan_tiles_every_receive_buffer_exactly(G, P, sb, data):
    F = P << sb
    counts = np.array(data.draw(st.lists(st.lists(st.integers(0, 9), min_size=F, max_size=F), min_size=G, max_size=G)),
                      dtype=np.int64)
    blocks = [b << sb for b in shuffle.owner_blocks(P, G)]
    ac = torch.from_numpy(counts)
    covered = [np.zeros(int(counts[:, blocks[d]:blocks[d + 1]].sum()), dtype=np.int32) for d in range(G)]
    for s in range(G):
        send_first, dst_first, rows, recv_total = peer.push_plan(ac, blocks, s)
        assert recv_total.tolist() == [len(c) for c in covered]
        assert int(rows.sum()) == int(counts[s].sum())            # everything a rank holds is sent somewhere
        ends = (send_first + rows).tolist()
        assert send_first.tolist() == [0] + ends[:-1]             # its blocks are adjacent in its own buffer
        for d in range(G):
            a, c = int(dst_first[d]), int(rows[d])
            covered[d][a:a + c] += 1
    assert all((c == 1).all() for c in covered)                   # no gap, no overlap in any receive buffer


@settings(max_examples=40, deadline=None)
@given(st.lists(st.lists(st.floats(-1e9, 1e9), min_size=0, max_size=400), min_size=1, max_size=6),
       st.lists(st.floats(0, 100), min_size=1, max_size=8))
def test_digest_quantiles_are_monotone_and_inside_the_data(parts, pcts):
    flat = sorted(x for p in parts for x in p)
    got = quantiles.percentiles_of_partitions(parts, sorted(pcts))
    if not flat:
        assert all(math.isnan(g) for g in got)
        return
    assert all(flat[0] <= g <= flat[-1] for g in got)
    assert all(a <= b + 1e-6 * max(1.0, abs(b)) for a, b in zip(got, got[1:]))
    # rank error of the estimate stays small (compression 100: a few percent at most in the middle)
    for p, g in zip(sorted(pcts), got):
        lo = bisect.bisect_left(flat, g) / len(flat)
        hi = bisect.bisect_right(flat, g) / len(flat)
        q = p / 100.
        err = 0.0 if lo <= q <= hi else min(abs(lo - q), abs(hi - q))
        assert err <= 0.05 + 1.0 / len(flat)


@given(st.lists(st.integers(-2 ** 62, 2 ** 62), min_size=0, max_size=300), st.integers(2, 9))
def test_skew_thresholds_are_strictly_increasing_ints(hashes, splits):
    thr, n = quantiles.skew_thresholds([hashes[::2], hashes[1::2]], splits)
    assert n == len(thr) + 1 <= splits
    assert all(isinstance(t, int) for t in thr) and all(a < b for a, b in zip(thr, thr[1:]))
