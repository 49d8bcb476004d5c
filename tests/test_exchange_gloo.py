"""The N>1 path on CPU: world_size-2 `gloo` run of the exchange step (the one
collective of the shuffle) and of the ownership plan.  The bucket-major send
buffers are produced by the ORACLE here (tests may use it; the kernels need a
GPU); what is under test is dpark_b200.shuffle.exchange / owner_blocks: split
sizes, the counts matrix (MapOutputTracker replacement), and that every
reducer receives exactly its partitions' rows in source-rank order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_rank_data(rank, n, seed):
    rng = np.random.default_rng(seed + rank)
    k = rng.integers(-5000, 5000, n, dtype=np.int64)
    v = rng.integers(0, 1000, n, dtype=np.int64) + rank * 1_000_000
    return k, v


def _worker(rank, world, port, P, sub_bits, n, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dpark_b200 import shuffle
        k, v = _make_rank_data(rank, n, 77)
        # bucket-major map output with the product's layout: partition-major, sub-buckets inside
        # (oracle partition ids; the sub-bucket split is emulated with a second stable key)
        pid = orc.partition_vec(orc.hash_vec(k), P).astype(np.int64)
        sub = (np.abs(k) % (1 << sub_bits)).astype(np.int64) if sub_bits else np.zeros(n, np.int64)
        bucket = pid * (1 << sub_bits) + sub
        order = np.argsort(bucket, kind="stable")
        F = P << sub_bits
        offs = np.zeros(F + 1, dtype=np.int64)
        offs[1:] = np.cumsum(np.bincount(bucket, minlength=F))
        mo = shuffle.MapOutput(torch.from_numpy(k[order]), torch.from_numpy(v[order]), torch.from_numpy(offs), P, sub_bits)
        rx = shuffle.exchange(mo)
        out_q.put((rank, rx.part_first, rx.nparts, rx.keys.numpy().copy(), rx.vals.numpy().copy(),
                   rx.seg.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P,sub_bits", [(8, 0), (5, 2), (2, 3), (3, 0)])
def test_exchange_world2_gloo(P, sub_bits):
    world, n = 2, 20000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, P, sub_bits, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, first, nparts, keys, vals, seg = q.get(timeout=120)
        got[r] = (first, nparts, keys, vals, seg)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from dpark_b200.shuffle import owner_blocks
    blocks = owner_blocks(P, world)
    data = [_make_rank_data(r, n, 77) for r in range(world)]
    for r in range(world):
        first, nparts, keys, vals, seg = got[r]
        assert (first, first + nparts) == (blocks[r], blocks[r + 1])
        # expected: for each source rank in order, its rows whose partition this rank owns,
        # bucket-major and stable inside each bucket
        exp_k, exp_v = [], []
        for s in range(world):
            k, v = data[s]
            pid = orc.partition_vec(orc.hash_vec(k), P).astype(np.int64)
            sub = (np.abs(k) % (1 << sub_bits)).astype(np.int64) if sub_bits else np.zeros(n, np.int64)
            bucket = pid * (1 << sub_bits) + sub
            order = np.argsort(bucket, kind="stable")
            mine = (pid[order] >= blocks[r]) & (pid[order] < blocks[r + 1])
            exp_k.append(k[order][mine])
            exp_v.append(v[order][mine])
            cnt = np.bincount(bucket[(pid >= blocks[r]) & (pid < blocks[r + 1])] - (blocks[r] << sub_bits),
                              minlength=nparts << sub_bits)
            assert np.array_equal(seg[s], cnt)
        assert np.array_equal(keys, np.concatenate(exp_k))
        assert np.array_equal(vals, np.concatenate(exp_v))
    # every row arrived exactly once
    assert sum(len(got[r][2]) for r in range(world)) == world * n


def test_owner_blocks_cover_all_partitions():
    from dpark_b200.shuffle import owner_blocks
    for P in (1, 2, 7, 8, 64, 100):
        for G in (1, 2, 3, 4, 8):
            b = owner_blocks(P, G)
            assert b[0] == 0 and b[-1] == P and len(b) == G + 1
            assert all(b[i] <= b[i + 1] for i in range(G))
