#!/usr/bin/env python
"""Multi-GPU parity check of the shuffle engine (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/multi_gpu_check.py

Every rank holds its own map splits; reduceByKey and groupByKey run through
map_side -> exchange (NCCL alltoallv) -> reduce_side / group_side; the per-partition
results are gathered on rank 0 and compared with the oracle run on the union of all
splits in (rank, split) order.  Test infrastructure: the oracle is only the checker."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from dpark_b200 import shuffle
    from oracle import oracle as orc

    dev = torch.device("cuda", local)
    M, n = 3, 400_000
    ok_all = True
    px = None
    try:
        from dpark_b200 import peer
        px = peer.PeerExchange(world * n + 4096, torch.int64, torch.int64, dev)   # P=1: one rank receives everything
    except Exception as e:
        if rank == 0:
            print("peer exchange unavailable:", type(e).__name__, e)
    for case, P, lo, hi, sb in (("uniform", 8 * world, 0, 2 ** 31, None), ("dups", 5, -3000, 3000, 2),
                                ("one_partition", 1, 0, 1000, 0)):
        def make(r):
            rng = np.random.default_rng(1000 * r + 7)
            k = rng.integers(lo, hi, n, dtype=np.int64)
            v = rng.integers(-1000, 1000, n, dtype=np.int64)
            return np.array_split(k, M), np.array_split(v, M)
        ks, vs = make(rank)
        res = shuffle.reduce_by_key([torch.from_numpy(x).to(dev) for x in ks],
                                    [torch.from_numpy(x).to(dev) for x in vs], P, "sum", sub_bits=sb)
        mine = [(p, k.cpu().numpy(), v.cpu().numpy()) for p, k, v in res]
        if px is not None:
            # fused scatter + exchange must deliver bit-identical receive buffers to the NCCL alltoallv
            sb_eff = shuffle.choose_sub_bits(n, P, world) if sb is None else sb
            kd = [torch.from_numpy(x).to(dev) for x in ks]
            vd = [torch.from_numpy(x).to(dev) for x in vs]
            rx_nccl = shuffle.exchange(shuffle.map_side(kd, vd, P, None, False, sb_eff))
            rx_peer = peer.map_side_push(px, kd, vd, P, None, sb_eff, unordered=False)   # stable mode: bit-comparable
            nrx = int(rx_peer.seg.sum().item())          # bound view: the whole receive buffer
            same = (nrx == rx_nccl.keys.numel() and torch.equal(rx_nccl.keys, rx_peer.keys[:nrx])
                    and torch.equal(rx_nccl.vals, rx_peer.vals[:nrx])
                    and torch.equal(rx_nccl.seg, rx_peer.seg) and rx_nccl.part_first == rx_peer.part_first)
            px.check()
            # the unordered fused form (TMA bulk stores into the peers' buffers) feeds the same merge: same partitions
            rx_f = peer.map_side_push(px, kd, vd, P, None, sb_eff, unordered=True)
            fk, fv, fpo, fcnt = shuffle.reduce_side(rx_f, "sum", P)
            px.check()
            fpo, fcnt = fpo.cpu().tolist(), fcnt.cpu().tolist()
            for j, (p, wk_, wv_) in enumerate(mine):
                gk_ = fk[fpo[j]:fpo[j] + fcnt[j]].cpu().numpy()
                gv_ = fv[fpo[j]:fpo[j] + fcnt[j]].cpu().numpy()
                o1, o2 = np.argsort(gk_), np.argsort(wk_)
                same = same and np.array_equal(gk_[o1], wk_[o2]) and np.array_equal(gv_[o1], wv_[o2])
            # ... and so must the block push (local scatter + dpk_copy_segments)
            rx_push = peer.exchange_push(px, shuffle.map_side(kd, vd, P, None, False, sb_eff), need_host_count=True)
            same = (same and torch.equal(rx_nccl.keys, rx_push.keys) and torch.equal(rx_nccl.vals, rx_push.vals)
                    and torch.equal(rx_nccl.seg, rx_push.seg) and rx_nccl.nparts == rx_push.nparts)
            # the pipelined step (groups of map splits x parts of every block, push under the kernels on both sides)
            kall, vall = torch.from_numpy(np.concatenate(ks)).to(dev), torch.from_numpy(np.concatenate(vs)).to(dev)
            cuts = [0, n // 4, n // 2, n - 7, n]
            kcs = [kall[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
            vcs = [vall[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
            for gq in ((2, 2), (4, 1), (3, 4)):
                parts_res = peer.shuffle_pipelined(px, kcs, vcs, P, "sum", None, sb_eff, gq[0], gq[1])
                px.check()
                got_parts = {}
                for ok_, ov_, po_, cnt_, pf_, np_ in parts_res:
                    po_l, cnt_l = po_.cpu().tolist(), cnt_.cpu().tolist()
                    for j in range(np_):
                        got_parts[pf_ + j] = (ok_[po_l[j]:po_l[j] + cnt_l[j]].cpu().numpy(),
                                              ov_[po_l[j]:po_l[j] + cnt_l[j]].cpu().numpy())
                same = same and sorted(got_parts) == [p for p, _, _ in mine]
                for p, wk_, wv_ in mine:
                    if p in got_parts:
                        gk_, gv_ = got_parts[p]
                        o1, o2 = np.argsort(gk_), np.argsort(wk_)
                        same = same and np.array_equal(gk_[o1], wk_[o2]) and np.array_equal(gv_[o1], wv_[o2])
            flags = [None] * world
            dist.all_gather_object(flags, bool(same))
            if rank == 0:
                print("case %-14s fused scatter == block push == NCCL alltoallv, pipelined step == plain step on every rank: %s" % (case, all(flags)))
                ok_all &= all(flags)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        # groupByKey: values = global row ids so the (rank, split, position) order is checkable
        base = rank * n
        vid = np.array_split(np.arange(base, base + n, dtype=np.int64), M)
        mo = shuffle.map_side([torch.from_numpy(x).to(dev) for x in ks],
                              [torch.from_numpy(x).to(dev) for x in vid], P)
        rx = shuffle.exchange(mo)
        gk, gs, ng, ov, off = shuffle.group_side(rx, P)
        G = int(ng.item())
        gmine = (rx.part_first, rx.nparts, gk[:G].cpu().numpy(), gs[:G + 1].cpu().numpy(), ov.cpu().numpy(),
                 off.cpu().numpy())
        ggath = [None] * world
        dist.all_gather_object(ggath, gmine)
        if rank == 0:
            allk, allv, allid = [], [], []
            for r in range(world):
                a, b = make(r)
                allk += a
                allv += b
                allid += np.array_split(np.arange(r * n, r * n + n, dtype=np.int64), M)
            want = orc.reduce_by_key(allk, allv, P, "sum")
            seen = set()
            for r in range(world):
                for p, gk_, gv_ in gathered[r]:
                    assert p not in seen
                    seen.add(p)
                    o1, o2 = np.argsort(gk_), np.argsort(want[p][0])
                    good = np.array_equal(gk_[o1], want[p][0][o2]) and np.array_equal(gv_[o1], want[p][1][o2])
                    ok_all &= good
                    if not good:
                        print("MISMATCH reduce", case, "partition", p, "from rank", r)
            assert seen == set(range(P)), (seen, P)
            wantg = orc.group_by_key(allk, allid, P)
            for r in range(world):
                first, nparts, gkeys, gstarts, vals, poff = ggath[r]
                fg = np.searchsorted(gstarts[:-1], poff, side="left")
                for p in range(first, first + nparts):
                    wk, wo, wv = wantg[p]
                    g0, g1 = fg[p - first], fg[p - first + 1]    # poff: offsets of the rank's own partitions
                    got = {int(gkeys[g]): vals[gstarts[g]:gstarts[g + 1]] for g in range(g0, g1)}
                    good = len(got) == len(wk)
                    for i, key in enumerate(wk.tolist()):
                        good &= key in got and np.array_equal(got[key], wv[wo[i]:wo[i + 1]])
                    ok_all &= bool(good)
                    if not good:
                        print("MISMATCH group", case, "partition", p, "from rank", r)
            print("case %-14s P=%-3d world=%d: %s" % (case, P, world, "ok" if ok_all else "FAILED"))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok_all:
        sys.exit(1)


if __name__ == "__main__":
    main()
