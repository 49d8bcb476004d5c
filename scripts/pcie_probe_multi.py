#!/usr/bin/env python
"""Host-link bandwidth with ALL ranks copying at once (run under torchrun): pinned H2D and D2H, both directions
together, per rank and summed.  The e2e leg of bench.py moves 1.6 GB in and 1.56 GB out per batch and rank; this is
its bound at N GPUs (GPUs behind one PCIe switch share its uplink)."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dpark_b200 import shuffle
    local = int(os.environ.get("LOCAL_RANK", "0"))
    node = shuffle.bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 100_000_000          # int64 -> 0.8 GB
    h_in = torch.empty(n, dtype=torch.int64).pin_memory()
    h_out = torch.empty(n, dtype=torch.int64).pin_memory()
    d_in = torch.empty(n, dtype=torch.int64, device="cuda")
    d_out = torch.ones(n, dtype=torch.int64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    gb = n * 8 / 1e9

    def timed(fn, reps=4):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    res = []
    for name, fn in (("H2D alone", h2d), ("D2H alone", d2h), ("both", lambda: (h2d(), d2h()))):
        t = timed(fn)
        res.append((name, gb / t))
    out = [None] * world
    dist.all_gather_object(out, (rank, node, res))
    if rank == 0:
        for name_i in range(3):
            per = [o[2][name_i][1] for o in out]
            print("%-10s all %d ranks at once: per rank %s GB/s per direction, sum %.1f GB/s  (NUMA nodes %s)"
                  % (out[0][2][name_i][0], world, " ".join("%.1f" % x for x in per), sum(per), [o[1] for o in out]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
