#!/usr/bin/env python
"""Probe: can this torch build give each rank a device pointer into its peers' memory
(torch.distributed._symmetric_memory), so that the map-side scatter kernel can store rows
straight into the destination GPU's receive buffer over NVLink?  Run under torchrun."""
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    try:
        import torch.distributed._symmetric_memory as symm
        n = 1 << 26   # 64 Mi int64 = 512 MiB
        t = symm.empty(n, dtype=torch.int64, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
        t.fill_(rank)
        hdl.barrier()
        peer = (rank + 1) % world
        pbuf = hdl.get_buffer(peer, (n,), torch.int64)
        print("rank", rank, "local ptr %x peer ptr %x" % (t.data_ptr(), pbuf.data_ptr()), "ptrs", list(hdl.buffer_ptrs)[:world],
              flush=True)
        src = torch.full((n,), 100 + rank, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        hdl.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            pbuf.copy_(src)           # P2P store into the peer's buffer
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        hdl.barrier()
        want = 100 + (rank - 1) % world
        ok = bool((t == want).all())
        print("rank %d: peer write %s, %.1f GB/s" % (rank, "ok" if ok else "WRONG", n * 8 / ms / 1e6), flush=True)
        # our own kernel storing through the peer pointer: partition into the peer buffer
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from dpark_b200 import _native as nv
        k = torch.randint(0, 2 ** 31, (1 << 22,), dtype=torch.int64, device=dev)
        v = torch.arange(1 << 22, dtype=torch.int64, device=dev)
        counts, ws = nv.partition_count(k, 8)
        base = torch.zeros(8, dtype=torch.int64, device=dev)
        base[1:] = torch.cumsum(counts, 0)[:-1]
        outk = pbuf[: 1 << 22]
        outv = pbuf[1 << 22: 1 << 23]
        hdl.barrier()
        nv.partition_scatter(k, v, 8, base, outk, outv, ws)
        torch.cuda.synchronize()
        hdl.barrier()
        print("rank %d: kernel stores through the peer pointer ok (checksum %d)" % (rank, int(t[: 1 << 22].sum() & 0xffff)),
              flush=True)
    except Exception as e:
        print("rank", rank, "symmetric memory probe FAILED:", type(e).__name__, e, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
