"""The operator surface under torch.distributed: one process per GPU, every rank runs the SAME driver script.

The reference has one driver and many executors (schedule.py / executor.py); the B200 path under `torchrun` is SPMD:
every rank builds the same lineage, and

  * a job (`ctx.runJob`) computes on each rank only the partitions the rank OWNS (contiguous blocks, the same rule
    as the shuffle's partition ownership, `shuffle.owner_blocks`) and all-gathers the per-partition results, so every
    rank's script sees the full result and continues identically (`saveAsTextFile` writes each part file on its owner);
  * a shuffle ingests on each rank the parent splits the rank owns, exchanges rows between the GPUs (device columns
    over NCCL / NVLink for numeric reduceByKey; rows whose values are Python objects or whose keys are strings are
    routed to the owning rank as pickled columns and grouped / merged there on the GPU), and shares the small
    per-partition results with every rank;
  * accumulators are summed over the ranks after every job.

Collectives must be entered by every rank at the same point of the script: before a job runs, every shuffle of its
lineage that has not run yet is materialised in a deterministic order (`materialize_lineage`), also on ranks that
own none of the requested partitions.  With one process (no process group) every function here is the identity.
"""
import pickle

import torch


def rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def owner_of(index, n, world):
    """Rank owning split `index` of an RDD with n splits: contiguous blocks of ceil(n / world)."""
    per = (n + world - 1) // world
    return min(world - 1, index // per) if per else 0


def my_indices(n, rank, world):
    return [i for i in range(n) if owner_of(i, n, world) == rank]


def _comm_device():
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def all_gather_objects(obj):
    """[obj of rank 0, obj of rank 1, ...] on every rank."""
    import torch.distributed as dist
    rank, world = rank_world()
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def all_to_all_objects(per_dest):
    """per_dest[d] = object for rank d  ->  [object from rank 0, object from rank 1, ...].  Pickled once, moved as one
    variable-size all_to_all of bytes (not G all-gathers)."""
    import torch.distributed as dist
    rank, world = rank_world()
    if world == 1:
        return [per_dest[0]]
    dev = _comm_device()
    blobs = [pickle.dumps(o, protocol=pickle.HIGHEST_PROTOCOL) for o in per_dest]
    send_sizes = torch.tensor([len(b) for b in blobs], dtype=torch.int64, device=dev)
    recv_sizes = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_sizes, send_sizes)
    rs = recv_sizes.cpu().tolist()
    send = torch.frombuffer(bytearray(b"".join(blobs)) or bytearray(1), dtype=torch.uint8)
    if not sum(len(b) for b in blobs):
        send = send[:0]
    send = send.to(dev)
    recv = torch.empty(sum(rs), dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv, send, rs, [len(b) for b in blobs])
    raw = recv.cpu().numpy().tobytes()
    out, at = [], 0
    for n in rs:
        out.append(pickle.loads(raw[at:at + n]))
        at += n
    return out


def barrier():
    import torch.distributed as dist
    if rank_world()[1] > 1:
        dist.barrier()


def agree_max(value):
    """max of an int over the ranks (sizes that every rank must choose identically)."""
    import torch.distributed as dist
    rank, world = rank_world()
    if world == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=_comm_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def materialize_lineage(rdd, seen=None):
    """Run every shuffle of rdd's lineage that has not run yet, parents first, in the (deterministic) order of the
    lineage walk -- on every rank."""
    if seen is None:
        seen = set()
    if id(rdd) in seen:
        return
    seen.add(id(rdd))
    if rdd.should_cache and rdd._cache is not None and len(rdd._cache) == len(rdd.splits):
        return                                  # fully cached: nothing upstream will be read
    for p in rdd.parents():
        materialize_lineage(p, seen)
    mat = getattr(rdd, "_materialize", None)
    if mat is not None:
        mat()


# ---- accumulators --------------------------------------------------------------------------------------------------
_live_accumulators = []


def register_accumulator(acc):
    import weakref
    _live_accumulators.append(weakref.ref(acc))


def sync_accumulators():
    """After a job: every rank adds the other ranks' contributions since the last sync."""
    rank, world = rank_world()
    if world == 1:
        return
    alive = [r() for r in _live_accumulators]
    _live_accumulators[:] = [r for r, a in zip(list(_live_accumulators), alive) if a is not None]
    accs = sorted((a for a in alive if a is not None), key=lambda a: a.id)
    deltas = all_gather_objects([(a.id, a._take_delta()) for a in accs])
    for r, lst in enumerate(deltas):
        if r == rank:
            continue
        for (aid, d), a in zip(lst, accs):
            if aid == a.id and d is not None:
                a._absorb(d)
