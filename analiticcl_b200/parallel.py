"""Query-batch partitioning across the GPUs of one box (mode 1).

The path shards by query: the index is replicated on every GPU, every rank looks up its own
contiguous slice of the batch, and nothing is exchanged on the data path.  `torch.distributed` is
only plumbing: a barrier for timing and an object gather of the result lists to rank 0.
Works with any backend (NCCL on GPUs; gloo in the CPU tests).
"""
import torch.distributed as dist


def partition(n, world):
    """Contiguous, balanced slices of range(n): the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def lookup_partitioned(queries, lookup, group=None, dst=0):
    """Each rank runs `lookup(list[str]) -> list[result]` on its slice; rank `dst` gets the full,
    input-ordered result list (other ranks get None)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = partition(len(queries), world)[rank]
    mine = lookup(queries[lo:hi])
    if world == 1:
        return mine
    gathered = [None] * world if rank == dst else None
    dist.gather_object(mine, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = []
    for part in gathered:
        out.extend(part)
    return out


def reduce_step_time(ms_local, units_local, device=None, group=None):
    """bench.py's reduction: (max over ranks of the step time, sum over ranks of the units)."""
    import torch
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()), float(u.item())
