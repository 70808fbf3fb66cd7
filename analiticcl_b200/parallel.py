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


def reduce_max_sum(times, units, device=None, group=None):
    """bench.py's reduction over the ranks: element-wise MAX of `times` (a step is as slow as its slowest rank) and
    element-wise SUM of `units` (every rank processed its own share).  Returns two lists of floats."""
    import torch
    t = torch.tensor([float(x) for x in times], dtype=torch.float64, device=device)
    u = torch.tensor([float(x) for x in units], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return t.tolist(), u.tolist()


def reduce_step_time(ms_local, units_local, device=None, group=None):
    """(max over ranks of the step time, sum over ranks of the units)."""
    t, u = reduce_max_sum([ms_local], [units_local], device=device, group=group)
    return t[0], u[0]
