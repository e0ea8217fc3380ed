"""Multi-GPU plumbing of the hot path: the path shards by SCENE (independent scenes never interact), so
ranks own disjoint scene ranges and no data-path collective exists.  torch.distributed is used only for the
barrier and the max-over-ranks time / sum-over-ranks unit reductions of a measurement."""
from __future__ import annotations


def shard_range(n_units: int, rank: int, world: int) -> range:
    """contiguous, balanced range of scene (or item) indices owned by `rank`: the first n % world ranks get one more"""
    per, rem = divmod(n_units, world)
    lo = rank * per + min(rank, rem)
    return range(lo, lo + per + (1 if rank < rem else 0))


def reduce_job(ms_local: float, units_local: float, dist=None, device=None):
    """(max over ranks of the step time, sum over ranks of the units processed) → whole-job throughput inputs"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms_local, units_local
    import torch
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    u = torch.tensor([units_local], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
