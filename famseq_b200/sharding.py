"""Variant sharding across ranks (one process per GPU).

Every variant is independent, so a run over V sites on G GPUs is G runs over contiguous slices with no
data-path collective: rank r owns [lo, hi) and passes `v_offset = lo` to the
engine, which keys the Gibbs sampler's random stream by the GLOBAL site index -- any sharding produces the
same bytes.  `torch.distributed` is used only for timing (max over ranks) and, optionally, to gather outputs.
"""
from __future__ import annotations


def shard_range(n_variants: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks get the extras."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_variants, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(dist, value: float, device=None) -> float:
    """Max of a per-rank scalar (elapsed milliseconds) over all ranks; identity when not distributed."""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_to_rank0(dist, array, n_variants: int):
    """Gathers the per-rank output slices (numpy arrays whose first axis is the rank's variants) on rank 0, in
    site order.  This is the only communication a multi-GPU run needs, and only if one process must hold all
    results; the command line writes per-rank slices straight to their place in the output instead."""
    import numpy as np
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_variants, r, world) for r in range(world)]
    # dist.gather needs equally sized tensors: pad every slice to the largest one, trim on rank 0
    rows = max(hi - lo for lo, hi in sizes)
    mine = np.ascontiguousarray(array)
    padded = np.zeros((rows,) + mine.shape[1:], dtype=mine.dtype)
    padded[:mine.shape[0]] = mine
    mine_t = torch.from_numpy(padded)
    if rank == 0:
        parts = [torch.empty_like(mine_t) for _ in sizes]
        dist.gather(mine_t, parts, dst=0)
        return np.concatenate([p.numpy()[:hi - lo] for p, (lo, hi) in zip(parts, sizes)])
    dist.gather(mine_t, None, dst=0)
    return None
