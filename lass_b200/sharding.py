"""Clip sharding for multi-GPU inference: one process per GPU, clips are independent in eval mode
(BatchNorm uses running statistics, FiLM is per clip), so rank r separates a contiguous slab of the clips with
replicated weights and NO data-path collective.  ``torch.distributed`` is used only for the barrier / max-reduce
around timed regions and for the optional final gather of waveforms."""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slab [lo, hi) of rank ``rank``; the first ``n_clips % world`` ranks get one extra clip."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_bounds(n_clips: int, world: int) -> List[Tuple[int, int]]:
    return [shard_bounds(n_clips, r, world) for r in range(world)]


def max_over_ranks(value: float, device="cpu") -> float:
    """MAX all-reduce of a scalar (timings are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_waveforms(local: torch.Tensor, n_clips: int):
    """Optional: assemble the (n_clips, 1, L) result on rank 0 from per-rank slabs (ragged slabs padded)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = all_bounds(n_clips, world)
    width = max(hi - lo for lo, hi in bounds)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, parts, dst=0)
    if rank != 0:
        return None
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, bounds)], dim=0)
