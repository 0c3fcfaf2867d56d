"""Channel sharding across ranks (one process per GPU). Channels are independent
(no cross-channel term anywhere in src/main.cpp:1232-1308 of the reference), so ranks never
exchange samples: the only collective is the reduction of the timing scalar."""
from __future__ import annotations


def channels_of_rank(rank: int, world: int, channels_per_rank: int) -> range:
    """Weak scaling: every rank owns `channels_per_rank` consecutive global channel ids."""
    if not (0 <= rank < world) or channels_per_rank < 0:
        raise ValueError("bad rank / world / channels_per_rank")
    return range(rank * channels_per_rank, (rank + 1) * channels_per_rank)


def split_total(total_channels: int, world: int) -> list[range]:
    """Strong scaling: a fixed set of channels dealt out in contiguous, near-equal shares."""
    if world < 1 or total_channels < 0:
        raise ValueError("bad world / total_channels")
    base, extra = divmod(total_channels, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append(range(start, start + n))
        start += n
    return out


def max_over_ranks(value: float, device=None) -> float:
    """MAX-reduce a scalar over the default process group (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    """SUM-reduce a scalar over the default process group (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(samples_per_rank_per_step: int, steps: int, world: int, max_ms: float) -> float:
    """Whole-job MS/s: all ranks' samples over the slowest rank's device time."""
    return world * samples_per_rank_per_step * steps / (max_ms * 1e-3) / 1e6
