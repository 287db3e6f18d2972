"""Multi-GPU plumbing: scenarios are independent, so a batch is split into contiguous
per-rank shards with no per-tick traffic; the only collective is one all-reduce (sum) of the int64
episode statistics — the distributed form of AgentManager.get_collision_rate / get_offroad_rate
(reference: lcsim/entity/agent/manager.py:329-345; the reference's own sharding of scenarios over Ray
workers is lcsim/envs/gym_warpper.py:35-45)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_scenarios: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [s0, s1) of rank `rank`; sizes differ by at most one, earlier ranks get the extras."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(int(n_scenarios), world_size)
    s0 = rank * base + min(rank, extra)
    return s0, s0 + base + (1 if rank < extra else 0)


def reduce_episode_stats(stats):
    """Sum of the (8,) int64 statistics over all ranks (NCCL over NVLink on GPUs, gloo on CPU).
    Returns the input unchanged when torch.distributed is not initialised."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        stats = stats.clone()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def rates(stats) -> dict:
    """collision / off-road rates per agent-tick from the reduced statistics (manager.py:329-345 pooled)."""
    s = [int(x) for x in stats.tolist()]
    active = max(s[1], 1)
    return {"agent_steps": s[0], "collision_rate": s[2] / active, "offroad_rate": s[3] / active,
            "arrivals": s[4], "departures": s[5], "ticks": s[6]}
