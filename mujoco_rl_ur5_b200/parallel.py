"""Multi-GPU plumbing: environments are independent, so N_total envs are split into contiguous index ranges, one process
(rank) per GPU, with NO collective on the physics / render path.  Seeds are functions of the GLOBAL env index,
so per-env results do not depend on the number of ranks.  The only collective is an all-gather of the small per-env result
record (reward u8, result i32, steps i32) for reporting; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n_total: int, rank: int, world: int):
    """[start, stop) of the env indices owned by `rank`; ranges are contiguous, disjoint and cover [0, n_total)."""
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def env_seed(global_env_index: int, episode: int = 0, base: int = 20000) -> int:
    """Reset seed of one environment (RandomState(20000 + i)); independent of the sharding."""
    return int(base + global_env_index + 100003 * episode)


def action_seed(global_env_index: int, base: int = 30000) -> int:
    return int(base + global_env_index)


def gather_results(reward, result, steps, n_total, group=None):
    """All-gather the per-env result record of every rank into global-index order.  Inputs are this rank's 1-D tensors
    (any device supported by the process group's backend).  Returns (reward, result, steps) of length n_total on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return reward, result, steps
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    cap = max(b - a for a, b in sizes)
    n_local = sizes[rank][1] - sizes[rank][0]
    rec = torch.zeros((cap, 3), dtype=torch.int32, device=reward.device)
    rec[:n_local, 0] = reward.to(torch.int32)
    rec[:n_local, 1] = result.to(torch.int32)
    rec[:n_local, 2] = steps.to(torch.int32)
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec, group=group)
    full = torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)
    return full[:, 0].to(torch.uint8), full[:, 1], full[:, 2]


def shard_actions(actions: np.ndarray, rank: int, world: int):
    """Slice a global [N_total, 2] action array to this rank's envs."""
    a, b = shard_range(len(actions), rank, world)
    return actions[a:b]
