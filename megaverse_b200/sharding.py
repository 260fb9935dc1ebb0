"""Multi-GPU layout of the path: envs are independent (own RNG, own colliders, own scene), so the batch is sharded by
contiguous env ranges, one process per GPU, with NO data-path collective (the reference is process-per-GPU as
well).  torch.distributed is only used for the barrier and the max-over-ranks of the device time.

The functions here are pure host logic (usable with the gloo backend on CPU); bench.py and the tests share them."""
from __future__ import annotations


def shard_range(total_envs: int, world: int, rank: int) -> tuple[int, int]:
    """[begin, end) of the global env indices owned by `rank`; contiguous, sizes differ by at most one"""
    if world < 1 or not 0 <= rank < world or total_envs < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total_envs, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def env_seeds(begin: int, end: int, base_seed: int = 42) -> list[int]:
    """seed of global env i is base_seed + i (megaverse_test_app.cpp:250-254), whichever rank owns it"""
    return [base_seed + i for i in range(begin, end)]


def view_offset(begin: int, agents_per_env: int) -> int:
    """first row of this rank's observations in the concatenated [N, H, W, 4] tensor (agent-views are env-major)"""
    return begin * agents_per_env


def aggregate_throughput(local_units: int, local_ms: float, dist=None) -> tuple[float, float, int]:
    """whole-job units per second = sum of units over ranks / max device time over ranks.
    `dist` is torch.distributed (initialised) or None for a single process.  Returns (units_per_s, max_ms, total_units)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_units / (local_ms / 1e3), local_ms, local_units
    import torch

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    u = torch.tensor([local_units], dtype=torch.int64, device=dev)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    max_ms, total = float(t.item()), int(u.item())
    return total / (max_ms / 1e3), max_ms, total
