"""Multi-GPU plumbing: one process per GPU, environments sharded by contiguous
blocks, and the only collective of the path -- an all-reduce of the episode
statistics vector.

Replaces ``guard_utils/mpi_tools.py`` of the reference (``mpi_statistics_scalar``
:70-92, ``mpi_sum`` :63-68: CPU MPI ``Allreduce`` on numpy scalars) with
``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests).  Envs are
independent (``jax.vmap``, engine.py:247-258), so there is no collective in the
step itself; ``global_env_id = offset + local_id`` keys every Philox draw, which
makes results independent of the number of ranks.
"""
from __future__ import annotations

import math

import torch

# layout of the statistics vector (gx_stats_read)
N_EPISODES, SUM_RET, SUM_COST, SUM_LEN, SUM_RET2, SUM_COST2, TOTAL_COST, TOTAL_STEPS = range(8)


def rank() -> int:
    """proc_id() (mpi_tools.py:38-40); 0 without a process group."""
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(total_envs: int, rank: int, world: int) -> tuple[int, int]:
    """(local env count, global id of local env 0) for a contiguous block split."""
    assert 0 <= rank < world and total_envs >= world
    base, rem = divmod(total_envs, world)
    n = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return n, offset


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the 8-double statistics vector over all ranks (in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def episode_summary(stats) -> dict:
    """What the reference logs per epoch (ppo_train.py:405-411, logx.py:366):
    mean/std of EpRet and EpCost, mean EpLen, cumulative cost and cost rate."""
    s = [float(v) for v in stats]
    n = max(s[N_EPISODES], 1.0)
    mean_ret, mean_cost = s[SUM_RET] / n, s[SUM_COST] / n
    var = lambda sq, m: max(sq / n - m * m, 0.0)
    return {
        'Episodes': s[N_EPISODES],
        'EpRet': mean_ret, 'StdEpRet': math.sqrt(var(s[SUM_RET2], mean_ret)),
        'EpCost': mean_cost, 'StdEpCost': math.sqrt(var(s[SUM_COST2], mean_cost)),
        'EpLen': s[SUM_LEN] / n,
        'CumulativeCost': s[TOTAL_COST],
        'CostRate': s[TOTAL_COST] / max(s[TOTAL_STEPS], 1.0),
        'TotalEnvInteracts': s[TOTAL_STEPS],
    }


def global_mean(x: torch.Tensor, group=None) -> torch.Tensor:
    """mpi_avg (mpi_tools.py:56-58) of a scalar tensor, on the device."""
    import torch.distributed as dist
    x = x.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        x /= dist.get_world_size(group)
    return x


def sync_params(module: torch.nn.Module, group=None) -> None:
    """sync_params (mpi_pytorch.py:29-35): broadcast rank 0's parameters."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in module.parameters():
            dist.broadcast(p.data, src=0, group=group)
