"""Multi-GPU plumbing: one process per GPU, the trajectory batch sharded contiguously, and ONE
all-reduce of (sum loss, sum dp, count) per outer iteration.  Trajectories are
independent (reference PDP/PDP.py:1266 loop body, Examples/IRL/quadrotor/uav_PDP.py:45-75), so there
is no data-path collective; the average replicates ``dp / n_batch`` (PDP.py:1293-1294) over the
GLOBAL batch.  Works with the ``nccl`` backend on CUDA tensors and with ``gloo`` on CPU tensors."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int):
    """Contiguous split [lo, hi) of ``global_batch`` trajectories; the first ``global_batch % world``
    ranks take one extra trajectory."""
    base, extra = divmod(int(global_batch), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def batch_sums(loss_dp_local: torch.Tensor) -> torch.Tensor:
    """loss_dp_local[B_local, r+1] -> partial[r+2] = (sum loss, sum dp, B_local).  CUDA tensors go through the C ABI's
    deterministic reduction kernel (``pdp_reduce_loss_dp``: one launch, nothing eager between the sweep kernel and the
    collective); CPU tensors (the gloo tests of the host logic) use torch."""
    if loss_dp_local.is_cuda:
        from . import engine
        return engine.reduce_loss_dp(loss_dp_local.contiguous())
    count = torch.full((1,), float(loss_dp_local.shape[0]), dtype=loss_dp_local.dtype, device=loss_dp_local.device)
    return torch.cat([loss_dp_local.sum(dim=0), count])


def all_reduce_sums(partial: torch.Tensor, group=None) -> torch.Tensor:
    """The single collective of an outer iteration: SUM all-reduce of r+2 float64 (in place); identity when the process
    group is not initialised or has one rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def reduce_loss_dp(loss_dp_local: torch.Tensor, group=None):
    """loss_dp_local[B_local, r+1] (per-trajectory loss and half-gradient) -> (mean loss, mean dp[r])
    over the global batch.  One reduction kernel + one all-reduce of r+2 float64 values."""
    partial = all_reduce_sums(batch_sums(loss_dp_local), group)
    count = partial[-1]
    return partial[0] / count, partial[1:-1] / count
