"""Host-side frame sharding protocol (backend-agnostic: NCCL on the GPUs, gloo in the CPU
tests).  Frames are the data-parallel axis: rank r owns the contiguous block
``frame_range(T, r, R)``; the ranks exchange exactly two messages per pass --

  1. ``global_mean``  : all-reduce(sum) of the per-column sums (3N doubles) -> the global
                        mean every rank centres with;
  2. ``reduce_sums``  : ONE all-reduce(sum) of the contiguous ``[G || D]`` buffer (partial
                        Gram and distance sums, 2 * Npad^2 doubles) before the epilogue.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def frame_range(n_frames, rank, world):
    """Contiguous, balanced frame block of `rank` (sizes differ by at most one)."""
    return n_frames * rank // world, n_frames * (rank + 1) // world


def is_distributed(group=None):
    """group=False: this pipeline is local even inside an initialised process group (used for the
    single-rank recomputation that multi-rank results are checked against)."""
    if group is False:
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def global_mean(colsum, mean_out, n_frames_total, group=None):
    """colsum: this rank's column sums (tensor, overwritten with the global sums);
    mean_out <- global sums / total frames."""
    if is_distributed(group):
        dist.all_reduce(colsum, op=dist.ReduceOp.SUM, group=group)
    torch.div(colsum, float(n_frames_total), out=mean_out)
    return mean_out


def reduce_sums(sums, group=None):
    """sums: contiguous [G || D] partial-sum tensor, reduced in place over the ranks."""
    if is_distributed(group):
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums
