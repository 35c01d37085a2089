"""Token-shard data parallelism for the KD hot path.

Every rank holds the full LM-head weight and its own sequences.  The only exchanges are
  (1) all-reduce(SUM) of the valid-row count  -> global N used inside the gradient kernels,
  (2) all-reduce(SUM) of the 8-float sums record -> identical loss scalars on every rank,
  (3) all-reduce(SUM) of dW (done by the caller / DDP; helper below).
With the global N the result equals the single-process reference evaluated on the concatenated
batch (HF-DDP would average per-rank means instead, which differs when valid counts differ).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def make_reduce_fns(group=None):
    """Returns (reduce_fn, count_reduce_fn) for kd_loss_on_logits / fused_linear_kd_loss."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None, None

    def reduce_fn(sums):
        out = sums.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    def count_reduce_fn(n_valid):
        out = n_valid.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    return reduce_fn, count_reduce_fn


def losses_from_sums(sums, tau, alpha, sparse=False):
    """Host/torch mirror of kd_finalize_losses (distillation_loss.py:68,116-118,123,126) for logging
    and for the CPU tests of the reduction logic; works on any device."""
    n = sums[3]
    if float(n) <= 0:
        z = torch.zeros((), dtype=sums.dtype)
        return z, z.clone(), z.clone(), z.clone()
    task = sums[0] / n
    distill = (tau * tau) * sums[1] / n
    if sparse:
        teacher = -sums[2] / sums[4] if float(sums[4]) > 0 else torch.zeros((), dtype=sums.dtype)
    else:
        teacher = sums[2] / n
    return alpha * task + (1 - alpha) * distill, task, distill, teacher


def allreduce_grad_(grad, group=None, bucket_rows=0):
    """In-place SUM all-reduce of a [V, H] LM-head gradient; ``bucket_rows`` > 0 splits it into row
    blocks so that NCCL can start on finished vocabulary chunks while later ones are computed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grad
    if bucket_rows <= 0:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
        return grad
    handles = []
    for r0 in range(0, grad.size(0), bucket_rows):
        handles.append(dist.all_reduce(grad[r0 : r0 + bucket_rows], op=dist.ReduceOp.SUM, group=group, async_op=True))
    for h in handles:
        h.wait()
    return grad
