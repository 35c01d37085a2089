"""Vocab-parallel mode of the fused LM-head KD loss (SURVEY.md 8e "optional mode", BASELINE configs[4]).

Rank g holds ``W[v0_g:v1_g, :]`` and the matching teacher-logit columns (or the whole top-k cache: indices are
global); every rank sees the same hidden states and labels.  Per step:

  forward : kd_fused_linear_fwd_partial on the slice -> one 12-float record per row
            all-gather of the records ([G][R][12], 1.5 MB at R = 4096, G = 8)
            kd_fused_merge_ranks (split-V merge rule, appendix C) -> identical losses + row_stats on every rank
  backward: kd_fused_linear_bwd_range on the slice (labels / indices shifted by v_offset inside the kernels)
            dW slice is final and local; dH is a partial sum -> all-reduce(SUM) in fp32, then rounded once

All arithmetic is in libkd_b200.so; this file is plumbing (torch.distributed for the two exchanges).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, dtype_code, require_cuda, stream_ptr
from .loss import (IGNORE_INDEX, _fused_backward, _fused_workspace, _ptr, _teacher_kind, _workspace, alloc_logit_cache,
                   finalize_losses,
                   prepare_rows)

RANK_REC_FLOATS = 12  # kRankRecFloats of csrc/kd_fused.cu


def vocab_slices(V, world, align=256):
    """Contiguous vocabulary slices, one per rank, starting on multiples of ``align`` (TMA needs 16-byte aligned
    slice bases; 256 also keeps the tile grid of every rank identical except the last).  Pure host logic."""
    while True:
        units = -(-V // align)
        base, extra = divmod(units, world)
        if base > 0 or align <= 8:
            break
        align //= 2
    out, v0 = [], 0
    for g in range(world):
        v1 = min(v0 + (base + (1 if g < extra else 0)) * align, V)
        out.append((v0, v1))
        v0 = v1
    return out


def forward_partial(h2, W_slice, y_slice, topk, row_target, v_offset, tau, v_chunk=0, cache=None):
    """kd_fused_linear_fwd_partial: the slice's per-row record [R, 12] (fp32)."""
    lib = _lib.load()
    R, H = h2.shape
    V = W_slice.shape[0]
    dev = h2.device
    teacher_kind = _teacher_kind(y_slice, topk)
    K = topk[0].size(-1) if teacher_kind == _lib.KD_TEACHER_SPARSE else 0
    rec = torch.empty((R, RANK_REC_FLOATS), dtype=torch.float32, device=dev)
    ws = _fused_workspace(R, H, V, v_chunk, dev, K)
    rc = lib.kd_fused_linear_fwd_partial(
        h2.data_ptr(), h2.stride(0), W_slice.data_ptr(), W_slice.stride(0), teacher_kind,
        _ptr(y_slice), dtype_code(y_slice.dtype) if y_slice is not None else 0,
        y_slice.stride(0) if y_slice is not None else 0,
        _ptr(topk[0]) if K else 0, _ptr(topk[1]) if K else 0, K, row_target.data_ptr(), 0, R, H, V, int(v_offset),
        float(tau), rec.data_ptr(), _ptr(cache), cache.numel() if cache is not None else 0, ws.data_ptr(), ws.numel(),
        stream_ptr(dev))
    check(rc, "kd_fused_linear_fwd_partial")
    return rec, ws


def merge_ranks(recs, row_target, teacher_kind, tau):
    """kd_fused_merge_ranks: recs [G, R, 12] -> (sums[8], row_stats[R, 4])."""
    lib = _lib.load()
    G, R, _ = recs.shape
    dev = recs.device
    recs = recs.contiguous()
    sums = torch.empty(8, dtype=torch.float32, device=dev)
    row_stats = torch.empty((R, 4), dtype=torch.float32, device=dev)
    ws = _workspace(lib.kd_fused_merge_workspace_bytes(), dev)
    check(lib.kd_fused_merge_ranks(recs.data_ptr(), G, row_target.data_ptr(), R, teacher_kind, float(tau),
                                   sums.data_ptr(), row_stats.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)),
          "kd_fused_merge_ranks")
    return sums, row_stats


def _default_gather(group):
    def gather(rec):
        world = dist.get_world_size(group)
        out = torch.empty((world,) + tuple(rec.shape), dtype=rec.dtype, device=rec.device)
        dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
        return out

    return gather


def _default_reduce(group):
    def reduce(x):
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        return x

    return reduce


class _KDVocabParallel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, W_slice, y_slice, topk_v, topk_i, row_target, n_valid, v_offset, tau, alpha, v_chunk,
                gather_fn, reduce_fn):
        topk = (topk_v, topk_i) if topk_v is not None else None
        teacher_kind = _teacher_kind(y_slice, topk)
        cache = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            cache = alloc_logit_cache(h.shape[0], W_slice.shape[0], v_chunk, h.device)
        rec, ws = forward_partial(h, W_slice, y_slice, topk, row_target, v_offset, tau, v_chunk, cache)
        sums, row_stats = merge_ranks(gather_fn(rec), row_target, teacher_kind, tau)
        eff_alpha = alpha if teacher_kind != _lib.KD_TEACHER_NONE else 1.0
        losses = finalize_losses(sums, tau, eff_alpha, teacher_kind == _lib.KD_TEACHER_SPARSE)
        ctx.set_materialize_grads(False)
        ctx.cfg = (tau, eff_alpha, teacher_kind, int(v_offset), int(v_chunk), reduce_fn)
        ctx.save_for_backward(h, W_slice, y_slice, row_target, row_stats, n_valid, topk_v, topk_i)
        ctx.ws = ws
        ctx.cache = cache
        total, task, distill, teacher = losses.unbind(0)
        ctx.mark_non_differentiable(teacher)
        return total, task, distill, teacher

    @staticmethod
    def backward(ctx, g_total, g_task, g_distill, g_teacher):
        h, W_slice, y_slice, row_target, row_stats, n_valid, topk_v, topk_i = ctx.saved_tensors
        topk = (topk_v, topk_i) if topk_v is not None else None
        tau, alpha, teacher_kind, v_offset, v_chunk, reduce_fn = ctx.cfg
        zero = torch.zeros((), dtype=torch.float32, device=h.device)
        gt = zero if g_total is None else g_total.detach().float()
        w_ce = gt * alpha + (zero if g_task is None else g_task.detach().float())
        w_kl = gt * (1.0 - alpha) + (zero if g_distill is None else g_distill.detach().float())
        coef = torch.stack([w_ce.reshape(()), w_kl.reshape(())]).contiguous()
        need_h, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dH32, dW = _fused_backward(h, W_slice, y_slice, row_target, row_stats, n_valid, coef, tau, teacher_kind, 0,
                                   v_chunk, torch.bfloat16, need_h, need_w, ctx.ws, topk, None, v_offset, dh_fp32=True,
                                   cache=ctx.cache)
        dH = None
        if need_h:
            dH = reduce_fn(dH32).to(h.dtype)  # partial sums over vocabulary slices -> one rounding
        return (dH, dW) + (None,) * 11


def fused_linear_kd_loss_vocab_parallel(hidden, weight_slice, labels, v_offset, teacher_logits_slice=None,
                                        teacher_top_k_v=None, teacher_top_k_i=None, speech_token_mask=None,
                                        temperature=2.0, alpha=0.5, ignore_index=IGNORE_INDEX, v_chunk=0, group=None,
                                        gather_fn=None, reduce_fn=None):
    """``DistillationLoss`` on ``hidden @ W.T`` with W (and dense teacher columns) sharded over the vocabulary.

    hidden [B,T,H] bf16 (identical on every rank), weight_slice [V_g,H] bf16 = W[v_offset : v_offset + V_g],
    teacher_logits_slice [B,T,V_g] (same columns) or the global top-k cache, labels [B,T] (global ids).
    Returns the reference's 4-tuple (fp32), identical on every rank; ``total.backward()`` gives the full dH
    (all-reduced) and this rank's dW slice.  ``gather_fn`` / ``reduce_fn`` default to NCCL over ``group``.
    """
    require_cuda(hidden, weight_slice)
    if hidden.dtype != torch.bfloat16 or weight_slice.dtype != torch.bfloat16:
        raise TypeError("vocab-parallel KD computes in bf16 with fp32 accumulation: pass bf16 hidden and weight")
    if hidden.dim() != 3:
        raise ValueError("hidden must be [B, T, H]")
    B, T, H = hidden.shape
    Vg = weight_slice.shape[0]
    dev = hidden.device
    h2 = hidden.reshape(B * T, H)
    if h2.stride(-1) != 1:
        h2 = h2.contiguous()
    W = weight_slice if weight_slice.stride(-1) == 1 else weight_slice.contiguous()
    row_target, n_valid = prepare_rows(labels, speech_token_mask, B, T, ignore_index, dev)
    y = topk_v = topk_i = None
    if teacher_logits_slice is not None:
        y = teacher_logits_slice.detach()
        if y.shape[-1] != Vg or y.numel() != B * T * Vg:
            raise ValueError(f"teacher_logits_slice shape {tuple(y.shape)} does not match [B={B}, T={T}, V_g={Vg}]")
        y = y.reshape(B * T, Vg)  # a column slice of a [B,T,V] tensor stays a view (row stride V)
        if y.stride(-1) != 1:
            y = y.contiguous()
        if y.dtype == torch.float16:
            y = y.float()
    elif teacher_top_k_v is not None and teacher_top_k_i is not None:
        K = teacher_top_k_v.size(-1)
        topk_v = teacher_top_k_v.detach().to(device=dev, dtype=torch.float32).reshape(B * T, K).contiguous()
        topk_i = teacher_top_k_i.detach().to(device=dev, dtype=torch.int32).reshape(B * T, K).contiguous()
    if gather_fn is None:
        gather_fn = _default_gather(group)
    if reduce_fn is None:
        reduce_fn = _default_reduce(group)
    return _KDVocabParallel.apply(h2, W, y, topk_v, topk_i, row_target, n_valid, int(v_offset), float(temperature),
                                  float(alpha), int(v_chunk), gather_fn, reduce_fn)
