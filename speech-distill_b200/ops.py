"""``torch.library`` registration of the path's two loss entry points (north_star: "a PyTorch custom op over a thin
C-ABI extension").

The default Python mirror (``loss.py``) drives the C ABI from ``torch.autograd.Function`` nodes, which eager training
loops such as ``train.py`` / ``stage1.py`` need nothing more than; those nodes are opaque to ``torch.compile``.  The
ops below wrap the same C calls as ``speech_distill_b200::kd_loss`` (K2, loss on materialised logits,
``distillation_loss.py:14-128``) and ``speech_distill_b200::fused_linear_kd`` / ``..._bwd`` (K1, LM head + loss
without logits) with fake (meta) implementations and registered autograd formulas, so a compiled model traces through
the loss without a graph break (``tests/test_gpu_ops.py`` compiles with ``fullgraph=True``).

Differences to the ``autograd.Function`` path, both by construction of an op with a fixed schema:
* only ``total`` is differentiable (what every caller in the reference back-propagates, ``train.py:97-116``); ``task``,
  ``distill`` and ``teacher_task`` come back detached, as the logged metrics they are;
* the data-parallel hooks (``GradSync`` ranges, row compaction) stay on the ``autograd.Function`` path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from . import loss as L
from ._lib import require_cuda
from .loss import IGNORE_INDEX

_NS = "speech_distill_b200"


# ---------------------------------------------------------------------------------------------------------------------
# K2: loss on materialised logits
# ---------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{_NS}::kd_loss", mutates_args=(), device_types="cuda")
def _kd_loss(student_logits: Tensor, labels: Tensor, teacher_logits: Optional[Tensor], teacher_top_k_v: Optional[Tensor],
             teacher_top_k_i: Optional[Tensor], speech_token_mask: Optional[Tensor], temperature: float, alpha: float,
             ignore_index: int) -> Tuple[Tensor, Tensor]:
    """-> (losses fp32 [4] = total, task, distill, teacher_task; d total / d student_logits [B,T,V])."""
    z = L._as_btv(student_logits, "student_logits")
    B, T, V = z.shape
    dev = z.device
    row_target, n_valid = L.prepare_rows(labels, speech_token_mask, B, T, ignore_index, dev)
    y = v = i = None
    if teacher_logits is not None:
        y = L._as_btv(teacher_logits, "teacher_logits")
    else:
        K = teacher_top_k_v.size(-1)
        v = teacher_top_k_v.to(device=dev, dtype=torch.float32).reshape(B, T, K).contiguous()
        i = teacher_top_k_i.to(device=dev, dtype=torch.int32).reshape(B, T, K).contiguous()
    sums, dlogits = L._stream_call(z, y, v, i, row_target, n_valid, temperature, alpha, 1.0, True)
    return L.finalize_losses(sums, temperature, alpha, y is None), dlogits


@_kd_loss.register_fake
def _(student_logits, labels, teacher_logits, teacher_top_k_v, teacher_top_k_i, speech_token_mask, temperature, alpha,
      ignore_index):
    B, T, V = student_logits.shape
    return (student_logits.new_empty((4,), dtype=torch.float32),
            student_logits.new_empty((B, T, V)))


def _kd_loss_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])


def _kd_loss_backward(ctx, g_losses, g_dlogits):
    (dlogits,) = ctx.saved_tensors
    grad = dlogits * g_losses[0].to(dlogits.dtype)  # total is the differentiable output (see the module docstring)
    return grad, None, None, None, None, None, None, None, None


_kd_loss.register_autograd(_kd_loss_backward, setup_context=_kd_loss_setup)


def kd_loss(student_logits, labels, teacher_logits=None, teacher_top_k_v=None, teacher_top_k_i=None,
            speech_token_mask=None, temperature=2.0, alpha=0.5, ignore_index=IGNORE_INDEX):
    """Reference forward (distillation_loss.py:14-128) as a registered op: ``(total, task, distill, teacher_task)``."""
    if teacher_logits is None and (teacher_top_k_v is None or teacher_top_k_i is None):
        raise ValueError("Either teacher_logits or top_k must be provided")  # distillation_loss.py:120
    if student_logits.dim() != 3:
        raise ValueError("student_logits must be [B, T, V]")
    if teacher_logits is not None:
        teacher_logits = teacher_logits.detach()
        teacher_top_k_v = teacher_top_k_i = None  # dense wins when both are given (:56 before :73)
    else:
        teacher_top_k_v, teacher_top_k_i = teacher_top_k_v.detach(), teacher_top_k_i.detach()
    losses, _ = _kd_loss(student_logits, labels, teacher_logits, teacher_top_k_v, teacher_top_k_i, speech_token_mask,
                         float(temperature), float(alpha), int(ignore_index))
    rest = losses.detach()
    return losses[0], rest[1], rest[2], rest[3]


# ---------------------------------------------------------------------------------------------------------------------
# K1: LM head + loss without logits
# ---------------------------------------------------------------------------------------------------------------------
def _teacher_args(B, T, dev, teacher_logits, teacher_top_k_v, teacher_top_k_i):
    y = topk = None
    if teacher_logits is not None:
        y = teacher_logits.reshape(B * T, teacher_logits.size(-1))
        if y.stride(-1) != 1:
            y = y.contiguous()
    elif teacher_top_k_v is not None:
        K = teacher_top_k_v.size(-1)
        topk = (teacher_top_k_v.to(device=dev, dtype=torch.float32).reshape(B * T, K).contiguous(),
                teacher_top_k_i.to(device=dev, dtype=torch.int32).reshape(B * T, K).contiguous())
    return y, topk


@torch.library.custom_op(f"{_NS}::fused_linear_kd", mutates_args=(), device_types="cuda")
def _fused_linear_kd(hidden: Tensor, weight: Tensor, labels: Tensor, teacher_logits: Optional[Tensor],
                     teacher_top_k_v: Optional[Tensor], teacher_top_k_i: Optional[Tensor],
                     speech_token_mask: Optional[Tensor], temperature: float, alpha: float, ignore_index: int,
                     logit_cache_mb: float, old_vocab_size: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """hidden [B,T,H] bf16, weight [V,H] bf16 -> (losses [4], row_stats [R,4], row_target [R] int32, n_valid [1] int32,
    logit cache uint8 [n]); everything the backward op needs travels as op outputs.  ``old_vocab_size`` is only
    handed on to the backward (stage1's frozen rows of dW)."""
    B, T, H = hidden.shape
    dev = hidden.device
    h2 = hidden.reshape(B * T, H)
    if h2.stride(-1) != 1:
        h2 = h2.contiguous()
    W = weight if weight.stride(-1) == 1 else weight.contiguous()
    row_target, n_valid = L.prepare_rows(labels, speech_token_mask, B, T, ignore_index, dev)
    y, topk = _teacher_args(B, T, dev, teacher_logits, teacher_top_k_v, teacher_top_k_i)
    cache = L.alloc_logit_cache(B * T, W.shape[0], 0, dev, logit_cache_mb)
    sums, row_stats, _ = L._fused_forward(h2, W, y, row_target, temperature, alpha, 0, topk, None, cache)
    kind = L._teacher_kind(y, topk)
    eff_alpha = alpha if kind != _lib.KD_TEACHER_NONE else 1.0
    losses = L.finalize_losses(sums, temperature, eff_alpha, kind == _lib.KD_TEACHER_SPARSE)
    if cache is None:
        cache = torch.empty(0, dtype=torch.uint8, device=dev)
    return losses, row_stats, row_target, n_valid, cache


@_fused_linear_kd.register_fake
def _(hidden, weight, labels, teacher_logits, teacher_top_k_v, teacher_top_k_i, speech_token_mask, temperature, alpha,
      ignore_index, logit_cache_mb, old_vocab_size):
    B, T, H = hidden.shape
    R, V = B * T, weight.shape[0]
    nbytes = int(_lib.load().kd_fused_logit_cache_bytes(int(R), int(V), 0, L.logit_cache_budget(logit_cache_mb)))
    return (hidden.new_empty((4,), dtype=torch.float32), hidden.new_empty((R, 4), dtype=torch.float32),
            hidden.new_empty((R,), dtype=torch.int32), hidden.new_empty((1,), dtype=torch.int32),
            hidden.new_empty((nbytes,), dtype=torch.uint8))


@torch.library.custom_op(f"{_NS}::fused_linear_kd_bwd", mutates_args=(), device_types="cuda")
def _fused_linear_kd_bwd(hidden: Tensor, weight: Tensor, teacher_logits: Optional[Tensor], teacher_top_k_v: Optional[Tensor],
                         teacher_top_k_i: Optional[Tensor], row_stats: Tensor, row_target: Tensor, n_valid: Tensor,
                         cache: Tensor, g_total: Tensor, temperature: float, alpha: float,
                         old_vocab_size: int) -> Tuple[Tensor, Tensor]:
    """-> (d total / d hidden [B,T,H], d total / d weight [V,H]) x g_total, both bf16; rows below ``old_vocab_size``
    of dW are never computed and stay zero (stage1.py:29-73)."""
    B, T, H = hidden.shape
    dev = hidden.device
    h2 = hidden.reshape(B * T, H)
    if h2.stride(-1) != 1:
        h2 = h2.contiguous()
    W = weight if weight.stride(-1) == 1 else weight.contiguous()
    y, topk = _teacher_args(B, T, dev, teacher_logits, teacher_top_k_v, teacher_top_k_i)
    kind = L._teacher_kind(y, topk)
    eff_alpha = alpha if kind != _lib.KD_TEACHER_NONE else 1.0
    g = g_total.detach().to(torch.float32).reshape(())
    coef = torch.stack([g * eff_alpha, g * (1.0 - eff_alpha)]).contiguous()
    dH, dW = L._fused_backward(h2, W, y, row_target, row_stats, n_valid, coef, temperature, kind, int(old_vocab_size), 0,
                               torch.bfloat16, True, True, None, topk, None,
                               cache=cache if cache.numel() > 0 else None)
    return dH.reshape(B, T, H), dW


@_fused_linear_kd_bwd.register_fake
def _(hidden, weight, teacher_logits, teacher_top_k_v, teacher_top_k_i, row_stats, row_target, n_valid, cache, g_total,
      temperature, alpha, old_vocab_size):
    return hidden.new_empty(hidden.shape, dtype=torch.bfloat16), weight.new_empty(weight.shape, dtype=torch.bfloat16)


def _fused_setup(ctx, inputs, output):
    (hidden, weight, labels, teacher_logits, topk_v, topk_i, mask, temperature, alpha, ignore_index, cache_mb,
     old_vocab_size) = inputs
    losses, row_stats, row_target, n_valid, cache = output
    ctx.save_for_backward(hidden, weight, teacher_logits, topk_v, topk_i, row_stats, row_target, n_valid, cache)
    ctx.cfg = (temperature, alpha, old_vocab_size)


def _fused_backward_formula(ctx, g_losses, g_row_stats, g_row_target, g_n_valid, g_cache):
    hidden, weight, teacher_logits, topk_v, topk_i, row_stats, row_target, n_valid, cache = ctx.saved_tensors
    temperature, alpha, old_vocab_size = ctx.cfg
    dH, dW = _fused_linear_kd_bwd(hidden, weight, teacher_logits, topk_v, topk_i, row_stats, row_target, n_valid, cache,
                                  g_losses[0], temperature, alpha, old_vocab_size)
    return dH, dW, None, None, None, None, None, None, None, None, None, None


_fused_linear_kd.register_autograd(_fused_backward_formula, setup_context=_fused_setup)


def fused_linear_kd_loss(hidden, lm_head_weight, labels, teacher_logits=None, teacher_top_k_v=None, teacher_top_k_i=None,
                         speech_token_mask=None, temperature=2.0, alpha=0.5, ignore_index=IGNORE_INDEX,
                         logit_cache_mb=None, old_vocab_size=0):
    """LM head + KD loss without logits as registered ops: ``(total, task, distill, teacher_task)``; hidden
    ``[B,T,H]`` and lm_head_weight ``[V,H]`` in bf16.  With no teacher at all the loss is the plain cross-entropy
    (``total == task``), as in ``loss.fused_linear_kd_loss``; ``old_vocab_size`` > 0 leaves the dW rows of the old
    vocabulary uncomputed and zero (stage1.py:29-73)."""
    require_cuda(hidden, lm_head_weight)
    if hidden.dim() != 3:
        raise ValueError("hidden must be [B, T, H]")
    if hidden.dtype != torch.bfloat16 or lm_head_weight.dtype != torch.bfloat16:
        raise TypeError("the op form takes bf16 hidden states and weights (use loss.fused_linear_kd_loss for fp32)")
    if teacher_logits is not None:
        teacher_logits = teacher_logits.detach()
        teacher_top_k_v = teacher_top_k_i = None
    elif teacher_top_k_v is not None:
        teacher_top_k_v, teacher_top_k_i = teacher_top_k_v.detach(), teacher_top_k_i.detach()
    mb = float(L.logit_cache_budget(logit_cache_mb)) / float(1 << 20)
    losses = _fused_linear_kd(hidden, lm_head_weight, labels, teacher_logits, teacher_top_k_v, teacher_top_k_i,
                              speech_token_mask, float(temperature), float(alpha), int(ignore_index), mb,
                              int(old_vocab_size))[0]
    rest = losses.detach()
    return losses[0], rest[1], rest[2], rest[3]
