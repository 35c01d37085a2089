"""Host-side mirror of the reference loss interface over the C ABI (include/kd_b200.h).

``DistillationLoss`` keeps the reference's constructor, ``forward`` signature, 4-tuple return and
error behaviour (reference ``distillation_loss.py:6-128``); ``fused_linear_kd_loss`` is the opt-in
form that takes hidden states + LM-head weight instead of logits.  Everything here
is plumbing: row bookkeeping, dtype/stride checks, autograd registration.  All arithmetic runs in
libkd_b200.so on the current CUDA stream; there is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import KdError, check, dtype_code, require_cuda, stream_ptr

IGNORE_INDEX = -100


# --------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------
def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _as_btv(x, name):
    """View logits-like input as [B, T, V] with a unit last stride (copy only if unavoidable)."""
    if x.dim() < 2:
        raise ValueError(f"{name} must have at least 2 dimensions [..., T, V]")
    if x.dim() == 2:
        x = x.unsqueeze(0)
    elif x.dim() > 3:
        x = x.reshape(-1, x.size(-2), x.size(-1))
    if x.stride(-1) != 1:
        x = x.contiguous()
    return x


def prepare_rows(labels, speech_token_mask, B, T, ignore_index, device):
    """kd_prepare_rows: row_target int32 [B*T] (-1 = row not scored) and n_valid int32 [1]."""
    lib = _lib.load()
    labels = labels.to(device=device, dtype=torch.int64).reshape(B, T).contiguous()
    mask_u8 = None
    if speech_token_mask is not None:
        mask_u8 = (speech_token_mask.to(device).reshape(B, T) != 0).to(torch.uint8).contiguous()
    row_target = torch.empty(B * T, dtype=torch.int32, device=device)
    n_valid = torch.empty(1, dtype=torch.int32, device=device)
    check(
        lib.kd_prepare_rows(labels.data_ptr(), _ptr(mask_u8), B, T, int(ignore_index), row_target.data_ptr(),
                            n_valid.data_ptr(), stream_ptr(device)),
        "kd_prepare_rows",
    )
    return row_target, n_valid


def finalize_losses(sums, tau, alpha, sparse):
    lib = _lib.load()
    losses = torch.empty(4, dtype=torch.float32, device=sums.device)
    check(lib.kd_finalize_losses(sums.data_ptr(), float(tau), float(alpha), int(bool(sparse)), losses.data_ptr(),
                                 stream_ptr(sums.device)), "kd_finalize_losses")
    return losses


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _stream_call(z, y, topk_v, topk_i, row_target, n_norm, tau, alpha, grad_scale, want_grad):
    """One launch of K2 (dense or sparse).  Returns (sums[8], dlogits or None)."""
    lib = _lib.load()
    B, T, V = z.shape
    dev = z.device
    sums = torch.empty(8, dtype=torch.float32, device=dev)
    dlogits = torch.empty((B, T, V), dtype=z.dtype, device=dev) if want_grad else None
    ws = _workspace(lib.kd_stream_workspace_bytes(), dev)
    if y is not None:
        rc = lib.kd_dense_fwd_bwd(
            z.data_ptr(), dtype_code(z.dtype), z.stride(0), z.stride(1),
            y.data_ptr(), dtype_code(y.dtype), y.stride(0), y.stride(1),
            row_target.data_ptr(), B, T, V, float(tau), float(alpha), n_norm.data_ptr(), float(grad_scale),
            sums.data_ptr(), _ptr(dlogits), ws.data_ptr(), ws.numel(), stream_ptr(dev))
        check(rc, "kd_dense_fwd_bwd")
    else:
        K = topk_v.size(-1)
        rc = lib.kd_sparse_fwd_bwd(
            z.data_ptr(), dtype_code(z.dtype), z.stride(0), z.stride(1),
            topk_v.data_ptr(), topk_i.data_ptr(), K,
            row_target.data_ptr(), B, T, V, float(tau), float(alpha), n_norm.data_ptr(), float(grad_scale),
            sums.data_ptr(), _ptr(dlogits), ws.data_ptr(), ws.numel(), stream_ptr(dev))
        check(rc, "kd_sparse_fwd_bwd")
    return sums, dlogits


# --------------------------------------------------------------------------------------------
# K2 behind autograd: loss on materialised logits
# --------------------------------------------------------------------------------------------
class _KDOnLogits(torch.autograd.Function):
    """(total, task, distill, teacher_task) = f(student_logits); gradient formed in the same sweep."""

    @staticmethod
    def forward(ctx, z, y, topk_v, topk_i, row_target, n_valid, n_norm, tau, alpha, reduce_fn):
        need_grad = bool(ctx.needs_input_grad[0])
        sums, dlogits = _stream_call(z, y, topk_v, topk_i, row_target, n_norm, tau, alpha, 1.0, need_grad)
        if reduce_fn is not None:  # data-parallel: all-reduce the 8-float record before normalising
            sums = reduce_fn(sums)
        losses = finalize_losses(sums, tau, alpha, y is None)
        ctx.set_materialize_grads(False)
        ctx.dlogits = dlogits
        ctx.cfg = (tau, alpha)
        ctx.shape = z.shape
        ctx.save_for_backward(z, y, topk_v, topk_i, row_target, n_norm)
        total, task, distill, teacher = losses.unbind(0)
        ctx.mark_non_differentiable(teacher)
        return total, task, distill, teacher

    @staticmethod
    def backward(ctx, g_total, g_task, g_distill, g_teacher):
        lib = _lib.load()
        z, y, topk_v, topk_i, row_target, n_norm = ctx.saved_tensors
        tau, alpha = ctx.cfg
        grad = None
        if g_total is not None:
            if ctx.dlogits is None:
                raise KdError("backward called twice on the fused KD loss; recompute the loss instead")
            grad = ctx.dlogits
            ctx.dlogits = None  # scaled in place below: single use
            scale = g_total.detach().to(torch.float32).reshape(1).contiguous()
            check(lib.kd_scale_inplace(grad.data_ptr(), dtype_code(grad.dtype), grad.numel(), scale.data_ptr(),
                                       stream_ptr(grad.device)), "kd_scale_inplace")
        # rare: a caller differentiates task / distill on their own (the reference's autograd allows it)
        for g_part, a_part in ((g_task, 1.0), (g_distill, 0.0)):
            if g_part is None:
                continue
            _, d_part = _stream_call(z, y, topk_v, topk_i, row_target, n_norm, tau, a_part, 1.0, True)
            d_part = d_part * g_part.to(d_part.dtype)
            grad = d_part if grad is None else grad + d_part
        if grad is not None:
            grad = grad.view(ctx.shape)
        return grad, None, None, None, None, None, None, None, None, None


def kd_loss_on_logits(student_logits, labels, teacher_logits=None, teacher_top_k_v=None, teacher_top_k_i=None,
                      speech_token_mask=None, temperature=2.0, alpha=0.5, ignore_index=IGNORE_INDEX,
                      reduce_fn=None, count_reduce_fn=None):
    """Functional form of the reference forward (distillation_loss.py:14-128) on CUDA tensors.

    Returns four 0-dim fp32 tensors (total, task, distill, teacher_task); ``total``, ``task`` and
    ``distill`` are differentiable w.r.t. ``student_logits``.  ``reduce_fn`` / ``count_reduce_fn``
    are the data-parallel hooks (all-reduce of the sums record / of the valid-row count).
    """
    require_cuda(student_logits)
    if teacher_logits is None and (teacher_top_k_v is None or teacher_top_k_i is None):
        raise ValueError("Either teacher_logits or top_k must be provided")  # distillation_loss.py:120
    z = _as_btv(student_logits, "student_logits")
    B, T, V = z.shape
    dev = z.device
    if labels.numel() != B * T:
        raise ValueError(f"labels has {labels.numel()} elements, expected {B * T}")
    row_target, n_valid = prepare_rows(labels, speech_token_mask, B, T, ignore_index, dev)
    n_norm = count_reduce_fn(n_valid) if count_reduce_fn is not None else n_valid
    y = v = i = None
    if teacher_logits is not None:  # dense wins when both are given (:56 before :73)
        y = _as_btv(teacher_logits.detach(), "teacher_logits")
        if y.device != dev:
            y = y.to(dev)
        if tuple(y.shape) != (B, T, V):
            raise ValueError(f"teacher_logits shape {tuple(y.shape)} != student_logits shape {(B, T, V)}")
    else:
        K = teacher_top_k_v.size(-1)
        # :82-90 - values to fp32 on the student's device, indices to integer
        v = teacher_top_k_v.detach().to(device=dev, dtype=torch.float32).reshape(B, T, K).contiguous()
        i = teacher_top_k_i.detach().to(device=dev, dtype=torch.int32).reshape(B, T, K).contiguous()
    return _KDOnLogits.apply(z, y, v, i, row_target, n_valid, n_norm, float(temperature), float(alpha), reduce_fn)


# --------------------------------------------------------------------------------------------
# K1 behind autograd: LM head + loss without materialising logits
# --------------------------------------------------------------------------------------------
def _fused_workspace(R, H, V, v_chunk, device, K=0):
    lib = _lib.load()
    return _workspace(lib.kd_fused_workspace_bytes(R, H, V, int(v_chunk), int(K)), device)


def logit_cache_budget(mb=None):
    """Bytes the forward may spend on the logit cache (include/kd_b200.h, "Logit cache"): a constant chosen by the
    caller - ``mb`` megabytes, else KD_LOGIT_CACHE_MB, else 6144 - never a function of V.  (The buffer itself is
    min(budget, what the R x V logits need in fp16): 1.23 GB at configs[1], the full 6 GB at configs[3]'s 16,384 rows.)"""
    import os

    if mb is None:
        mb = float(os.environ.get("KD_LOGIT_CACHE_MB", "6144"))
    return max(int(mb * (1 << 20)), 0)


def alloc_logit_cache(R, V, v_chunk, device, mb=None):
    """uint8 buffer for the encoded logits of the first vocabulary chunks (None when the budget holds none)."""
    lib = _lib.load()
    nbytes = lib.kd_fused_logit_cache_bytes(int(R), int(V), int(v_chunk), logit_cache_budget(mb))
    return torch.empty(nbytes, dtype=torch.uint8, device=device) if nbytes > 0 else None


def _teacher_kind(y, topk):
    if y is not None:
        return _lib.KD_TEACHER_DENSE
    return _lib.KD_TEACHER_SPARSE if topk is not None else _lib.KD_TEACHER_NONE


def compact_rows(row_target):
    """kd_compact_rows: (perm, inv, target_c, n_valid) - valid rows first, order preserved, no host sync."""
    lib = _lib.load()
    R = row_target.numel()
    dev = row_target.device
    perm = torch.empty(R, dtype=torch.int32, device=dev)
    inv = torch.empty(R, dtype=torch.int32, device=dev)
    target_c = torch.empty(R, dtype=torch.int32, device=dev)
    n_valid = torch.empty(1, dtype=torch.int32, device=dev)
    check(lib.kd_compact_rows(row_target.data_ptr(), R, perm.data_ptr(), inv.data_ptr(), target_c.data_ptr(),
                              n_valid.data_ptr(), stream_ptr(dev)), "kd_compact_rows")
    return perm, inv, target_c, n_valid


_compact = compact_rows  # the public functions below have a keyword argument of the same name


def gather_rows(src, row_map, zero_fill=True):
    """kd_gather_rows on a 2-D tensor with unit inner stride: out[j] = src[row_map[j]] (zeros where row_map < 0)."""
    lib = _lib.load()
    R = row_map.numel()
    out = torch.empty((R, src.size(1)), dtype=src.dtype, device=src.device)
    es = src.element_size()
    check(lib.kd_gather_rows(src.data_ptr(), src.stride(0) * es, row_map.data_ptr(), R, out.data_ptr(),
                             out.stride(0) * es, src.size(1) * es, int(bool(zero_fill)), stream_ptr(src.device)),
          "kd_gather_rows")
    return out


def _fused_forward(h, W, y, row_target, tau, alpha, v_chunk, topk=None, n_rows=None, cache=None):
    """topk = (v fp32 [R,K], i int32 [R,K]) for the sparse teacher, else None; n_rows = device count of live
    (compacted) rows or None; cache = logit-cache buffer (alloc_logit_cache) the backward will read, or None."""
    lib = _lib.load()
    R, H = h.shape
    V = W.shape[0]
    dev = h.device
    teacher_kind = _teacher_kind(y, topk)
    K = topk[0].size(-1) if teacher_kind == _lib.KD_TEACHER_SPARSE else 0
    sums = torch.empty(8, dtype=torch.float32, device=dev)
    row_stats = torch.empty((R, 4), dtype=torch.float32, device=dev)
    ws = _fused_workspace(R, H, V, v_chunk, dev, K)
    rc = lib.kd_fused_linear_fwd(
        h.data_ptr(), h.stride(0), W.data_ptr(), W.stride(0), teacher_kind,
        _ptr(y), dtype_code(y.dtype) if y is not None else 0, y.stride(0) if y is not None else 0,
        _ptr(topk[0]) if K else 0, _ptr(topk[1]) if K else 0, K, row_target.data_ptr(), _ptr(n_rows), R, H, V,
        float(tau), float(alpha), sums.data_ptr(), row_stats.data_ptr(), _ptr(cache),
        cache.numel() if cache is not None else 0, ws.data_ptr(), ws.numel(), stream_ptr(dev))
    check(rc, "kd_fused_linear_fwd")
    return sums, row_stats, ws


class _KDFusedLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, W, y, row_target, n_valid, n_norm, tau, alpha, dw_row_begin, v_chunk, reduce_fn, grad_dtype,
                topk_v=None, topk_i=None, grad_sync=None, compact=False, cache_mb=None):
        # fp32 / fp16 operands (fp32 master weights, autocast training) are cast to bf16 HERE, inside the node, so that
        # the backward can hand their gradients back from the kernels' fp32 accumulators without a bf16 rounding
        ctx.in_dtypes = (h.dtype, W.dtype)
        if h.dtype != torch.bfloat16:
            h = h.to(torch.bfloat16)
        if W.dtype != torch.bfloat16:
            W = W.to(torch.bfloat16)
        inv = n_rows = None
        if compact:  # valid rows to the front; every GEMM tile behind them is skipped (kd_rows.cu)
            perm, inv, row_target, n_rows = compact_rows(row_target)
            h = gather_rows(h, perm)
            if y is not None:
                y = gather_rows(y, perm, zero_fill=False)
            if topk_v is not None:
                topk_v, topk_i = gather_rows(topk_v, perm), gather_rows(topk_i, perm)
        topk = (topk_v, topk_i) if topk_v is not None else None
        teacher_kind = _teacher_kind(y, topk)
        cache = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:  # a backward will follow: keep the logits it needs
            cache = alloc_logit_cache(h.shape[0], W.shape[0], v_chunk, h.device, cache_mb)
        sums, row_stats, ws = _fused_forward(h, W, y, row_target, tau, alpha, v_chunk, topk, n_rows, cache)
        if reduce_fn is not None:
            sums = reduce_fn(sums)
        eff_alpha = alpha if teacher_kind != _lib.KD_TEACHER_NONE else 1.0
        losses = finalize_losses(sums, tau, eff_alpha, teacher_kind == _lib.KD_TEACHER_SPARSE)
        ctx.set_materialize_grads(False)
        ctx.cfg = (tau, eff_alpha, teacher_kind, int(dw_row_begin), int(v_chunk), grad_dtype)
        ctx.save_for_backward(h, W, y, row_target, row_stats, n_norm, topk_v, topk_i, inv, n_rows)
        ctx.ws = ws
        ctx.cache = cache
        ctx.grad_sync = grad_sync
        total, task, distill, teacher = losses.unbind(0)
        ctx.mark_non_differentiable(teacher)
        return total, task, distill, teacher

    @staticmethod
    def backward(ctx, g_total, g_task, g_distill, g_teacher):
        h, W, y, row_target, row_stats, n_norm, topk_v, topk_i, inv, n_rows = ctx.saved_tensors
        topk = (topk_v, topk_i) if topk_v is not None else None
        tau, alpha, teacher_kind, dw_row_begin, v_chunk, grad_dtype = ctx.cfg
        dev = h.device
        zero = torch.zeros((), dtype=torch.float32, device=dev)
        gt = zero if g_total is None else g_total.detach().float()
        w_ce = gt * alpha + (zero if g_task is None else g_task.detach().float())
        w_kl = gt * (1.0 - alpha) + (zero if g_distill is None else g_distill.detach().float())
        coef = torch.stack([w_ce.reshape(()), w_kl.reshape(())]).contiguous()
        h_dt, w_dt = ctx.in_dtypes
        w_grad_dtype = grad_dtype if w_dt == torch.bfloat16 else torch.float32
        dh_fp32 = h_dt != torch.bfloat16 and w_grad_dtype != torch.float32  # dH fp32 beside a bf16 dW
        dH, dW = _fused_backward(h, W, y, row_target, row_stats, n_norm, coef, tau, teacher_kind, dw_row_begin,
                                 v_chunk, w_grad_dtype, ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.ws, topk,
                                 ctx.grad_sync, dh_fp32=dh_fp32, n_rows=n_rows, cache=ctx.cache)
        if inv is not None and dH is not None:
            dH = gather_rows(dH, inv)  # back to the original row order; rows that are not scored get zeros
        if dH is not None and h_dt != torch.bfloat16 and dH.dtype != h_dt:
            dH = dH.to(h_dt)
        if dW is not None and w_dt != torch.bfloat16 and dW.dtype != w_dt:
            dW = dW.to(w_dt)
        return dH, dW, None, None, None, None, None, None, None, None, None, None, None, None, None, None, None


def _fused_backward(h, W, y, row_target, row_stats, n_norm, coef, tau, teacher_kind, dw_row_begin, v_chunk,
                    grad_dtype, need_h, need_w, ws=None, topk=None, grad_sync=None, v_offset=0, dh_fp32=False,
                    n_rows=None, cache=None):
    """kd_fused_linear_bwd, or - with ``grad_sync`` (dist.GradSync) - kd_fused_linear_bwd_range over a few
    vocabulary ranges, handing each finished dW row block to the all-reduce while the next range runs."""
    lib = _lib.load()
    R, H = h.shape
    V = W.shape[0]
    dev = h.device
    K = topk[0].size(-1) if teacher_kind == _lib.KD_TEACHER_SPARSE else 0
    dH = torch.empty((R, H), dtype=torch.float32 if dh_fp32 else grad_dtype, device=dev) if need_h else None
    dW = None
    gcode = dtype_code(grad_dtype) | (_lib.KD_GRAD_DH_F32 if dh_fp32 else 0)
    symm = None
    if need_w:
        # rows below dw_row_begin are never computed (stage1 frozen vocabulary): they stay zero
        symm = None
        if grad_sync is not None and hasattr(grad_sync, "grad_buffer"):
            symm = grad_sync.grad_buffer(V, H, grad_dtype, dev, int(dw_row_begin))
        if symm is not None:  # multimem backend: dW is written into (and reduced inside) the symmetric buffer
            dW = symm
        else:
            dW = (torch.zeros if dw_row_begin > 0 else torch.empty)((V, H), dtype=grad_dtype, device=dev)
        if n_rows is not None and dw_row_begin <= 0:
            # compacted rows and not a single live one: the dW GEMM is skipped, the gradient is zero (:47-53)
            check(lib.kd_zero_if_empty(dW.data_ptr(), dW.numel() * dW.element_size(), n_rows.data_ptr(),
                                       stream_ptr(dev)), "kd_zero_if_empty")
    if ws is None:
        ws = _fused_workspace(R, H, V, v_chunk, dev, K)
    ranges = [(0, V)]
    if grad_sync is not None and need_w:
        ranges = grad_sync.ranges(V, int(dw_row_begin), int(v_chunk))
    # the stream on which the all-reduce of a finished range is enqueued (0: the current stream)
    ready_stream = 0
    if grad_sync is not None and need_w and len(ranges) > 1 and hasattr(grad_sync, "ready_stream_ptr"):
        ready_stream = grad_sync.ready_stream_ptr(dev)
    for n, (v0, v1) in enumerate(ranges):
        flags = (_lib.KD_RANGE_FIRST if n == 0 else 0) | (_lib.KD_RANGE_LAST if n == len(ranges) - 1 else 0)
        # while an all-reduce is in flight its CTAs keep their SMs: the persistent GEMMs take the others
        sm_limit = grad_sync.sm_limit() if (grad_sync is not None and n > 0) else 0
        rc = lib.kd_fused_linear_bwd_range(
            h.data_ptr(), h.stride(0), W.data_ptr(), W.stride(0), teacher_kind,
            _ptr(y), dtype_code(y.dtype) if y is not None else 0, y.stride(0) if y is not None else 0,
            _ptr(topk[0]) if K else 0, _ptr(topk[1]) if K else 0, K, row_target.data_ptr(), _ptr(n_rows),
            row_stats.data_ptr(), R, H, V, float(tau), n_norm.data_ptr(), coef.data_ptr(), gcode, _ptr(dH), H, _ptr(dW), H,
            int(dw_row_begin), int(v_chunk), int(v0), int(v1), flags, int(sm_limit), int(v_offset), _ptr(cache),
            cache.numel() if cache is not None else 0, ws.data_ptr(), ws.numel(), ready_stream, stream_ptr(dev))
        check(rc, "kd_fused_linear_bwd_range")
        if grad_sync is not None and need_w:
            grad_sync.reduce_rows(dW, max(v0, int(dw_row_begin)), v1, last=(n == len(ranges) - 1) or not ready_stream)
    if grad_sync is not None and need_w:
        grad_sync.finish()
        if symm is not None:
            dW = grad_sync.result()  # the reduced rows, copied out of the symmetric buffer range by range
    return dH, dW


def fused_linear_kd_value_and_grad(hidden, lm_head_weight, labels, teacher_logits=None, speech_token_mask=None,
                                   temperature=2.0, alpha=0.5, ignore_index=IGNORE_INDEX, dw_row_begin=0, v_chunk=0,
                                   grad_dtype=torch.float32, teacher_top_k_v=None, teacher_top_k_i=None,
                                   compact_rows=None, logit_cache_mb=None):
    """Forward + backward in one call, outside autograd: returns (losses[4] fp32, dHidden, dWeight) with the
    gradients of ``total`` in ``grad_dtype``.  fp32 exposes the kernels' accumulators before the final
    rounding to bf16 that autograd imposes on bf16 leaves (used by the parity tests and by callers that keep
    fp32 master gradients)."""
    with torch.no_grad():
        out = fused_linear_kd_loss(hidden, lm_head_weight, labels, teacher_logits, speech_token_mask, temperature,
                                   alpha, ignore_index, dw_row_begin, v_chunk, teacher_top_k_v=teacher_top_k_v,
                                   teacher_top_k_i=teacher_top_k_i, compact_rows=compact_rows,
                                   logit_cache_mb=logit_cache_mb, _return_ctx=True)
    losses, saved = out
    h2, W, y, row_target, row_stats, n_norm, teacher_kind, eff_alpha, topk, inv, n_rows, cache = saved
    coef = torch.tensor([eff_alpha, 1.0 - eff_alpha], dtype=torch.float32, device=h2.device)
    dH, dW = _fused_backward(h2, W, y, row_target, row_stats, n_norm, coef, float(temperature), teacher_kind,
                             int(dw_row_begin), int(v_chunk), grad_dtype, True, True, None, topk, n_rows=n_rows,
                             cache=cache)
    if inv is not None:
        dH = gather_rows(dH, inv)
    return losses, dH.view(hidden.shape), dW


def fused_linear_kd_loss(hidden, lm_head_weight, labels, teacher_logits=None, speech_token_mask=None,
                         temperature=2.0, alpha=0.5, ignore_index=IGNORE_INDEX, dw_row_begin=0, v_chunk=0,
                         reduce_fn=None, count_reduce_fn=None, grad_dtype=torch.bfloat16, teacher_top_k_v=None,
                         teacher_top_k_i=None, grad_sync=None, compact_rows=None, logit_cache_mb=None,
                         _return_ctx=False):
    """``DistillationLoss(student_logits = hidden @ lm_head_weight.T, ...)`` without the logits.

    hidden [B,T,H] (or [R,H] with labels [.., T]) bf16, lm_head_weight [V,H] bf16, labels [B,T].
    Teacher: ``teacher_logits`` [B,T,V] (bf16/fp32), or the top-k cache ``teacher_top_k_v`` / ``teacher_top_k_i``
    [B,T,K] (distillation_loss.py:73-118; dense wins when both are given, :56 before :73), or neither for
    plain causal-LM cross-entropy (stage1).
    ``dw_row_begin``: first vocabulary row that receives a weight gradient (stage1.py:46-57 passes
    the old vocabulary size; rows below stay exactly zero and are never computed).
    ``grad_dtype``: torch.bfloat16 (what autograd requires for bf16 leaves) or torch.float32 - the
    unrounded fp32 accumulators, usable only with non-leaf / fp32-grad consumers (verification).
    ``compact_rows``: move the scored rows to the front on the device and skip every GEMM tile behind them
    (the reference's boolean row gather, distillation_loss.py:37-45, without its host sync).  Default (None):
    on for the top-k cache and for plain CE, where only [R,H] / [R,K] rows move; off for a dense teacher, whose
    [R,V] rows would have to be copied (worth it from roughly 15 % ignored rows on: pass True).
    ``logit_cache_mb``: budget of the forward's logit cache (None = KD_LOGIT_CACHE_MB or 6144 MB, 0 = off).  The
    cache is a constant-size buffer, independent of V: vocabulary chunks that fit are differentiated from the cached
    logits by an HBM-bound kernel beside the dW / dH GEMMs, the rest is recomputed on the tensor cores.
    ``reduce_fn`` / ``count_reduce_fn`` / ``grad_sync``: token-shard data-parallel hooks (dist.py): all-reduce of
    the sums record and of the valid-row count, and the dW all-reduce overlapped with the backward
    (the weight gradient autograd receives is then already summed over ranks).
    """
    require_cuda(hidden, lm_head_weight)
    # the tensor-core path computes in bf16 with fp32 accumulation; fp32 / fp16 operands (fp32 master weights,
    # autocast training) are cast inside the autograd node, and their gradients come back in their own dtype straight
    # from the kernels' fp32 accumulators (no bf16 rounding in between)
    if hidden.dtype not in (torch.bfloat16, torch.float32, torch.float16):
        raise TypeError(f"hidden states must be bf16, fp16 or fp32, got {hidden.dtype}")
    if lm_head_weight.dtype not in (torch.bfloat16, torch.float32, torch.float16):
        raise TypeError(f"lm_head weight must be bf16, fp16 or fp32, got {lm_head_weight.dtype}")
    if _return_ctx:
        hidden, lm_head_weight = hidden.to(torch.bfloat16), lm_head_weight.to(torch.bfloat16)
    if hidden.dim() != 3:
        raise ValueError("hidden must be [B, T, H]")
    B, T, H = hidden.shape
    V = lm_head_weight.shape[0]
    dev = hidden.device
    h2 = hidden.reshape(B * T, H)
    if h2.stride(-1) != 1:
        h2 = h2.contiguous()
    W = lm_head_weight if lm_head_weight.stride(-1) == 1 else lm_head_weight.contiguous()
    row_target, n_valid = prepare_rows(labels, speech_token_mask, B, T, ignore_index, dev)
    n_norm = count_reduce_fn(n_valid) if count_reduce_fn is not None else n_valid
    y = None
    topk_v = topk_i = None
    if teacher_logits is not None:
        y = teacher_logits.detach()
        if tuple(y.shape[-1:]) != (V,) or y.numel() != B * T * V:
            raise ValueError(f"teacher_logits shape {tuple(y.shape)} does not match [B={B}, T={T}, V={V}]")
        y = y.reshape(B * T, V)
        if y.stride(-1) != 1:
            y = y.contiguous()
        if y.dtype == torch.float16:
            y = y.float()
    elif teacher_top_k_v is not None and teacher_top_k_i is not None:
        K = teacher_top_k_v.size(-1)
        if teacher_top_k_v.numel() != B * T * K or teacher_top_k_i.numel() != B * T * K:
            raise ValueError(f"teacher top-k tensors do not match [B={B}, T={T}, K={K}]")
        # :82-90 - values to fp32 on the student's device, indices to integer
        topk_v = teacher_top_k_v.detach().to(device=dev, dtype=torch.float32).reshape(B * T, K).contiguous()
        topk_i = teacher_top_k_i.detach().to(device=dev, dtype=torch.int32).reshape(B * T, K).contiguous()
    topk = (topk_v, topk_i) if topk_v is not None else None
    teacher_kind = _teacher_kind(y, topk)
    compact = bool(compact_rows) if compact_rows is not None else teacher_kind != _lib.KD_TEACHER_DENSE
    if _return_ctx:  # fused_linear_kd_value_and_grad: forward pieces without an autograd node
        inv = n_rows = None
        hd = h2.detach()
        if compact:
            perm, inv, row_target, n_rows = _compact(row_target)
            hd = gather_rows(hd, perm)
            if y is not None:
                y = gather_rows(y, perm, zero_fill=False)
            if topk is not None:
                topk = (gather_rows(topk[0], perm), gather_rows(topk[1], perm))
        cache = alloc_logit_cache(hd.shape[0], V, v_chunk, dev, logit_cache_mb)
        sums, row_stats, _ = _fused_forward(hd, W.detach(), y, row_target, float(temperature), float(alpha), v_chunk,
                                            topk, n_rows, cache)
        if reduce_fn is not None:
            sums = reduce_fn(sums)
        eff_alpha = float(alpha) if teacher_kind != _lib.KD_TEACHER_NONE else 1.0
        losses = finalize_losses(sums, temperature, eff_alpha, teacher_kind == _lib.KD_TEACHER_SPARSE)
        return losses, (hd, W.detach(), y, row_target, row_stats, n_norm, teacher_kind, eff_alpha, topk, inv, n_rows,
                        cache)
    out = _KDFusedLinear.apply(h2, W, y, row_target, n_valid, n_norm, float(temperature), float(alpha),
                               int(dw_row_begin), int(v_chunk), reduce_fn, grad_dtype, topk_v, topk_i, grad_sync,
                               compact, logit_cache_mb)
    return out


# --------------------------------------------------------------------------------------------
# the reference-facing module
# --------------------------------------------------------------------------------------------
class DistillationLoss(nn.Module):
    """Drop-in for the reference ``DistillationLoss`` (distillation_loss.py:6-128).

    Same constructor (plus ``ignore_index``, default -100 as hard-coded at :39-41), same forward
    signature, same 4-tuple.  Opt-in fused form: pass ``student_hidden=`` and ``lm_head_weight=``
    (and ``student_logits=None``) to skip the logits tensor altogether.
    """

    def __init__(self, temperature=2.0, alpha=0.5, ignore_index=IGNORE_INDEX, reference_dtypes=True):
        super().__init__()
        self.temperature = temperature
        self.alpha = alpha
        self.ignore_index = ignore_index
        self.reference_dtypes = reference_dtypes

    def forward(self, student_logits, labels, teacher_logits=None, teacher_top_k_v=None, teacher_top_k_i=None,
                speech_token_mask=None, student_hidden=None, lm_head_weight=None):
        from .lazy import LazyLogits  # outputs.logits of a model patched by enable_lazy_logits

        if isinstance(teacher_logits, LazyLogits):
            # dense teacher from a patched teacher model: its logits are needed as a tensor (kd_linear_bf16)
            teacher_logits = teacher_logits.materialize()
        if isinstance(student_logits, LazyLogits):
            if student_logits.log_probs:
                raise ValueError("student_logits must be logits, not log-probabilities")
            student_hidden, lm_head_weight = student_logits.hidden, student_logits.head_weight()
            student_logits = None
        if student_logits is None:
            if student_hidden is None or lm_head_weight is None:
                raise ValueError("pass student_logits, or student_hidden together with lm_head_weight")
            if teacher_logits is None and (teacher_top_k_v is None or teacher_top_k_i is None):
                raise ValueError("Either teacher_logits or top_k must be provided")  # distillation_loss.py:120
            out = fused_linear_kd_loss(student_hidden, lm_head_weight, labels, teacher_logits=teacher_logits,
                                       speech_token_mask=speech_token_mask, temperature=self.temperature,
                                       alpha=self.alpha, ignore_index=self.ignore_index,
                                       teacher_top_k_v=teacher_top_k_v, teacher_top_k_i=teacher_top_k_i)
            ref_dtype = student_hidden.dtype
            dense = teacher_logits is not None
        else:
            out = kd_loss_on_logits(student_logits, labels, teacher_logits, teacher_top_k_v, teacher_top_k_i,
                                    speech_token_mask, self.temperature, self.alpha, self.ignore_index)
            ref_dtype = student_logits.dtype
            dense = teacher_logits is not None
        total, task, distill, teacher = out
        if self.reference_dtypes and ref_dtype != torch.float32:
            # dtypes the reference returns: dense -> all in the logits dtype;
            # sparse -> task in the logits dtype, the rest fp32
            task = task.to(ref_dtype)
            if dense:
                total, distill, teacher = total.to(ref_dtype), distill.to(ref_dtype), teacher.to(ref_dtype)
        return total, task, distill, teacher
