"""Teacher top-k log-prob compaction (K3) - host mirror of the three torch calls at
reference ``extract_teacher_logits.py:114-129`` and ``train.py:82-91``."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, dtype_code, require_cuda, stream_ptr

MAX_K = 512


def teacher_topk_logprobs(teacher_logits, k, vocab_size=None):
    """``log_softmax(logits[..., :vocab_size]) -> topk(k) -> (fp16 values, int32 indices)``.

    ``vocab_size`` mirrors train.py:82-83 (truncate the teacher to the student's vocabulary).
    Order: log-prob descending; equal logits by ascending index (hard part 4).
    No [.., V] temporary is written: one read of the logits (plus the ~k 64-byte pieces per row that can hold a
    top-k entry).
    """
    require_cuda(teacher_logits)
    lib = _lib.load()
    x = teacher_logits.detach()
    if vocab_size is not None and vocab_size < x.size(-1):
        x = x[..., :vocab_size]
    V = x.size(-1)
    if not (1 <= k <= min(V, MAX_K)):
        raise ValueError(f"k={k} must be in [1, min(V={V}, {MAX_K})]")  # torch.topk raises on k > V as well
    lead = x.shape[:-1]
    x2 = x.reshape(-1, V) if x.dim() != 2 else x
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    R = x2.size(0)
    dev = x2.device
    out_v = torch.empty((R, k), dtype=torch.float16, device=dev)
    out_i = torch.empty((R, k), dtype=torch.int32, device=dev)
    if R > 0:
        # workspace of the two-kernel form (piece maxima + partial log-sum-exp records, ~1/32 of the logits); the
        # library's internal sweep stream reads and writes it, but every use is joined back into the current stream
        # before the call's last kernel, so the caching allocator's stream bookkeeping stays valid
        ws_bytes = int(lib.kd_topk_workspace_bytes(R, V))
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (ws.data_ptr() + 255) & ~255
        check(lib.kd_topk_logprobs_ws(x2.data_ptr(), dtype_code(x2.dtype), R, V, x2.stride(0), int(k), out_v.data_ptr(),
                                      out_i.data_ptr(), ws_ptr, ws_bytes, stream_ptr(dev)), "kd_topk_logprobs_ws")
    return out_v.reshape(*lead, k), out_i.reshape(*lead, k)


def extract_batch(teacher_logits, attention_mask, k):
    """Per-sample truncation of extract_teacher_logits.py:120-129 without its per-sample D2H loop:
    one kernel, one device->host copy.  Returns two lists of numpy arrays ([len_b, k] fp16 / int32)."""
    v, i = teacher_topk_logprobs(teacher_logits, k)
    lengths = attention_mask.sum(dim=1).tolist()
    v_cpu, i_cpu = v.cpu().numpy(), i.cpu().numpy()
    return ([v_cpu[b, : int(n)] for b, n in enumerate(lengths)], [i_cpu[b, : int(n)] for b, n in enumerate(lengths)])


def linear_bf16(hidden, weight, out=None):
    """``hidden @ weight.T`` in bf16 with fp32 accumulation on the K1 tensor pipeline (kd_linear_bf16).
    hidden [R,H], weight [V,H] bf16; ``out`` [R, >=V] bf16 with a row stride that is a multiple of 8."""
    require_cuda(hidden, weight)
    lib = _lib.load()
    if hidden.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise TypeError("linear_bf16 needs bf16 operands")
    R, H = hidden.shape
    V = weight.shape[0]
    if out is None:
        ld = -(-V // 8) * 8
        out = torch.empty((R, ld), dtype=torch.bfloat16, device=hidden.device)[:, :V]
    check(lib.kd_linear_bf16(hidden.data_ptr(), hidden.stride(0), weight.data_ptr(), weight.stride(0), out.data_ptr(),
                             out.stride(0), R, H, V, stream_ptr(hidden.device)), "kd_linear_bf16")
    return out


def head_topk_layout(V):
    """(pmax_stride, part_stride) of the selection statistics for a vocabulary of V columns (kd_head_topk_layout)."""
    import ctypes

    a, b = ctypes.c_int(0), ctypes.c_int(0)
    check(_lib.load().kd_head_topk_layout(int(V), ctypes.byref(a), ctypes.byref(b)), "kd_head_topk_layout")
    return a.value, b.value


def teacher_head_topk(hidden, lm_head_weight, k, vocab_size=None, row_block=2048, fused=True):
    """Teacher LM head -> log_softmax -> top-k without the teacher's [B,T,V] logits
    (train.py:60-94 on-the-fly mode, extract_teacher_logits.py:109-129).

    hidden [..., H_t] bf16 (the teacher body's last hidden states), lm_head_weight [V_t, H_t] bf16;
    ``vocab_size`` truncates the teacher vocabulary to the student's (train.py:82-83) by dropping weight rows,
    which is the same as slicing the logits.  Rows are processed ``row_block`` at a time through two fixed scratch
    buffers of row_block x V bf16: the head GEMM of block b + 1 (tensor-core bound, current stream) runs beside the
    selection of block b (side stream).  Every block re-reads the teacher's lm_head weight (626 MB for SoulX-1.7B), so
    larger blocks cost less: measured at configs[2] size (tools/head_topk_sweep.py, one B200, fused) 4.90 ms with
    1024-row blocks, 4.70 ms with 2048 (default: 2 x 0.63 GB of scratch), 4.33 ms with 4096 (2 x 1.25 GB - the size of
    the whole [B*T, V] logits this function exists to avoid at B*T = 8192), 3.34 ms for the head GEMM alone.

    ``fused=True`` (default): the GEMM's epilogue (kd_head_logits_stats) also leaves the maximum of every 32-column
    piece and partial log-sum-exp records, and the selection (kd_head_topk_select) reads only the ~k pieces per row
    that can hold a top-k entry - a few KB per row, so it no longer competes with the next block's GEMM for HBM and
    SMs.  ``fused=False``: plain kd_linear_bf16 + the full-row compaction kernel (kd_topk_logprobs).  Both return
    (values fp16 [..., k], indices int32 [..., k]); the indices are identical, the values can differ in the last bit
    of the fp32 log-sum-exp (summation order) before their rounding to bf16 / fp16.
    """
    import ctypes

    require_cuda(hidden, lm_head_weight)
    lead = hidden.shape[:-1]
    H = hidden.shape[-1]
    if hidden.dtype != torch.bfloat16:          # fp32 / fp16 teachers: the head runs in bf16 like the student's
        hidden = hidden.detach().to(torch.bfloat16)
    if lm_head_weight.dtype != torch.bfloat16:
        lm_head_weight = lm_head_weight.detach().to(torch.bfloat16)
    h2 = hidden.detach().reshape(-1, H)
    if h2.stride(-1) != 1:
        h2 = h2.contiguous()
    W = lm_head_weight.detach()
    if vocab_size is not None and vocab_size < W.size(0):
        W = W[:vocab_size]
    if W.stride(-1) != 1:
        W = W.contiguous()
    V = W.size(0)
    if not (1 <= k <= min(V, MAX_K)):
        raise ValueError(f"k={k} must be in [1, min(V={V}, {MAX_K})]")
    R = h2.size(0)
    dev = h2.device
    out_v = torch.empty((R, k), dtype=torch.float16, device=dev)
    out_i = torch.empty((R, k), dtype=torch.int32, device=dev)
    if R == 0:
        return out_v.reshape(*lead, k), out_i.reshape(*lead, k)
    rb = max(1, min(int(row_block), R))
    ld = -(-V // 8) * 8
    n_sets = 2 if R > rb else 1
    scratch = [torch.empty((rb, ld), dtype=torch.bfloat16, device=dev) for _ in range(n_sets)]
    if fused:
        pmax_stride, part_stride = head_topk_layout(V)
        pmax = [torch.empty((rb, pmax_stride), dtype=torch.bfloat16, device=dev) for _ in range(n_sets)]
        part = [torch.empty((rb, part_stride, 2), dtype=torch.float32, device=dev) for _ in range(n_sets)]
    caller = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    # fused: the GEMMs go to an internal high-priority stream, so that when a GEMM and a selection kernel are both
    # waiting for SMs the persistent GEMM CTAs are placed first and the selection's CTAs fill the shared memory the
    # GEMM leaves free (the caller's stream usually has the lowest priority there is and cannot be outranked by less)
    main = _side_stream(dev, high_priority=True) if fused and n_sets > 1 else caller
    if main is not caller:
        main.wait_stream(caller)
    filled = [torch.cuda.Event() for _ in scratch]
    drained = [None for _ in scratch]
    lib = _lib.load()
    for n, r0 in enumerate(range(0, R, rb)):
        r1 = min(r0 + rb, R)
        s = n % n_sets
        if drained[s] is not None:
            main.wait_event(drained[s])  # the selection that read this scratch two blocks ago is done
        hb = h2[r0:r1]
        logits = scratch[s][: r1 - r0, :V]
        if fused:
            n_part = ctypes.c_int(0)
            check(lib.kd_head_logits_stats(hb.data_ptr(), hb.stride(0), W.data_ptr(), W.stride(0), logits.data_ptr(),
                                           logits.stride(0), pmax[s].data_ptr(), pmax_stride, part[s].data_ptr(),
                                           part_stride, ctypes.byref(n_part), r1 - r0, H, V, main.cuda_stream),
                  "kd_head_logits_stats")
        else:
            linear_bf16(hb, W, logits)
        filled[s].record(main)
        with torch.cuda.stream(side):
            side.wait_event(filled[s])
            if fused:
                check(lib.kd_head_topk_select(logits.data_ptr(), logits.stride(0), pmax[s].data_ptr(), pmax_stride,
                                              part[s].data_ptr(), part_stride, n_part.value, r1 - r0, V, int(k),
                                              out_v[r0:r1].data_ptr(), out_i[r0:r1].data_ptr(), stream_ptr(dev)),
                      "kd_head_topk_select")
            else:
                check(lib.kd_topk_logprobs(logits.data_ptr(), dtype_code(logits.dtype), r1 - r0, V, logits.stride(0),
                                           int(k), out_v[r0:r1].data_ptr(), out_i[r0:r1].data_ptr(), stream_ptr(dev)),
                      "kd_topk_logprobs")
            ev = torch.cuda.Event()
            ev.record(side)
            drained[s] = ev
    for ev in drained:
        if ev is not None:
            caller.wait_event(ev)
    if main is not caller:
        caller.wait_stream(main)
    return out_v.reshape(*lead, k), out_i.reshape(*lead, k)


_SIDE_STREAMS = {}


def _side_stream(dev, high_priority=False):
    idx = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    key = (idx, bool(high_priority))
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev, priority=-1 if high_priority else 0)
    return _SIDE_STREAMS[key]
