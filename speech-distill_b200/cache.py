"""Top-k teacher cache wire format and collator fast path (plumbing, no arithmetic).

``extract_teacher_logits.py:120-145`` stores, per sample, ``teacher_top_k_v`` fp16 ``[len, K]`` and
``teacher_top_k_i`` int32 ``[len, K]``; the reference collator (``data.py:261-271, 330-348``) rebuilds tensors from
Python lists - which turns them into fp32 / int64 - pads each with ``torch.cat`` and stacks.  ``pad_logits`` keeps
the wire dtypes end to end (2 + 4 bytes per entry instead of 4 + 8), writes every sample straight into one
(optionally pinned) batch buffer and honours the same contract: pad the sequence dimension to ``max_length`` with
``padding_value``, truncate longer samples.  The loss casts on the device (``distillation_loss.py:82-90``).
"""
from __future__ import annotations

import numpy as np
import torch

WIRE_DTYPES = {"teacher_top_k_v": torch.float16, "teacher_top_k_i": torch.int32}


def _as_tensor(x, dtype):
    if isinstance(x, torch.Tensor):
        return x if x.dtype == dtype else x.to(dtype)
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)  # nested Python lists (datasets' default decoding)


def pad_logits(logit_list, max_length, padding_value=0.0, dtype=None, pin_memory=False):
    """Same contract as the reference ``_pad_logits`` (data.py:330-348): ``[len_b, K]`` samples ->
    ``[B, max_length, K]``, padded with ``padding_value`` / truncated along the sequence dimension.
    ``dtype`` defaults to the first sample's dtype when it is a tensor / array, else fp32 for float padding and
    int32 for integer padding; pass ``WIRE_DTYPES[...]`` to pin the cache's wire format."""
    if not logit_list:
        raise ValueError("pad_logits needs at least one sample")
    first = logit_list[0]
    if dtype is None:
        if isinstance(first, torch.Tensor):
            dtype = first.dtype
        elif isinstance(first, np.ndarray):
            dtype = torch.from_numpy(first[:0]).dtype
        else:
            dtype = torch.float32 if isinstance(padding_value, float) else torch.int32
    K = int(_as_tensor(first, dtype).shape[1])
    out = torch.full((len(logit_list), int(max_length), K), padding_value, dtype=dtype,
                     pin_memory=bool(pin_memory) and torch.cuda.is_available())
    for b, sample in enumerate(logit_list):
        t = _as_tensor(sample, dtype)
        if t.dim() != 2 or t.shape[1] != K:
            raise ValueError(f"sample {b} has shape {tuple(t.shape)}, expected [len, {K}]")
        n = min(t.shape[0], int(max_length))
        out[b, :n] = t[:n]
    return out


def collate_teacher_topk(features, max_length, pin_memory=False):
    """The collator's "Handle Pre-calculated Teacher Logprobs" step (data.py:261-271) in wire dtypes: returns
    {} when the features carry no cache, else {"teacher_top_k_v": fp16 [B,T,K], "teacher_top_k_i": int32 [B,T,K]}."""
    top_v = [f.get("teacher_top_k_v") for f in features if "teacher_top_k_v" in f]
    top_i = [f.get("teacher_top_k_i") for f in features if "teacher_top_k_i" in f]
    if not top_v or top_v[0] is None:
        return {}
    return {
        "teacher_top_k_v": pad_logits(top_v, max_length, 0.0, WIRE_DTYPES["teacher_top_k_v"], pin_memory),
        "teacher_top_k_i": pad_logits(top_i, max_length, 0, WIRE_DTYPES["teacher_top_k_i"], pin_memory),
    }
