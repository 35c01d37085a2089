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
    No [.., V] temporary is written: one read of the logits.
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
        check(lib.kd_topk_logprobs(x2.data_ptr(), dtype_code(x2.dtype), R, V, x2.stride(0), int(k), out_v.data_ptr(),
                                   out_i.data_ptr(), stream_ptr(dev)), "kd_topk_logprobs")
    return out_v.reshape(*lead, k), out_i.reshape(*lead, k)


def extract_batch(teacher_logits, attention_mask, k):
    """Per-sample truncation of extract_teacher_logits.py:120-129 without its per-sample D2H loop:
    one kernel, one device->host copy.  Returns two lists of numpy arrays ([len_b, k] fp16 / int32)."""
    v, i = teacher_topk_logprobs(teacher_logits, k)
    lengths = attention_mask.sum(dim=1).tolist()
    v_cpu, i_cpu = v.cpu().numpy(), i.cpu().numpy()
    return ([v_cpu[b, : int(n)] for b, n in enumerate(lengths)], [i_cpu[b, : int(n)] for b, n in enumerate(lengths)])
