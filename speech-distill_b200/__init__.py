"""speech_distill_b200 - B200 (sm_100a) implementation of the speech-distill KD hot path.

Host-side mirror of the reference interface over libkd_b200.so (C ABI: include/kd_b200.h):

* ``DistillationLoss``            - reference ``distillation_loss.py`` (drop-in module)
* ``kd_loss_on_logits``           - functional form, with data-parallel hooks
* ``fused_linear_kd_loss``        - LM head + KD without materialising logits (K1)
* ``teacher_topk_logprobs``       - reference ``extract_teacher_logits.py:114-129`` / ``train.py:82-91``
* ``freeze_model_weights`` / ``fused_linear_cross_entropy`` / ``mask_old_rows_`` - reference ``stage1.py:29-93``
* ``ops``                         - the two loss entry points as ``torch.library`` custom ops (``torch.compile``-traceable)
"""
from ._lib import KdError, LIB_PATH, load as load_library  # noqa: F401
from .loss import (DistillationLoss, fused_linear_kd_loss, fused_linear_kd_value_and_grad,  # noqa: F401
                   kd_loss_on_logits)
from .cache import collate_teacher_topk, pad_logits  # noqa: F401
from .dist import GradSync, allreduce_grad_, make_reduce_fns, plan_ranges  # noqa: F401
from .lazy import LazyLogits, enable_fused_ce, enable_lazy_logits  # noqa: F401
from .stage1 import freeze_model_weights, fused_linear_cross_entropy, mask_old_rows_  # noqa: F401
from .vocab_parallel import fused_linear_kd_loss_vocab_parallel, vocab_slices  # noqa: F401
from .topk import extract_batch, linear_bf16, teacher_head_topk, teacher_topk_logprobs  # noqa: F401

__all__ = [
    "DistillationLoss", "kd_loss_on_logits", "fused_linear_kd_loss", "fused_linear_kd_value_and_grad",
    "teacher_topk_logprobs", "teacher_head_topk", "linear_bf16", "extract_batch",
    "fused_linear_kd_loss_vocab_parallel", "vocab_slices", "pad_logits", "collate_teacher_topk", "LazyLogits", "enable_lazy_logits", "enable_fused_ce", "GradSync", "make_reduce_fns", "allreduce_grad_", "plan_ranges",
    "freeze_model_weights", "fused_linear_cross_entropy", "mask_old_rows_", "KdError", "load_library", "LIB_PATH",
]
