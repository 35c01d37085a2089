"""Drop-in replacement module for the reference's ``distillation_loss.py``.

Put this repository's root ahead of the reference on ``sys.path`` (or copy this file next to
``train.py``) and ``from distillation_loss import DistillationLoss`` (reference ``train.py:13``)
resolves to the B200 implementation; ``train.py`` itself stays unchanged.
"""
from speech_distill_b200 import DistillationLoss  # noqa: F401

__all__ = ["DistillationLoss"]
