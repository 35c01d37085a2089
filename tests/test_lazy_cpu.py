"""Host logic of LazyLogits: metadata, the [..., :vocab] slice, torch-function dispatch of
log_softmax; materialising without CUDA must fail loudly (no CPU fallback)."""
import pytest
import torch
import torch.nn.functional as F


def test_lazy_logits_facade():
    import speech_distill_b200 as K

    h, W = torch.randn(2, 5, 8), torch.randn(100, 8)
    lz = K.LazyLogits(h, W)
    assert lz.shape == (2, 5, 100) and lz.size(-1) == 100 and lz.dim() == 3 and lz.dtype == h.dtype
    cut = lz[..., :60]                                  # train.py:82-83
    assert isinstance(cut, K.LazyLogits) and cut.size(-1) == 60 and cut.head_weight().shape == (60, 8)
    assert lz[..., :1000].size(-1) == 100               # slicing past the end clamps like a tensor slice
    lp = F.log_softmax(cut, dim=-1)                     # train.py:85
    assert isinstance(lp, K.LazyLogits) and lp.log_probs and lp.size(-1) == 60
    lp2 = torch.log_softmax(cut, -1)
    assert isinstance(lp2, K.LazyLogits) and lp2.log_probs
    with pytest.raises(K.KdError):
        torch.topk(lp, k=4, dim=-1)                     # needs the CUDA kernels
    with pytest.raises(K.KdError):
        lz.materialize()
    with pytest.raises(ValueError):
        K.LazyLogits(torch.randn(2, 7), W)


def test_distillation_loss_rejects_lazy_on_cpu():
    import speech_distill_b200 as K

    lz = K.LazyLogits(torch.randn(1, 4, 8).bfloat16(), torch.randn(16, 8).bfloat16())
    with pytest.raises(K.KdError):
        K.DistillationLoss()(lz, torch.randint(0, 16, (1, 4)), teacher_logits=torch.randn(1, 4, 16))
