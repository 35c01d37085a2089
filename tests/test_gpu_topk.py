"""GPU parity of K3 (teacher top-k log-prob compaction): indices bit-exact on tie-free inputs,
documented tie rule otherwise (hard part 4)."""
import os

import numpy as np
import pytest
import torch

from oracle import kd_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_golden_fp32_indices_bit_exact():
    import speech_distill_b200 as K

    d = np.load(os.path.join(GOLDEN, "topk_f32.npz"))
    v, i = K.teacher_topk_logprobs(torch.from_numpy(d["logits"]).cuda(), int(d["k"]))
    assert v.dtype == torch.float16 and i.dtype == torch.int32
    np.testing.assert_array_equal(i.cpu().numpy(), d["i"])
    assert np.abs(v.float().cpu().numpy() - d["v"].astype(np.float32)).max() <= 2 ** -7  # <= 1 fp16 ulp at |v|~8


@pytest.mark.parametrize("R,V,k,dtype", [(7, 152936, 64, torch.float32), (5, 152936, 128, torch.float32),
                                         (3, 159488, 100, torch.float32), (9, 1031, 16, torch.float32),
                                         (4, 700, 512, torch.float32), (2, 64, 64, torch.float32), (3, 5000, 1, torch.float32)])
def test_fp32_matches_torch_topk(R, V, k, dtype):
    import speech_distill_b200 as K

    g = torch.Generator().manual_seed(R * V + k)
    x = torch.randn(R, V, generator=g) * 3
    v_ref, i_ref = O.topk_logprobs_reference(x, k)
    v, i = K.teacher_topk_logprobs(x.cuda(), k)
    np.testing.assert_array_equal(i.cpu().numpy(), i_ref.numpy())
    assert (v.float().cpu() - v_ref.float()).abs().max() <= 2 ** -6


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_16bit_inputs_tie_rule(dtype):
    import speech_distill_b200 as K

    g = torch.Generator().manual_seed(11)
    x = (torch.randn(6, 152936, generator=g) * 2).to(dtype)
    k = 64
    v, i = K.teacher_topk_logprobs(x.cuda(), k)
    v_spec, i_spec = O.topk_spec(x, k)
    np.testing.assert_array_equal(i.cpu().numpy(), i_spec.numpy())  # deterministic spec: bit-exact
    assert (v.float().cpu() - v_spec.float()).abs().max() <= 2 ** -5
    # and a valid top-k of the reference: same multiset of selected logits as torch.topk on the raw logits
    top_ref = torch.topk(x.float(), k, -1).values
    sel = torch.gather(x.float(), -1, i.cpu().long())
    assert torch.equal(sel, top_ref)
    # values sorted descending, indices ascending inside equal values
    xs = sel.numpy()
    ii = i.cpu().numpy()
    assert (np.diff(xs, axis=-1) <= 0).all()
    assert ((np.diff(xs, axis=-1) < 0) | (np.diff(ii, axis=-1) > 0)).all()


def test_massive_ties_take_the_exact_path():
    import speech_distill_b200 as K

    x = torch.zeros(3, 20000)
    x[1, ::2] = 1.0  # 10000 ties at the top
    x[2, 5] = 3.0
    x[2, 17] = 3.0
    v, i = K.teacher_topk_logprobs(x.cuda(), 32)
    i = i.cpu().numpy()
    np.testing.assert_array_equal(i[0], np.arange(32))
    np.testing.assert_array_equal(i[1], np.arange(0, 64, 2))
    np.testing.assert_array_equal(i[2], np.array([5, 17] + [j for j in range(40) if j not in (5, 17)][:30]))
    v_spec, i_spec = O.topk_spec(x, 32)
    np.testing.assert_array_equal(i, i_spec.numpy())


def test_truncation_and_batch_shape():
    import speech_distill_b200 as K

    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 5, 3000, generator=g)
    v, i = K.teacher_topk_logprobs(x.cuda(), 8, vocab_size=2500)  # train.py:82-83
    v_ref, i_ref = O.topk_logprobs_reference(x[..., :2500], 8)
    assert v.shape == (2, 5, 8)
    np.testing.assert_array_equal(i.cpu().numpy(), i_ref.numpy())
    mask = torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1]])
    vs, is_ = K.extract_batch(x.cuda(), mask.cuda(), 8)
    assert [a.shape for a in vs] == [(3, 8), (5, 8)] and is_[0].dtype == np.int32 and vs[0].dtype == np.float16


def test_topk_feeds_sparse_loss_like_train_py():
    """train.py:74-104: on-the-fly top-k of the teacher, then the sparse loss."""
    import speech_distill_b200 as K

    g = torch.Generator().manual_seed(9)
    z = torch.randn(2, 6, 8192, generator=g) * 2
    y = torch.randn(2, 6, 8192, generator=g) * 2
    lab = torch.randint(0, 8192, (2, 6), generator=g)
    v_ref, i_ref = O.topk_logprobs_reference(y, 64)
    ref = O.reference_loss(z, lab, teacher_top_k_v=v_ref, teacher_top_k_i=i_ref)
    v, i = K.teacher_topk_logprobs(y.cuda(), 64)
    out = K.kd_loss_on_logits(z.cuda(), lab.cuda(), teacher_top_k_v=v, teacher_top_k_i=i)
    np.testing.assert_allclose([float(o.detach()) for o in out], [float(r.detach()) for r in ref], rtol=1e-3)


# ---- teacher LM head -> top-k without the teacher's [B,T,V] logits -------------------------
@pytest.mark.parametrize("R,H,V", [(300, 256, 1031), (128, 512, 5000), (1000, 2048, 20000)])
def test_linear_bf16_is_the_bf16_lm_head(R, H, V):
    """kd_linear_bf16 = fp32-accumulated h W^T rounded once to bf16 (what HF's bf16 nn.Linear returns up to the
    summation order): within one bf16 ulp of the fp64 product, ragged R / V handled by TMA clipping."""
    import speech_distill_b200 as K

    g = torch.Generator(device="cuda").manual_seed(R + V)
    h = torch.randn(R, H, device="cuda", generator=g).bfloat16()
    W = (torch.randn(V, H, device="cuda", generator=g) * (2.0 / H ** 0.5)).bfloat16()
    guard = torch.full((R, -(-V // 8) * 8), 777.0, dtype=torch.bfloat16, device="cuda")
    out = K.linear_bf16(h, W, guard[:, :V])
    ref = h.double() @ W.double().t()
    err = (out.double() - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -8 + 1e-4).all())  # half a bf16 ulp + fp32 accumulation noise near 0
    pad = guard[:, V:]  # padding up to the 16-byte row granule: untouched or zero-filled (TMA stores whole granules)
    assert bool(((pad == 777.0) | (pad == 0.0)).all())
    exact = (out == ref.to(torch.bfloat16)).float().mean()
    assert float(exact) > 0.98  # the rest are round-to-nearest ties decided by fp32 vs fp64 accumulation


def _values_agree(v, v2):
    """fused vs full-row compaction: same fp16 values except where the fp32 log-sum-exp (summed in another order)
    moves a log-prob across a bf16 rounding boundary - then by one bf16 ulp, in well under 1 % of the entries"""
    diff = (v.float() - v2.float()).abs()
    assert bool((diff <= 2.0 ** -7 * v2.float().abs() + 1e-6).all())
    assert float((diff > 0).float().mean()) < 0.01


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("B,T,H,V,k,rb", [(2, 100, 256, 5000, 64, 64), (3, 128, 512, 20000, 100, 1024), (1, 50, 128, 700, 16, 32),
                                          (2, 300, 256, 40000, 200, 256)])
def test_teacher_head_topk_equals_topk_of_its_logits(B, T, H, V, k, rb, fused):
    """Row-block pipeline (head GEMM on the current stream, selection on a side stream, two scratch buffers)
    = kd_topk_logprobs on the materialised bf16 logits of the same GEMM: indices bit for bit (both forms), values bit
    for bit (fused=False) or to the fp32 summation order of the log-sum-exp (fused=True: piece maxima + partial
    records from the GEMM epilogue); and the deterministic tie-rule spec of the oracle on those logits."""
    import speech_distill_b200 as K

    g = torch.Generator(device="cuda").manual_seed(B * T + V)
    h = torch.randn(B, T, H, device="cuda", generator=g).bfloat16()
    W = (torch.randn(V + 40, H, device="cuda", generator=g) * (3.0 / H ** 0.5)).bfloat16()  # teacher vocab > student's
    v, i = K.teacher_head_topk(h, W, k, vocab_size=V, row_block=rb, fused=fused)
    assert v.shape == (B, T, k) and v.dtype == torch.float16 and i.dtype == torch.int32
    logits = K.linear_bf16(h.reshape(-1, H), W[:V])
    v2, i2 = K.teacher_topk_logprobs(logits.reshape(B, T, V), k)
    assert torch.equal(i, i2)
    if fused:
        _values_agree(v, v2)
    else:
        assert torch.equal(v, v2)
    v_spec, i_spec = O.topk_spec(logits.cpu().reshape(B, T, V), k)
    np.testing.assert_array_equal(i.cpu().numpy(), i_spec.numpy())
    # against the reference pipeline on torch's own bf16 lm_head: same indices wherever its logits agree with ours
    ref_logits = torch.nn.functional.linear(h, W[:V])
    if torch.equal(ref_logits, logits.reshape(B, T, V)):
        assert torch.equal(torch.topk(ref_logits.float(), k, -1).values,
                           torch.gather(ref_logits.float(), -1, i.long()))


def test_teacher_head_topk_fused_massive_ties_and_tiny_vocab():
    """Selection behind the head GEMM on degenerate rows: a zero hidden row (all logits 0: V-way tie -> exact slow
    path, lowest indices win) and a vocabulary smaller than one tile."""
    import speech_distill_b200 as K

    g = torch.Generator(device="cuda").manual_seed(4)
    H, V, k = 128, 6000, 32
    h = torch.randn(64, H, device="cuda", generator=g).bfloat16()
    h[5] = 0
    W = (torch.randn(V, H, device="cuda", generator=g) * (3.0 / H ** 0.5)).bfloat16()
    v, i = K.teacher_head_topk(h, W, k, row_block=64)
    assert torch.equal(i[5].cpu(), torch.arange(k, dtype=torch.int32))
    logits = K.linear_bf16(h, W)
    v2, i2 = K.teacher_topk_logprobs(logits, k)
    assert torch.equal(i, i2)
    _values_agree(v, v2)
    Ws = W[:100].contiguous()
    v, i = K.teacher_head_topk(h, Ws, 100, row_block=64)
    v2, i2 = K.teacher_topk_logprobs(K.linear_bf16(h, Ws), 100)
    assert torch.equal(i, i2)
    _values_agree(v, v2)
