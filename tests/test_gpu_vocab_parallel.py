"""Vocab-parallel mode of K1 on one GPU: the ranks' slices are run one after the other, their
records merged by kd_fused_merge_ranks, and the result compared with the unsharded kernels and the oracle.
(The NCCL exchange itself is two stock collectives; tools/vp_check.py runs the real thing on 2+ GPUs.)"""
import numpy as np
import pytest
import torch

from oracle import kd_oracle as O

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _case(seed, B, T, H, V):
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(B, T, H, generator=g).bfloat16()
    W = (torch.randn(V, H, generator=g) * (2.0 / H ** 0.5)).bfloat16()
    y = (torch.randn(B, T, V, generator=g) * 2).bfloat16()
    labels = torch.randint(0, V, (B, T), generator=g)
    labels[:, : max(1, T // 4)] = -100
    return h, W, y, labels


@pytest.mark.parametrize("G,V,sparse", [(2, 5000, False), (3, 5000, False), (4, 1031, False), (3, 5000, True),
                                        (8, 20000, True)])
def test_vocab_parallel_slices_equal_unsharded(G, V, sparse):
    import speech_distill_b200 as KD
    from speech_distill_b200 import vocab_parallel as VP

    B, T, H, tau, alpha = 2, 96, 128, 2.0, 0.5
    h, W, y, labels = _case(900 + V + G, B, T, H, V)
    kw, kw_ref = {}, {}
    if sparse:
        lp = torch.log_softmax(y.float(), -1)
        tv, ti = torch.topk(lp, 64, -1)
        tv, ti = tv.half(), ti.int()
        labels[1, T // 2] = int(ti[1, T // 2 - 1, 3])
        kw = dict(teacher_top_k_v=tv.cuda(), teacher_top_k_i=ti.cuda())
        kw_ref = dict(teacher_top_k_v=tv, teacher_top_k_i=ti)
    else:
        kw_ref = dict(teacher_logits=y.double())
    ref, gh_ref, gw_ref = O.fused_linear_reference(h.double(), W.double(), labels, temperature=tau, alpha=alpha, **kw_ref)
    hc, Wc, yc, lc = h.cuda(), W.cuda(), y.cuda(), labels.cuda()
    slices = KD.vocab_slices(V, G, align=256)
    assert slices[0][0] == 0 and slices[-1][1] == V

    # all ranks' forward records first (what the all-gather would deliver), then every rank's autograd pass
    row_target, _ = KD.loss.prepare_rows(lc, None, B, T, -100, hc.device)
    recs = []
    for v0, v1 in slices:
        topk = (kw["teacher_top_k_v"].float().reshape(B * T, -1).contiguous(),
                kw["teacher_top_k_i"].reshape(B * T, -1).contiguous()) if sparse else None
        ys = None if sparse else yc.reshape(B * T, V)[:, v0:v1]
        recs.append(VP.forward_partial(hc.reshape(B * T, H), Wc[v0:v1], ys, topk, row_target, v0, tau)[0])
    recs = torch.stack(recs)

    dh_parts, dw_parts, losses = [], [], None
    for (v0, v1) in slices:
        hg = hc.clone().requires_grad_(True)
        Wg = Wc[v0:v1].clone().requires_grad_(True)
        part = []
        out = KD.fused_linear_kd_loss_vocab_parallel(
            hg, Wg, lc, v0, teacher_logits_slice=None if sparse else yc[..., v0:v1], temperature=tau, alpha=alpha,
            gather_fn=lambda rec: recs, reduce_fn=lambda x: (part.append(x.clone()), x)[1], **kw)
        out[0].backward()
        got = [float(o.detach()) for o in out]
        if losses is not None:
            assert got == losses  # identical on every rank, bit for bit
        losses = got
        dh_parts.append(part[0])
        dw_parts.append(Wg.grad)
    for got, want in zip(losses, [float(x.detach()) for x in ref]):
        assert abs(got - want) <= 1e-3 * max(1.0, abs(want)), (losses, ref)
    dH = torch.stack(dh_parts).sum(0).reshape(B, T, H)
    dW = torch.cat(dw_parts, 0)
    assert dW.shape == (V, H) and dW.dtype == torch.bfloat16
    eh, ew = rel_err(dH.cpu().numpy(), gh_ref.numpy()), rel_err(dW.float().cpu().numpy(), gw_ref.numpy())
    assert eh < 1e-3 and ew < 4e-3, (eh, ew)  # dH: fp32 partial sums; dW slices: bf16 outputs

    # against the unsharded kernels: same losses to fp32 merge-order noise, same gradients to bf16-G noise
    out1 = KD.fused_linear_kd_loss(hc, Wc, lc, teacher_logits=None if sparse else yc, temperature=tau, alpha=alpha, **kw)
    np.testing.assert_allclose(losses, [float(o.detach()) for o in out1], rtol=2e-6, atol=1e-7)
    _, gh1, gw1 = KD.fused_linear_kd_value_and_grad(hc, Wc, lc, teacher_logits=None if sparse else yc,
                                                    temperature=tau, alpha=alpha, **kw)
    assert rel_err(dH.cpu().numpy(), gh1.cpu().numpy()) < 1e-3
