"""torch.library registration (speech_distill_b200.ops): same numbers as the autograd.Function path, and a module
calling the ops compiles with fullgraph=True (no graph break at the loss)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(B=2, T=48, H=128, V=3001, seed=0):
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(B, T, H, generator=g).bfloat16().cuda()
    W = (torch.randn(V, H, generator=g) * (2.0 / H ** 0.5)).bfloat16().cuda()
    y = (torch.randn(B, T, V, generator=g) * 2).bfloat16().cuda()
    labels = torch.randint(0, V, (B, T), generator=g).cuda()
    labels[:, :5] = -100
    return h, W, y, labels


def test_ops_are_registered():
    import speech_distill_b200.ops  # noqa: F401

    for name in ("kd_loss", "fused_linear_kd", "fused_linear_kd_bwd"):
        assert hasattr(torch.ops.speech_distill_b200, name)


def test_kd_loss_op_matches_function_path():
    import speech_distill_b200 as K
    from speech_distill_b200 import ops

    h, W, y, labels = _inputs()
    z = (h.float() @ W.float().t()).bfloat16()
    za = z.clone().requires_grad_(True)
    zb = z.clone().requires_grad_(True)
    ref = K.kd_loss_on_logits(za, labels, teacher_logits=y, temperature=2.0, alpha=0.3)
    got = ops.kd_loss(zb, labels, teacher_logits=y, temperature=2.0, alpha=0.3)
    (ref[0] * 1.5).backward()
    (got[0] * 1.5).backward()
    for a, b in zip(ref, got):
        assert float(a.detach()) == float(b.detach())  # same kernel, same inputs
    assert not got[1].requires_grad and not got[3].requires_grad
    assert float((za.grad.float() - zb.grad.float()).abs().max()) <= 2.0 ** -8 * float(za.grad.float().abs().max())
    # sparse teacher
    tv, ti = K.teacher_topk_logprobs(y, 16)
    ref = K.kd_loss_on_logits(z, labels, teacher_top_k_v=tv, teacher_top_k_i=ti)
    got = ops.kd_loss(z, labels, teacher_top_k_v=tv, teacher_top_k_i=ti)
    for a, b in zip(ref, got):
        assert float(a.detach()) == float(b.detach())
    with pytest.raises(ValueError):
        ops.kd_loss(z, labels)


@pytest.mark.parametrize("teacher", ["dense", "topk", "none"])
def test_fused_op_matches_function_path(teacher):
    import speech_distill_b200 as K
    from speech_distill_b200 import ops

    h, W, y, labels = _inputs(seed=1)
    kw = {}
    if teacher == "dense":
        kw["teacher_logits"] = y
    elif teacher == "topk":
        kw["teacher_top_k_v"], kw["teacher_top_k_i"] = K.teacher_topk_logprobs(y, 16)
    ha, Wa = h.clone().requires_grad_(True), W.clone().requires_grad_(True)
    hb, Wb = h.clone().requires_grad_(True), W.clone().requires_grad_(True)
    ref = K.fused_linear_kd_loss(ha, Wa, labels, compact_rows=False, **kw)
    got = ops.fused_linear_kd_loss(hb, Wb, labels, **kw)
    ref[0].backward()
    got[0].backward()
    for a, b in zip(ref, got):
        assert float(a.detach()) == float(b.detach())
    assert torch.equal(ha.grad, hb.grad) and torch.equal(Wa.grad, Wb.grad)  # same kernels, same order


def test_ops_trace_under_torch_compile():
    from speech_distill_b200 import ops

    h, W, y, labels = _inputs(seed=2)

    class Head(torch.nn.Module):
        def __init__(self, W):
            super().__init__()
            self.weight = torch.nn.Parameter(W.clone())

        def forward(self, hidden, labels, teacher):
            total, task, distill, teacher_task = ops.fused_linear_kd_loss(hidden * 1.0, self.weight, labels,
                                                                          teacher_logits=teacher)
            return total, task

    m = Head(W)
    hc = h.clone().requires_grad_(True)
    eager = m(hc, labels, y)
    eager[0].backward()
    g_eager = (hc.grad.clone(), m.weight.grad.clone())
    hc.grad = None
    m.weight.grad = None
    compiled = torch.compile(m, backend="aot_eager", fullgraph=True)  # fullgraph: a graph break at the op would raise
    out = compiled(hc, labels, y)
    out[0].backward()
    assert float(out[0].detach()) == float(eager[0].detach()) and float(out[1]) == float(eager[1])
    assert torch.equal(hc.grad, g_eager[0]) and torch.equal(m.weight.grad, g_eager[1])

    def on_logits(z, labels, teacher):
        return ops.kd_loss(z * 1.0, labels, teacher_logits=teacher)[0]

    z = (h.float() @ W.float().t()).bfloat16().requires_grad_(True)
    e = on_logits(z, labels, y)
    c = torch.compile(on_logits, backend="aot_eager", fullgraph=True)(z, labels, y)
    assert float(e.detach()) == float(c.detach())
