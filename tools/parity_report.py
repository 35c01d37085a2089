"""Parity report for K1 at several sizes up to BASELINE's full configuration (run on the GPU box).

For each case: error of (a) our bf16 gradients, (b) our fp32 accumulators, (c) the reference's own
all-bf16 torch GPU pipeline, each against the reference loss evaluated in fp32 (fp64 for the small
cases) on the bf16-rounded inputs - the protocol of SURVEY.md 8(d).  Writes gpurun_out/parity_report.json.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K  # noqa: E402
from oracle import kd_oracle as O  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()))


def case(B, T, H, V, tau, alpha, oracle_dtype):
    g = torch.Generator(device="cuda").manual_seed(B * T + V)
    h = torch.randn(B, T, H, device="cuda", generator=g).bfloat16()
    W = (torch.randn(V, H, device="cuda", generator=g) * (2.0 / H ** 0.5)).bfloat16()
    y = torch.empty(B, T, V, device="cuda", dtype=torch.bfloat16)
    for b in range(B):
        y[b] = (torch.randn(T, V, device="cuda", generator=g) * 2).bfloat16()
    labels = torch.randint(0, V, (B, T), device="cuda", generator=g)
    labels[:, : T // 4] = -100
    # oracle (reference op sequence) in high precision, chunked over batch to bound memory
    hr = h.to(oracle_dtype).requires_grad_(True)
    Wr = W.to(oracle_dtype).requires_grad_(True)
    ref = O.reference_loss(hr @ Wr.t(), labels, teacher_logits=y.to(oracle_dtype), temperature=tau, alpha=alpha)
    ref[0].backward()
    gh_ref, gw_ref = hr.grad, Wr.grad
    ref_l = [float(x) for x in ref]
    del ref
    torch.cuda.empty_cache()
    # ours, bf16 grads through autograd
    hc, Wc = h.clone().requires_grad_(True), W.clone().requires_grad_(True)
    out = K.fused_linear_kd_loss(hc, Wc, labels, teacher_logits=y, temperature=tau, alpha=alpha)
    out[0].backward()
    ours_l = [float(x) for x in out]
    l32, gh32, gw32 = K.fused_linear_kd_value_and_grad(h, W, labels, teacher_logits=y, temperature=tau, alpha=alpha)
    # reference all-bf16 GPU pipeline
    hb, Wb = h.clone().requires_grad_(True), W.clone().requires_grad_(True)
    outb = O.reference_loss(torch.nn.functional.linear(hb, Wb), labels, teacher_logits=y, temperature=tau, alpha=alpha)
    outb[0].backward()
    rec = {
        "shape": dict(B=B, T=T, H=H, V=V, tau=tau, alpha=alpha, oracle=str(oracle_dtype)),
        "loss_ref": ref_l, "loss_ours": ours_l, "loss_torch_bf16": [float(x) for x in outb],
        "loss_rel_err_ours": max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(ours_l, ref_l)),
        "loss_rel_err_torch_bf16": max(abs(float(a) - b) / max(abs(b), 1e-30) for a, b in zip(outb, ref_l)),
        "dH": {"ours_bf16": rel(hc.grad, gh_ref), "ours_fp32": rel(gh32, gh_ref), "torch_bf16": rel(hb.grad, gh_ref),
               "cos_ours_bf16": cos(hc.grad, gh_ref)},
        "dW": {"ours_bf16": rel(Wc.grad, gw_ref), "ours_fp32": rel(gw32, gw_ref), "torch_bf16": rel(Wb.grad, gw_ref),
               "cos_ours_bf16": cos(Wc.grad, gw_ref)},
    }
    print(json.dumps(rec), flush=True)
    return rec


if __name__ == "__main__":
    out = []
    out.append(case(2, 128, 256, 5000, 2.0, 0.5, torch.float64))
    out.append(case(2, 256, 1024, 20000, 2.0, 0.5, torch.float64))
    out.append(case(4, 512, 1024, 152936, 2.0, 0.5, torch.float32))
    if "--full" in sys.argv:
        out.append(case(8, 512, 1024, 152936, 2.0, 0.5, torch.float32))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/parity_report.json", "w"), indent=1)
