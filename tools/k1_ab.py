"""K1 step A/B at BASELINE configs[1] size under the environment knobs of this process (one process per
configuration: the library reads its knobs once).  Prints one JSON line: ms/step over a seconds-long loop,
forward / backward split by CUDA events, and the fp32-accumulator + bf16 gradient error against the fp32 oracle.

    KD_LOGIT_CACHE_MB=0 python tools/k1_ab.py        # round-1 behaviour: recompute GEMM in the backward
    KD_GRAD_OVERLAP=0  python tools/k1_ab.py         # cached gradient kernel not overlapped with the GEMMs
    KD_DW_ORDER=m      python tools/k1_ab.py         # round-1 dW unit order
"""
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K  # noqa: E402
from speech_distill_b200 import loss as KL  # noqa: E402

B, T, H, V = 8, 512, 1024, 152936
SECONDS = float(os.environ.get("AB_SECONDS", "2.0"))
PARITY = os.environ.get("AB_PARITY", "1") == "1"
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1234)
h = torch.randn(B, T, H, device=dev, generator=g).bfloat16()
W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16()
y = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
for b in range(B):
    y[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
labels = torch.randint(0, V, (B, T), device=dev, generator=g)

h2, y2 = h.reshape(B * T, H), y.reshape(B * T, V)
row_target, n_valid = KL.prepare_rows(labels, None, B, T, -100, dev)
coef = torch.tensor([0.5, 0.5], dtype=torch.float32, device=dev)
cache = KL.alloc_logit_cache(B * T, V, 0, dev)


def step(evs=None):
    if evs:
        evs[0].record()
    sums, row_stats, ws = KL._fused_forward(h2, W, y2, row_target, 2.0, 0.5, 0, cache=cache)
    if evs:
        evs[1].record()
    KL._fused_backward(h2, W, y2, row_target, row_stats, n_valid, coef, 2.0, 1, 0, 0, torch.bfloat16, True, True, ws,
                       cache=cache)
    if evs:
        evs[2].record()


for _ in range(5):
    step()
torch.cuda.synchronize()
# how many steps fill SECONDS
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
n = max(20, int(SECONDS * 1e3 / (e0.elapsed_time(e1) / 10)))
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
t0 = time.time()
for i in range(n):
    step(evs[i])
torch.cuda.synchronize()
wall = time.time() - t0
total = evs[0][0].elapsed_time(evs[-1][2]) / n
# settled regime: the last half of the loop
half = n // 2
settled = evs[half][0].elapsed_time(evs[-1][2]) / (n - half)
fwd = statistics.median(e[0].elapsed_time(e[1]) for e in evs[half:])
bwd = statistics.median(e[1].elapsed_time(e[2]) for e in evs[half:])
rec = {
    "env": {k: v for k, v in os.environ.items() if k.startswith("KD_")},
    "cache_mb": (cache.numel() / 2 ** 20) if cache is not None else 0,
    "steps": n, "ms_per_step": total, "ms_per_step_settled": settled, "fwd_ms": fwd, "bwd_ms": bwd, "wall_s": wall,
    "tflops_algorithmic_settled": 6.0 * B * T * H * V / (settled * 1e-3) / 1e12,
}

if PARITY:
    from oracle import kd_oracle as O  # checker only

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max())

    hr = h.float().requires_grad_(True)
    Wr = W.float().requires_grad_(True)
    ref = O.reference_loss(hr @ Wr.t(), labels, teacher_logits=y.float(), temperature=2.0, alpha=0.5)
    ref[0].backward()
    ref_l = [float(x) for x in ref]
    gh_ref, gw_ref = hr.grad, Wr.grad
    del ref
    torch.cuda.empty_cache()
    l32, gh32, gw32 = K.fused_linear_kd_value_and_grad(h, W, labels, teacher_logits=y, temperature=2.0, alpha=0.5)
    hc, Wc = h.clone().requires_grad_(True), W.clone().requires_grad_(True)
    out = K.fused_linear_kd_loss(hc, Wc, labels, teacher_logits=y, temperature=2.0, alpha=0.5)
    out[0].backward()
    rec["parity"] = {
        "loss_rel": max(abs(float(a) - b) / max(abs(b), 1e-30) for a, b in zip(l32, ref_l)),
        "dH_fp32": rel(gh32.reshape(B * T, H), gh_ref.reshape(B * T, H)), "dW_fp32": rel(gw32, gw_ref),
        "dH_bf16": rel(hc.grad, gh_ref), "dW_bf16": rel(Wc.grad, gw_ref),
    }
print(json.dumps(rec), flush=True)
