"""Live timeline of the K1 backward's kernels on its three streams (kd_fused_bwd_trace_*), taken in the middle of a
settled loop of steps at BASELINE configs[1] size.  Prints one line per kernel: class, chunk, start, end, duration
(ms relative to the backward's first kernel) and per-class totals - the concurrent picture ncu cannot give.

    python tools/bwd_trace.py [steps_before_trace]
"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K  # noqa: E402
from speech_distill_b200 import loss as KL  # noqa: E402

B, T, H, V = 8, 512, 1024, 152936
NAMES = {0: "cast", 1: "grad_cached", 2: "dW", 3: "dH", 4: "grad_recompute"}
dev = torch.device("cuda")
lib = K.load_library()
g = torch.Generator(device=dev).manual_seed(1234)
h = torch.randn(B * T, H, device=dev, generator=g).bfloat16()
W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16()
y = torch.empty(B * T, V, device=dev, dtype=torch.bfloat16)
for b in range(B):
    y[b * T:(b + 1) * T] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
labels = torch.randint(0, V, (B, T), device=dev, generator=g)
row_target, n_valid = KL.prepare_rows(labels, None, B, T, -100, dev)
coef = torch.tensor([0.5, 0.5], dtype=torch.float32, device=dev)
cache = KL.alloc_logit_cache(B * T, V, 0, dev)


def fwd():
    return KL._fused_forward(h, W, y, row_target, 2.0, 0.5, 0, cache=cache)


def bwd(row_stats, ws):
    KL._fused_backward(h, W, y, row_target, row_stats, n_valid, coef, 2.0, 1, 0, 0, torch.bfloat16, True, True, ws,
                       cache=cache)


n_warm = int(sys.argv[1]) if len(sys.argv) > 1 else 250
for _ in range(n_warm):
    sums, row_stats, ws = fwd()
    bwd(row_stats, ws)
out = []
for rep in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    sums, row_stats, ws = fwd()
    e[1].record()
    lib.kd_fused_bwd_trace_begin()
    bwd(row_stats, ws)
    e[2].record()
    buf = (ctypes.c_float * (4 * 256))()
    n = lib.kd_fused_bwd_trace_read(ctypes.cast(buf, ctypes.c_void_p), 256)
    recs = [(int(buf[4 * i]), int(buf[4 * i + 1]), buf[4 * i + 2], buf[4 * i + 3]) for i in range(n)]
    tot = {}
    for c, ch, a, b_ in recs:
        tot.setdefault(NAMES[c], []).append(b_ - a)
    span = max(r[3] for r in recs) - min(r[2] for r in recs)
    summary = {k: {"n": len(v), "sum_ms": sum(v), "mean_us": 1e3 * sum(v) / len(v)} for k, v in tot.items()}
    out.append({"fwd_ms": e[0].elapsed_time(e[1]), "bwd_ms": e[1].elapsed_time(e[2]), "bwd_kernel_span_ms": span,
                "classes": summary})
    if rep == 2:
        for c, ch, a, b_ in sorted(recs, key=lambda r: r[2]):
            print(f"{NAMES[c]:15s} chunk {ch:3d}  {a:8.3f} -> {b_:8.3f}  ({1e3 * (b_ - a):7.1f} us)")
    # keep the loop hot between traced steps
    for _ in range(20):
        sums, row_stats, ws = fwd()
        bwd(row_stats, ws)
for o in out:
    print(json.dumps(o))
