"""Round-robin A/B of the backward's vocabulary chunk at BASELINE configs[1] size: every candidate is timed in turn,
several rounds, so that clock / thermal drift hits all of them alike."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_distill_b200 import loss as KL
B, T, H, V = 8, 512, 1024, 152936
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
h = torch.randn(B * T, H, device=dev, generator=g).bfloat16()
W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16()
y = torch.empty(B * T, V, device=dev, dtype=torch.bfloat16)
for b in range(B):
    y[b * T:(b + 1) * T] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
labels = torch.randint(0, V, (B, T), device=dev, generator=g)
row_target, n_valid = KL.prepare_rows(labels, None, B, T, -100, dev)
coef = torch.tensor([0.5, 0.5], dtype=torch.float32, device=dev)
cands = [int(c) for c in sys.argv[1:]] or [9472, 14336, 18944, 28416]
ws = {vc: KL._fused_workspace(B * T, H, V, vc, dev) for vc in cands}
sums, row_stats, _ = KL._fused_forward(h, W, y, row_target, 2.0, 0.5, 0)


def step(vc):
    KL._fused_forward(h, W, y, row_target, 2.0, 0.5, vc)
    KL._fused_backward(h, W, y, row_target, row_stats, n_valid, coef, 2.0, 1, 0, vc, torch.bfloat16, True, True, ws[vc])


res = {vc: [] for vc in cands}
for rnd in range(6):
    for vc in cands:
        for _ in range(2):
            step(vc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(8):
            step(vc)
        e1.record()
        torch.cuda.synchronize()
        res[vc].append(e0.elapsed_time(e1) / 8)
for vc in cands:
    r = sorted(res[vc])
    print(f"v_chunk {vc:6d}: median {r[len(r)//2]*1e3:6.0f} us/step, min {r[0]*1e3:6.0f}, max {r[-1]*1e3:6.0f}  (G scratch 2 x {B*T*vc*2/1e6:.0f} MB)")
