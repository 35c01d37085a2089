"""Step time of K1 at BASELINE configs[1] size with the collator's mask pattern (text prefix + padded tail not
scored), with and without valid-row compaction, for the dense teacher, the top-k cache and plain CE."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
B, T, H, V = 8, 512, 1024, 152936
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
h = torch.randn(B, T, H, device=dev, generator=g).bfloat16().requires_grad_(True)
W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16().requires_grad_(True)
y = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
for b in range(B):
    y[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
tv, ti = K.teacher_topk_logprobs(y, 64)


def timeit(fn, n=15):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for frac in (0.0, 0.25, 0.5):
    labels = torch.randint(0, V, (B, T), device=dev, generator=g)
    labels[:, : int(T * frac * 0.8)] = -100
    if frac > 0:
        labels[:, T - int(T * frac * 0.2):] = -100
    n_valid = int((labels[:, 1:] != -100).sum())
    for name, kw in (("dense", dict(teacher_logits=y)), ("top-k cache", dict(teacher_top_k_v=tv, teacher_top_k_i=ti)),
                     ("CE only", dict())):
        row = []
        for compact in (False, True):
            def step():
                h.grad = None
                W.grad = None
                out = K.fused_linear_kd_loss(h, W, labels, compact_rows=compact, **kw)
                out[0].backward()
            row.append(timeit(step))
        print(f"ignored {frac:4.0%} ({n_valid} of {B*T} rows scored) {name:12s}: plain {row[0]*1e3:6.0f} us, "
              f"compacted {row[1]*1e3:6.0f} us ({row[0]/row[1]:.2f}x)")
