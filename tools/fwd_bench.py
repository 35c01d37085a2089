"""Forward-only timing of K1 (dense teacher vs no teacher) without host syncs inside the loop."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
B, T, H, V = 8, 512, 1024, 152936
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
h = torch.randn(B, T, H, device=dev, generator=g).bfloat16()
W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16()
y = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
for b in range(B):
    y[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
labels = torch.randint(0, V, (B, T), device=dev, generator=g)
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    t_dense = timeit(lambda: K.fused_linear_kd_loss(h, W, labels, teacher_logits=y))
    t_none = timeit(lambda: K.fused_linear_kd_loss(h, W, labels, teacher_logits=None, temperature=1.0, alpha=1.0))
fl = 2.0 * B * T * H * V
print(f"skip_math={os.environ.get('KD_DEBUG_SKIP_MATH','0')} fwd dense {t_dense*1e3:.0f} us ({fl/t_dense/1e9:.0f} TF/s) | fwd no-teacher {t_none*1e3:.0f} us ({fl/t_none/1e9:.0f} TF/s)")
