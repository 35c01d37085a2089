"""One forward + backward of K1 at BASELINE configs[1] size (for ncu captures)."""
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
labels = torch.randint(0, V, (B, T), device=dev, generator=g)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(n):
    h.grad = None; W.grad = None
    out = K.fused_linear_kd_loss(h, W, labels, teacher_logits=y)
    out[0].backward()
torch.cuda.synchronize()
print([float(o) for o in out])
