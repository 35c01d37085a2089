"""A few K2 launches (dense teacher, forward + dlogits) at the configs[1] shape for ncu captures."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
B, T, V = 8, 512, 152936
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
z = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
y = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
for b in range(B):
    z[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
    y[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
labels = torch.randint(0, V, (B, T), device=dev, generator=g)
z.requires_grad_(True)
for _ in range(2):
    z.grad = None
    out = K.kd_loss_on_logits(z, labels, teacher_logits=y)
    out[0].backward()
torch.cuda.synchronize()
print([float(o.detach()) for o in out])
