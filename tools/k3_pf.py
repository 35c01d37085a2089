"""K3 at the configs[2] shape (R=8192, V=152936, k=64): best-of-3 time under this process's KD_TOPK_* knobs (one line)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
V, R, k, dev = 152936, 8192, 64, "cuda"
g = torch.Generator(device=dev).manual_seed(7)
x = torch.empty(R, V, device=dev, dtype=torch.bfloat16)
for r0 in range(0, R, 1024):
    x[r0:r0 + 1024] = (torch.randn(1024, V, device=dev, generator=g) * 2).bfloat16()
for _ in range(3):
    v, i = K.teacher_topk_logprobs(x, k)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    torch.cuda.synchronize()
    e0.record()
    for _ in range(30):
        K.teacher_topk_logprobs(x, k)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 30 * 1e3)
ti = torch.topk(x[:64].float(), k, -1).indices
ok = bool((i[:64].long().sort(-1).values == ti.sort(-1).values).all())
print(f"KD_TOPK_L2_AHEAD={os.environ.get('KD_TOPK_L2_AHEAD', 'default')} K3 {best:.0f} us  {2.0 * R * V / best / 1e3:.0f} GB/s  idx_ok={ok}")
