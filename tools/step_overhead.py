"""Where does the step time go beyond the two C calls?  Device time (CUDA events) and host enqueue time per step for
(1) the two C-ABI calls back to back, (2) value_and_grad (no autograd), (3) the autograd path bench.py times."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
from speech_distill_b200 import loss as KL
B, T, H, V = 8, 512, 1024, 152936
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
h = torch.randn(B, T, H, device=dev, generator=g).bfloat16().requires_grad_(True)
W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16().requires_grad_(True)
y = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
for b in range(B):
    y[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
labels = torch.randint(0, V, (B, T), device=dev, generator=g)
row_target, n_valid = KL.prepare_rows(labels, None, B, T, -100, dev)
coef = torch.tensor([0.5, 0.5], dtype=torch.float32, device=dev)
h2, y2, Wd = h.detach().reshape(B * T, H), y.reshape(B * T, V), W.detach()


def c_calls():
    sums, row_stats, ws = KL._fused_forward(h2, Wd, y2, row_target, 2.0, 0.5, 0)
    KL._fused_backward(h2, Wd, y2, row_target, row_stats, n_valid, coef, 2.0, 1, 0, 0, torch.bfloat16, True, True, ws)


def vag():
    K.fused_linear_kd_value_and_grad(h.detach(), Wd, labels, teacher_logits=y, grad_dtype=torch.bfloat16)


def autograd():
    h.grad = None
    W.grad = None
    out = K.fused_linear_kd_loss(h, W, labels, teacher_logits=y)
    out[0].backward()


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host * 1e3


for name, fn in (("C calls", c_calls), ("value_and_grad", vag), ("autograd", autograd), ("C calls", c_calls), ("autograd", autograd)):
    d, hst = timeit(fn)
    print(f"{name:15s}: device {d*1e3:7.0f} us/step, host enqueue {hst*1e3:7.0f} us/step")
