"""Backward time of K1 at BASELINE configs[1] size as a function of the vocabulary chunk (launch count),
forward time, and the whole step; CUDA events on the launching stream, no host sync inside the loops."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
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


def timeit(fn, n=10):
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


sums, row_stats, ws0 = KL._fused_forward(h, W, y, row_target, 2.0, 0.5, 0)
t_f = timeit(lambda: KL._fused_forward(h, W, y, row_target, 2.0, 0.5, 0))
fl = 2.0 * B * T * H * V
print(f"fwd {t_f*1e3:.0f} us ({fl/t_f/1e9:.0f} TF/s)")
chunks = [int(c) for c in sys.argv[1:]] or [4864, 9472, 18944, 37888, 75776, 153088]
for vc in chunks:
    ws = KL._fused_workspace(B * T, H, V, vc, dev)
    for need_h, need_w, tag in ((True, True, "dH+dW"), (True, False, "dH only"), (False, True, "dW only")):
        t = timeit(lambda: KL._fused_backward(h, W, y, row_target, row_stats, n_valid, coef, 2.0, 1, 0, vc,
                                              torch.bfloat16, need_h, need_w, ws))
        n_l = -(-V // vc) * (1 + int(need_h) + int(need_w))
        print(f"v_chunk {vc:6d} {tag:8s}: bwd {t*1e3:7.0f} us, {n_l:3d} launches, G scratch {B*T*vc*2/1e6:.0f} MB")
