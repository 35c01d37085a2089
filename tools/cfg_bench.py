"""BASELINE configs[2] and configs[3] timings (device-resident, CUDA events):
  configs[2]: SoulX-1.7B teacher head (hidden 2048) -> top-64 cache for B=16, T=512, then sparse KD fwd+bwd of the
              Qwen3-0.6B student head on the same tokens;
  configs[3]: stage-1 CE with the frozen-vocabulary mask, B=8, T=2048, 1,000 new rows."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
V, H, Ht = 152936, 1024, 2048
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


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
    return e0.elapsed_time(e1) / n * 1e-3


B, T = 16, 512
R = B * T
ht = torch.randn(B, T, Ht, device=dev, generator=g).bfloat16()
Wt = (torch.randn(V, Ht, device=dev, generator=g) * (2.5 / Ht ** 0.5)).bfloat16()
h = torch.randn(B, T, H, device=dev, generator=g).bfloat16().requires_grad_(True)
W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16().requires_grad_(True)
labels = torch.randint(0, V, (B, T), device=dev, generator=g)
t_head = timeit(lambda: K.teacher_head_topk(ht, Wt, 64))
fl_head = 2.0 * R * Ht * V
tv, ti = K.teacher_head_topk(ht, Wt, 64)
t_gemm = timeit(lambda: K.linear_bf16(ht.reshape(R, Ht), Wt))
print(f"configs[2] teacher head -> top-64 (R={R}, H_t={Ht}): {t_head*1e3:.2f} ms = {fl_head/t_head/1e12:.0f} TFLOP/s "
      f"(head GEMM alone into a full [R,V] buffer: {t_gemm*1e3:.2f} ms = {fl_head/t_gemm/1e12:.0f} TFLOP/s)")


def sparse_step():
    h.grad = None
    W.grad = None
    out = K.fused_linear_kd_loss(h, W, labels, teacher_top_k_v=tv, teacher_top_k_i=ti)
    out[0].backward()


t_sp = timeit(sparse_step)
print(f"configs[2] sparse KD fwd+bwd (R={R}, K=64): {t_sp*1e3:.2f} ms = {R/t_sp/1e6:.3f} M tokens/s, "
      f"{6.0*R*H*V/t_sp/1e12:.0f} TFLOP/s algorithmic")
del ht, Wt, tv, ti

B, T, new = 8, 2048, 1000
R = B * T
h4 = torch.randn(B, T, H, device=dev, generator=g).bfloat16().requires_grad_(True)
labels4 = torch.randint(V - 1500, V, (B, T), device=dev, generator=g)


def ce_step():
    h4.grad = None
    W.grad = None
    loss = K.fused_linear_cross_entropy(h4, W, labels4, old_vocab_size=V - new)
    loss.backward()


t_ce = timeit(ce_step)
alg = (4.0 * R * H * V + 2.0 * R * H * new)
print(f"configs[3] stage-1 CE, frozen vocabulary (R={R}, 1000 live dW rows): {t_ce*1e3:.2f} ms = {R/t_ce/1e6:.3f} M tokens/s, "
      f"{alg/t_ce/1e12:.0f} TFLOP/s algorithmic (4 R H V + 2 R H V_new)")
