"""K2 (kd_stream_row_kernel) times under this process's library (KD_B200_LIB): dense fwd+bwd, forward only and
sparse K=64 at the configs[0] / configs[1] shapes; one line per run for interleaved A/B of build.py --variant builds."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
V, dev = 152936, "cuda"


def timeit(fn, n=40):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


res = []
for B, T in ((2, 512), (8, 512)):
    g = torch.Generator(device=dev).manual_seed(B * T)
    z = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
    y = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
    for b in range(B):
        z[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
        y[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
    labels = torch.randint(0, V, (B, T), device=dev, generator=g)
    zr = z.clone().requires_grad_(True)

    def fwd_bwd():
        zr.grad = None
        o = K.kd_loss_on_logits(zr, labels, teacher_logits=y)
        o[0].backward()
        return o

    o = fwd_bwd()
    res.append(f"B{B} fb {timeit(fwd_bwd):.0f}us")
    with torch.no_grad():
        res.append(f"f {timeit(lambda: K.kd_loss_on_logits(z, labels, teacher_logits=y)):.0f}us")
    tv, ti = K.teacher_topk_logprobs(y, 64)

    def sp():
        zr.grad = None
        o = K.kd_loss_on_logits(zr, labels, teacher_top_k_v=tv, teacher_top_k_i=ti)
        o[0].backward()

    res.append(f"sp {timeit(sp):.0f}us")
    res.append("loss " + ",".join(f"{float(v.detach()):.6f}" for v in o) + f" gsum {float(zr.grad.float().abs().sum()):.6e}")
    del z, y, zr
print(os.path.basename(os.environ.get("KD_B200_LIB", "default")), " | ".join(res))
