"""Achieved HBM bandwidth of the streaming kernels against the algorithmic bytes of SURVEY.md 8(d):
K2 dense fwd+bwd (6 B / element for bf16), K2 forward only (4 B), K2 sparse (4 B V + 6 B K per row),
K3 top-k compaction (2 B R V read + 6 B R K write).  CUDA events, inputs much larger than L2."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
V = 152936
dev = "cuda"
peak = 6544.3
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p)).get("hbm_gbs", peak)


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


def mk(B, T):
    g = torch.Generator(device=dev).manual_seed(B * T)
    z = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
    y = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
    for b in range(B):
        z[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
        y[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
    labels = torch.randint(0, V, (B, T), device=dev, generator=g)
    return z, y, labels


out = []
for B, T, tag in ((2, 512, "configs[0] shape"), (8, 512, "configs[1] shape")):
    z, y, labels = mk(B, T)
    rows = B * (T - 1)  # the last position of a sequence is never scored: zero-filled without reads
    zr = z.clone().requires_grad_(True)

    def fwd_bwd():
        zr.grad = None
        o = K.kd_loss_on_logits(zr, labels, teacher_logits=y)
        o[0].backward()

    t = timeit(fwd_bwd)
    bytes_ = 6.0 * rows * V + 2.0 * B * V  # + zero rows written
    out.append((f"K2 dense fwd+bwd {tag} B={B} T={T}", t, bytes_))
    with torch.no_grad():
        t = timeit(lambda: K.kd_loss_on_logits(z, labels, teacher_logits=y))
    out.append((f"K2 dense forward only {tag}", t, 4.0 * rows * V))
    tv, ti = K.teacher_topk_logprobs(y, 64)

    def sp():
        zr.grad = None
        o = K.kd_loss_on_logits(zr, labels, teacher_top_k_v=tv, teacher_top_k_i=ti)
        o[0].backward()

    t = timeit(sp)
    out.append((f"K2 sparse K=64 fwd+bwd {tag}", t, rows * (4.0 * V + 6.0 * 64) + 2.0 * B * V))
    del z, y, zr
R, k = 16 * 512, 64
x = torch.empty(R, V, device=dev, dtype=torch.bfloat16)
g = torch.Generator(device=dev).manual_seed(1)
for r0 in range(0, R, 1024):
    x[r0:r0 + 1024] = (torch.randn(1024, V, device=dev, generator=g) * 2).bfloat16()
t = timeit(lambda: K.teacher_topk_logprobs(x, k))
out.append((f"K3 top-{k} compaction configs[2] shape R={R}", t, 2.0 * R * V + 6.0 * R * k))
for name, t, b in out:
    print(f"{name:55s} {t*1e6:8.0f} us  {b/t/1e9:7.0f} GB/s algorithmic = {b/t/1e9/peak:.2f} of measured HBM peak ({peak:.0f})")
