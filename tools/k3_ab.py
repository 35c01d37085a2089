"""K3 (kd_topk_logprobs) at the configs[2] shape under this process's library / environment knobs: time, achieved
HBM fraction, and a bit-exactness check of the indices against torch.topk on a tie-free fp32 row block.

    KD_B200_LIB=.../libkd_b200_NAME.so KD_TOPK_FORM=warp python tools/k3_ab.py [R] [k]
"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K

V = 152936
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = "cuda"
peak = 6544.3
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p)).get("hbm_gbs", peak)
g = torch.Generator(device=dev).manual_seed(7)
x = torch.empty(R, V, device=dev, dtype=torch.bfloat16)
for r0 in range(0, R, 1024):
    x[r0:r0 + 1024] = (torch.randn(min(1024, R - r0), V, device=dev, generator=g) * 2).bfloat16()


def run():
    return K.teacher_topk_logprobs(x, k)


for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        run()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 50 * 1e-3)
nbytes = 2.0 * R * V + 6.0 * R * k
# this box's copy bandwidth and a plain read-reduce of the same logits (calibration: boxes of the pool differ)
ca = torch.empty(1 << 29, dtype=torch.bfloat16, device=dev)
cb = torch.empty_like(ca)
for _ in range(2):
    cb.copy_(ca)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    cb.copy_(ca)
e1.record()
torch.cuda.synchronize()
copy_gbs = 2 * ca.numel() * 2 * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
del ca, cb
for _ in range(2):
    x.amax(dim=-1)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    x.amax(dim=-1)
e1.record()
torch.cuda.synchronize()
amax_us = e0.elapsed_time(e1) / 10 * 1e3
# parity: same selected-logit multiset as torch.topk on the bf16 rows, indices bit-exact on a tie-free fp32 block
v, i = run()
xs = x[:256].float()
tv, ti = torch.topk(xs, k, dim=-1)
sel = torch.gather(xs, 1, i[:256].long())
same_vals = bool(torch.equal(sel.sort(dim=-1, descending=True).values, tv))
xf = torch.randn(64, V, device=dev, generator=g) * 3
vf, jf = K.teacher_topk_logprobs(xf, k)
lf = torch.log_softmax(xf, dim=-1)
tvf, tif = torch.topk(lf, k, dim=-1)
print(json.dumps({"lib": os.environ.get("KD_B200_LIB", "default"), "form": os.environ.get("KD_TOPK_FORM", "auto"), "env": {a: b for a, b in os.environ.items() if a.startswith("KD_TOPK")}, "copy_gbs": round(copy_gbs), "torch_amax_us": round(amax_us), "R": R,
                  "k": k, "us": best * 1e6, "gbs": nbytes / best / 1e9, "frac_of_hbm_peak": nbytes / best / 1e9 / peak,
                  "bf16_same_selected_values": same_vals, "fp32_indices_bit_exact": bool(torch.equal(jf.long(), tif)),
                  "fp32_value_max_err": float((vf.float() - tvf).abs().max())}))
