"""Per-kernel numbers for every BASELINE config besides the headline one (bench.py's `kernels` block, rank 0, N = 1).

Each entry: CUDA-event time of the call through the public API (device-resident inputs much larger than L2, three
warm-ups, a loop of about `seconds` each), the algorithmic bytes or FLOPs of SURVEY.md 8(d) / BASELINE.md 4, the
achieved rate and its fraction of the measured peak (HBM copy bandwidth for the streaming kernels; for the
tensor-bound steps both the sustained and the burst cuBLAS figure - these loops are short, so the burst one applies).

    python tools/bench_kernels.py            # prints the block as JSON
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

V, H, HT = 152936, 1024, 2048


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0)


def _time(fn, seconds=0.3, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    n = max(5, min(400, int(seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3))))
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3, n


def _hbm(name, t, n, nbytes, what, peak):
    return {"kernel": name, "bound": "hbm", "us": t * 1e6, "iters": n, "algorithmic_bytes": nbytes, "bytes_are": what,
            "achieved": nbytes / t / 1e9, "peak": peak, "unit": "GB/s", "frac": nbytes / t / 1e9 / peak}


def _tensor(name, t, n, flops, what, tokens, burst, sustained):
    tf = flops / t / 1e12
    return {"kernel": name, "bound": "tensor", "ms": t * 1e3, "iters": n, "algorithmic_flops": flops, "flops_are": what,
            "achieved": tf, "unit": "TFLOP/s", "peak": burst, "frac": tf / burst, "frac_of_sustained_peak": tf / sustained,
            "tokens_per_s": tokens / t}


def _logits(B, T, dev, seed, scale=2.0):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
    for b in range(B):
        x[b] = (torch.randn(T, V, device=dev, generator=g) * scale).bfloat16()
    return x


def run(seconds=0.3):
    import speech_distill_b200 as K

    dev = torch.device("cuda")
    hbm, burst, sustained = _peaks()
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)

    # ---- K2: DistillationLoss on materialised logits (what an unmodified train.py:97-104 executes) ----
    for B, T, tag in ((2, 512, "configs[0] shape B=2 T=512"), (8, 512, "configs[1] shape B=8 T=512")):
        z, y = _logits(B, T, dev, 10 + B), _logits(B, T, dev, 20 + B)
        labels = torch.randint(0, V, (B, T), device=dev, generator=g)
        rows = B * (T - 1)  # the last position of a sequence is never scored: zero-filled without reads
        zr = z.clone().requires_grad_(True)

        def fwd_bwd():
            zr.grad = None
            o = K.kd_loss_on_logits(zr, labels, teacher_logits=y)
            o[0].backward()

        t, n = _time(fwd_bwd, seconds)
        out[f"k2_dense_fwd_bwd {tag}"] = _hbm("kd_stream_row_kernel (dense teacher, forward + dlogits)", t, n,
                                              6.0 * rows * V + 2.0 * B * V, "read z + read y + write dz (bf16)", hbm)
        with torch.no_grad():
            t, n = _time(lambda: K.kd_loss_on_logits(z, labels, teacher_logits=y), seconds)
        out[f"k2_dense_fwd {tag}"] = _hbm("kd_stream_row_kernel (dense teacher, forward only)", t, n, 4.0 * rows * V,
                                          "read z + read y (bf16)", hbm)
        tv, ti = K.teacher_topk_logprobs(y, 64)

        def sparse():
            zr.grad = None
            o = K.kd_loss_on_logits(zr, labels, teacher_top_k_v=tv, teacher_top_k_i=ti)
            o[0].backward()

        t, n = _time(sparse, seconds)
        out[f"k2_sparse_k64_fwd_bwd {tag}"] = _hbm("kd_stream_row_kernel (top-k cache, forward + dlogits)", t, n,
                                                   rows * (4.0 * V + 6.0 * 64) + 2.0 * B * V,
                                                   "read z + write dz (bf16) + top-k entries", hbm)
        del z, y, zr, tv, ti
    torch.cuda.empty_cache()

    # ---- K3: top-64 compaction of materialised teacher logits (extract_teacher_logits.py:114-129), configs[2] ----
    R, k = 16 * 512, 64
    x = _logits(16, 512, dev, 1).reshape(R, V)
    t, n = _time(lambda: K.teacher_topk_logprobs(x, k), seconds)
    out["k3_topk64 configs[2] shape R=8192"] = _hbm("kd_topk_warp_kernel (warp-per-row form: sweep + selection)", t, n,
                                                    2.0 * R * V + 6.0 * R * k,
                                                    "read logits (bf16) + write fp16 values and int32 indices", hbm)
    xs = x[:1024]
    t, n = _time(lambda: K.teacher_topk_logprobs(xs, k), seconds)
    out["k3_topk64 R=1024 (configs[0] row count)"] = _hbm(
        "kd_topk_stats_kernel + kd_head_select_kernel (two-kernel form; the rows fit the 126 MB L2 only in part)", t, n,
        2.0 * 1024 * V + 6.0 * 1024 * k, "read logits (bf16) + write fp16 values and int32 indices", hbm)
    del x, xs
    torch.cuda.empty_cache()

    # ---- configs[2]: teacher head -> top-64 without [R,V] logits, then the sparse K1 step on the same tokens ----
    B, T = 16, 512
    ht = torch.randn(B, T, HT, device=dev, generator=g).bfloat16()
    Wt = (torch.randn(V, HT, device=dev, generator=g) * (2.5 / HT ** 0.5)).bfloat16()
    h = torch.randn(B, T, H, device=dev, generator=g).bfloat16().requires_grad_(True)
    W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16().requires_grad_(True)
    labels = torch.randint(0, V, (B, T), device=dev, generator=g)
    t, n = _time(lambda: K.teacher_head_topk(ht, Wt, 64), seconds)
    e = _tensor("teacher head GEMM + top-64 compaction (teacher_head_topk)", t, n, 2.0 * R * HT * V,
                "2 R H_t V (SoulX-1.7B head, hidden 2048)", R, burst, sustained)
    tg, ng = _time(lambda: K.linear_bf16(ht.reshape(R, HT), Wt), seconds)
    e["head_gemm_alone_ms"] = tg * 1e3
    e["over_head_gemm_alone"] = t / tg
    out["teacher_head_topk64 configs[2]"] = e
    tv, ti = K.teacher_head_topk(ht, Wt, 64)
    del ht, Wt
    torch.cuda.empty_cache()

    def sparse_step():
        h.grad = None
        W.grad = None
        o = K.fused_linear_kd_loss(h, W, labels, teacher_top_k_v=tv, teacher_top_k_i=ti)
        o[0].backward()

    t, n = _time(sparse_step, seconds)
    out["k1_sparse_step configs[2]"] = _tensor("K1 fwd+bwd, top-k cache teacher (K=64), B=16 T=512", t, n,
                                               6.0 * R * H * V, "6 R H V", R, burst, sustained)
    del h, tv, ti
    torch.cuda.empty_cache()

    # ---- configs[3]: stage-1 CE with the frozen-vocabulary mask, B=8 T=2048, 1,000 new rows ----
    B, T, new = 8, 2048, 1000
    R = B * T
    h4 = torch.randn(B, T, H, device=dev, generator=g).bfloat16().requires_grad_(True)
    labels4 = torch.randint(V - 1500, V, (B, T), device=dev, generator=g)

    def ce_step():
        h4.grad = None
        W.grad = None
        loss = K.fused_linear_cross_entropy(h4, W, labels4, old_vocab_size=V - new)
        loss.backward()

    t, n = _time(ce_step, seconds)
    out["k1_stage1_ce_step configs[3]"] = _tensor("K1 CE fwd+bwd, dW of the 1,000 new rows only, B=8 T=2048", t, n,
                                                  4.0 * R * H * V + 2.0 * R * H * new, "4 R H V + 2 R H V_new", R,
                                                  burst, sustained)
    del h4, W
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    print(json.dumps(run(float(sys.argv[1]) if len(sys.argv) > 1 else 0.3), indent=1))
