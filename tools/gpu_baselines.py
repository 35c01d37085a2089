"""GPU baselines on the same box: what the reference executes on a GPU today.

  (i)  eager : bf16 `F.linear(h, W)` (the LM head, train.py:54) + the reference DistillationLoss
               (distillation_loss.py:14-128; the UNMODIFIED class from oracle/_ref when the recipe has run, else the
               oracle's op-for-op restatement), forward + backward to dH / dW, at BASELINE configs[1];
  (ii) Liger : Liger-Kernel's fused-linear-cross-entropy (what stage1.py:315 `use_liger_kernel=True` asks TRL for;
               third-party Triton + cuBLAS, liger_kernel 0.8.0 in this image) at BASELINE configs[3], plus the
               stage-1 clone-and-zero hook of stage1.py:53-57 on the 313 MB dW.
Each with CUDA-event time per step, tokens/s and peak allocated memory; our K1 beside them on the same inputs.
Prints one JSON object (also importable: bench.py calls run() on rank 0 at N = 1).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _time(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (torch.cuda.max_memory_allocated() - base) / 2 ** 20


def _inputs(B, T, H, V, dev, teacher=True, seed=1234):
    g = torch.Generator(device=dev).manual_seed(seed)
    h = torch.randn(B, T, H, device=dev, generator=g).bfloat16().requires_grad_(True)
    W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16().requires_grad_(True)
    y = None
    if teacher:
        y = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
        for b in range(B):
            y[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
    labels = torch.randint(0, V, (B, T), device=dev, generator=g)
    return h, W, y, labels


def eager_cfg1(iters=5, allow_port=True, B=8, T=512, H=1024, V=152936):
    import speech_distill_b200 as K
    from oracle.make_ref import load_reference_module

    dev = torch.device("cuda")
    ref_mod = load_reference_module()
    if ref_mod is None and not allow_port:
        return {"unavailable": "oracle/_ref/distillation_loss.py absent (run oracle/make_ref.py where the reference is mounted)"}
    h, W, y, labels = _inputs(B, T, H, V, dev)
    if ref_mod is not None:
        loss_fn, kind = ref_mod.DistillationLoss(temperature=2.0, alpha=0.5), "unmodified reference class (oracle/_ref)"
    else:
        from oracle import kd_oracle as O

        def loss_fn(student_logits, labels, teacher_logits):
            return O.reference_loss(student_logits, labels, teacher_logits=teacher_logits, temperature=2.0, alpha=0.5)
        kind = "oracle restatement (oracle/_ref absent)"

    def eager():
        h.grad = W.grad = None
        out = loss_fn(student_logits=torch.nn.functional.linear(h, W), labels=labels, teacher_logits=y)
        out[0].backward()
        return out

    def ours():
        h.grad = W.grad = None
        out = K.fused_linear_kd_loss(h, W, labels, teacher_logits=y, temperature=2.0, alpha=0.5)
        out[0].backward()
        return out

    ms_e, mem_e = _time(eager, iters)
    le = [float(x.detach()) for x in eager()]
    ms_o, mem_o = _time(ours, max(iters, 20), warm=5)
    lo = [float(x.detach()) for x in ours()]
    toks = B * T
    return {
        "config": f"configs[1]: B={B} T={T} H={H} V={V} bf16, dense teacher, fwd+bwd to dH/dW",
        "eager": {"what": "F.linear + " + kind + ", all bf16 (train.py:54 + distillation_loss.py:14-128)",
                  "ms_per_step": ms_e, "tokens_per_s": toks / (ms_e * 1e-3), "peak_mem_mb": mem_e, "losses": le},
        "ours": {"ms_per_step": ms_o, "tokens_per_s": toks / (ms_o * 1e-3), "peak_mem_mb": mem_o, "losses": lo},
        "speedup_over_eager": ms_e / ms_o, "peak_mem_ratio_eager_over_ours": mem_e / max(mem_o, 1e-9),
    }


def liger_cfg3(iters=5, allow_port=True, B=8, T=2048, H=1024, V=152936, new_rows=1000):
    import speech_distill_b200 as K

    dev = torch.device("cuda")
    h, W, _, labels = _inputs(B, T, H, V, dev, teacher=False)
    out = {"config": f"configs[3]: B={B} T={T} H={H} V={V} bf16 CE, {new_rows} trainable rows (stage1.py:29-73)"}
    toks = B * T
    shift = torch.full_like(labels, -100)
    shift[:, :-1] = labels[:, 1:]  # causal shift done by the caller, as transformers' lce_forward does

    def ours():
        h.grad = W.grad = None
        loss = K.fused_linear_cross_entropy(h, W, labels, old_vocab_size=V - new_rows)
        loss.backward()
        return loss

    ms_o, mem_o = _time(ours, max(iters, 10), warm=3)
    out["ours"] = {"ms_per_step": ms_o, "tokens_per_s": toks / (ms_o * 1e-3), "peak_mem_mb": mem_o,
                   "loss": float(ours().detach())}
    try:
        from liger_kernel.transformers import LigerFusedLinearCrossEntropyLoss

        lce = LigerFusedLinearCrossEntropyLoss(ignore_index=-100, reduction="mean")

        def liger():
            h.grad = W.grad = None
            loss = lce(W, h.reshape(B * T, H), shift.reshape(B * T))
            loss.backward()
            g = W.grad.clone()          # stage1.py:53-57: the hook clones the gradient ...
            g[: V - new_rows] = 0       # ... and zeroes the frozen rows
            return loss

        ms_l, mem_l = _time(liger, iters)
        out["liger_flce"] = {"what": "liger_kernel LigerFusedLinearCrossEntropyLoss (Triton + cuBLAS) + stage1's "
                                     "clone-and-zero gradient hook", "ms_per_step": ms_l,
                             "tokens_per_s": toks / (ms_l * 1e-3), "peak_mem_mb": mem_l, "loss": float(liger().detach())}
        out["speedup_over_liger"] = ms_l / ms_o
    except Exception as e:  # the baseline is third-party: report, do not fail
        out["liger_flce"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    return out


def run(iters=5, allow_port=True):
    """allow_port=False (bench.py): only the unmodified reference class is timed as the eager baseline; the oracle's
    restatement is used in its place only when this tool is run by hand without oracle/_ref."""
    res = {}
    for name, fn in (("eager_cfg1", eager_cfg1), ("liger_cfg3", liger_cfg3)):
        try:
            res[name] = fn(iters, allow_port)
        except Exception as e:
            res[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 5), indent=1))
