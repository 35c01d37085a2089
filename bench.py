#!/usr/bin/env python
"""bench.py - KD fwd+bwd tokens/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (CUDA, sm_100a kernels)
    python bench.py --impl reference [...]                        # reference arm: CPU port of the path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # N > 1

Workload (BASELINE.json configs[1]): Qwen3-0.6B student LM head (hidden 1024) fused with the
KD loss, B=8 T=512 per GPU, V=152,936, bf16, dense full-vocab teacher logits, tau=2 alpha=0.5,
forward + backward to dHidden and dWeight.  A "step" = one such pass over one batch of synthetic
input.  N > 1 = token-shard data parallel (weak scaling): global valid-row count + 8-float loss
record + dW all-reduced over NCCL inside the timed step.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B, T, H, V = 8, 512, 1024, 152936
TAU, ALPHA = 2.0, 0.5
METRIC = "kd_fwd_bwd_tokens_per_sec"
WORKLOAD = ("Qwen3-0.6B student fused lm_head+KD fwd+bwd: B=8 T=512 H=1024 V=152936 bf16, "
            "dense full-vocab bf16 teacher, tau=2 alpha=0.5 (BASELINE.json configs[1])")


def config(n_gpus):
    return {
        "workload": WORKLOAD, "B_per_gpu": B, "T": T, "H": H, "V": V, "tau": TAU, "alpha": ALPHA,
        "tokens_per_step_per_gpu": B * T, "global_batch": B * n_gpus,
        "parallelism": (f"token-shard dp{n_gpus}, dW all-reduce (NCCL) "
                        f"{os.environ.get('KD_BENCH_SYNC', 'overlap')} with the backward") if n_gpus > 1 else "single gpu",
        "l2": "inputs larger than L2 every step (teacher logits 1.25 GB + lm_head 313 MB vs 126 MB L2)",
    }


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference path on host cores
# ----------------------------------------------------------------------------------------------
def cpu_reference(sample_b, sample_t, iters, warmup):
    """LM head (F.linear) + reference loss, forward + backward, fp32 on every host core.
    This is the only place bench.py executes oracle/ (as the baseline being timed, never shipped)."""
    import torch

    from oracle import kd_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    h = torch.randn(sample_b, sample_t, H, generator=g).bfloat16().float()
    W = (torch.randn(V, H, generator=g) * (2.0 / H ** 0.5)).bfloat16().float()
    y = (torch.randn(sample_b, sample_t, V, generator=g) * 2).bfloat16().float()
    labels = torch.randint(0, V, (sample_b, sample_t), generator=g)
    times = []
    for it in range(warmup + iters):
        t0 = time.perf_counter()
        O.fused_linear_reference(h, W, labels, teacher_logits=y, temperature=TAU, alpha=ALPHA)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    best = min(times)
    mean = sum(times) / len(times)
    toks = sample_b * sample_t
    return {
        "value": toks / mean, "best": toks / best, "unit": "tokens/s", "cores": torch.get_num_threads(),
        "kind": "port",
        "sample": f"B={sample_b} T={sample_t} of the same workload (H={H}, V={V}), fp32 torch-CPU restatement "
                  f"of lm_head + distillation_loss.py fwd+bwd, mean of {iters} iters after {warmup} warm-up",
        "s_per_iter": mean,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    cb = cpu_reference(1, 256, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": cb["s_per_iter"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(args.gpus),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import speech_distill_b200 as K
    from speech_distill_b200 import dist as KD

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    K.load_library()

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    gw = torch.Generator(device=dev).manual_seed(99)  # identical weights on every rank
    h = torch.randn(B, T, H, device=dev, generator=g).bfloat16().requires_grad_(True)
    W = (torch.randn(V, H, device=dev, generator=gw) * (2.0 / H ** 0.5)).bfloat16().requires_grad_(True)
    y = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
    for b in range(B):  # chunked to bound the fp32 temporary
        y[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
    labels = torch.randint(0, V, (B, T), device=dev, generator=g)
    reduce_fn, count_fn = KD.make_reduce_fns()
    # dW all-reduce: overlapped with the backward, range by range (default), or one call after it
    sync_mode = os.environ.get("KD_BENCH_SYNC", "overlap") if world > 1 else "none"
    sync = None
    if sync_mode == "overlap":
        max_ctas = int(os.environ.get("KD_BENCH_NCCL_CTAS", "32"))
        sync = KD.GradSync(group=KD.GradSync.new_group(max_ctas), n_ranges=int(os.environ.get("KD_BENCH_RANGES", "6")),
                           max_ctas=max_ctas, reserve_sms=os.environ.get("KD_BENCH_SM_LIMIT", "0") == "1")

    def step(hh, yy, ll):
        out = K.fused_linear_kd_loss(hh, W, ll, teacher_logits=yy, temperature=TAU, alpha=ALPHA,
                                     reduce_fn=reduce_fn, count_reduce_fn=count_fn, grad_sync=sync)
        out[0].backward()
        if sync_mode == "serial":
            KD.allreduce_grad_(W.grad)
        return out

    def clear():
        h.grad = None
        W.grad = None

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    for _ in range(max(3, args.warmup)):
        clear()
        out = step(h, y, labels)
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        clear()
        out = step(h, y, labels)
    e1.record()
    sync_all()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    losses = [float(o) for o in out]
    if os.environ.get("KD_BENCH_QUICK") == "1":  # tuning sweeps: device-resident step time only
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.destroy_process_group()
        if rank == 0:
            print(json.dumps({"quick": True, "n_gpus": world, "ms_per_step": float(t[0]),
                              "value": B * T * world / (float(t[0]) * 1e-3), "sync": sync_mode}), flush=True)
        return

    # ---- per-phase timing for the roofline (forward kernel = the largest single launch) ----
    # events are recorded around the C-ABI calls on the launching stream with NO host sync inside the loop,
    # so the intervals are device time of the launches between them, not Python latency
    from speech_distill_b200 import loss as KL

    n_ph = min(10, args.steps)
    h2 = h.detach().reshape(B * T, H)
    y2 = y.reshape(B * T, V)
    Wd = W.detach()
    row_target, n_valid = KL.prepare_rows(labels, None, B, T, -100, dev)
    coef = torch.tensor([ALPHA, 1.0 - ALPHA], dtype=torch.float32, device=dev)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_ph)]
    torch.cuda.synchronize()
    for i in range(n_ph):
        evs[i][0].record()
        sums, row_stats, ws = KL._fused_forward(h2, Wd, y2, row_target, TAU, ALPHA, 0)
        evs[i][1].record()
        KL._fused_backward(h2, Wd, y2, row_target, row_stats, n_valid, coef, TAU, 1, 0, 0, torch.bfloat16, True, True, ws)
        evs[i][2].record()
    torch.cuda.synchronize()
    fwd_t = statistics.median(e[0].elapsed_time(e[1]) for e in evs)
    bwd_t = statistics.median(e[1].elapsed_time(e[2]) for e in evs)

    # ---- end to end through the public API with HOST buffers ----
    # every step's inputs start in pinned host memory; HostPrefetcher (speech_distill_b200.io) copies step i + 1
    # on its own stream while step i computes; the loss 4-tuple is read back to the host every step
    from speech_distill_b200.io import HostPrefetcher

    h_host = h.detach().cpu().pin_memory()
    y_host = torch.empty((B, T, V), dtype=torch.bfloat16).pin_memory()
    y_host.copy_(y)
    l_host = labels.cpu().pin_memory()
    host_batch = (h_host, y_host, l_host)
    out_host = torch.empty(4, dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 8))
    pf = HostPrefetcher(dev)

    def e2e_step(following):
        W.grad = None
        h_dev, y_dev, l_dev = pf.next(following)
        hh = h_dev.detach().requires_grad_(True)  # fresh leaf over the staging buffer
        o = step(hh, y_dev, l_dev)
        pf.release()
        out_host.copy_(torch.stack([x.detach().float() for x in o]), non_blocking=True)

    pf.submit(host_batch)
    e2e_step(host_batch)      # warm-up (allocates both staging sets); leaves one batch in flight
    e2e_step(host_batch)
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(e2e_steps):
        e2e_step(host_batch)  # K steps computed, K batches copied inside the timed region
    f1.record()
    sync_all()
    e2e_ms = f0.elapsed_time(f1) / e2e_steps
    e2e_losses = out_host.tolist()

    # ---- library GEMM of the same shape, for context: cuBLAS bf16 [B*T, H] x [H, V] -> bf16 logits ----
    cublas_tf = None
    if rank == 0:
        try:
            logits_buf = torch.empty(B * T, V, device=dev, dtype=torch.bfloat16)
            for _ in range(3):
                torch.matmul(h2, Wd.t(), out=logits_buf)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(10):
                torch.matmul(h2, Wd.t(), out=logits_buf)
            c1.record()
            torch.cuda.synchronize()
            cublas_tf = 2.0 * B * T * H * V / (c0.elapsed_time(c1) / 10 * 1e-3) / 1e12
            del logits_buf
        except Exception:  # context only
            cublas_tf = None

    # ---- max over ranks ----
    t = torch.tensor([ms, e2e_ms, fwd_t, bwd_t], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, fwd_t, bwd_t = [float(x) for x in t]

    if rank == 0:
        burst, sustained, hbm, src = peaks()
        tokens = B * T * world
        flops_fwd = 2.0 * B * T * H * V
        chunks = -(-V // 18944)  # kDefaultVChunk
        # prepare_rows + finalize, fwd (gemm, merge, reduce), bwd 3 GEMMs / chunk, fp16 operand copies (h once, W per chunk)
        g16 = os.environ.get("KD_G_FP16", "1") != "0"
        launches_per_step = 2 + 3 + 3 * chunks + ((1 + chunks) if g16 else 0)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("kd_umma_kernel_fwd_dram_bytes_per_launch")
        achieved = flops_fwd / (fwd_t * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": tokens / (ms * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config(world), "clocks": clocks,
            "e2e": {"value": tokens / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h_host.numel() * 2 + y_host.numel() * 2 + l_host.numel() * 8,
                    "d2h_bytes_per_step": 16,
                    "note": "pinned host h, teacher logits and labels copied to the device every step by a "
                            "double-buffered prefetcher (copy of step i+1 overlaps compute of step i; PCIe bound: "
                            "1.25 GB of teacher logits per step); loss 4-tuple read back every step",
                    "losses": e2e_losses},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {
                "kernel": "kd_umma_kernel<FwdEpi> (fused lm_head GEMM + online softmax statistics, forward)",
                "bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                "frac": achieved / sustained,
                "peak_source": f"{src} bf16_tflops_sustained: the launch is timed by CUDA events inside a loop of "
                               f"back-to-back forward+backward steps (power-capped regime), not alone",
                "frac_of_burst_peak": achieved / burst, "burst_peak": burst,
                "cublas_same_shape_tflops": cublas_tf,
                "frac_of_cublas_same_shape": (achieved / cublas_tf) if cublas_tf else None,
                "traffic": traffic, "flops_per_launch": flops_fwd, "ms_per_launch": fwd_t,
                "note": "launch duration = CUDA events around the kd_fused_linear_fwd C call, no host sync in the loop "
                        "(tcgen05 GEMM kernel + the row-merge and reduce kernels, ~25 us)",
            },
            "step_breakdown": {
                "fwd_ms": fwd_t, "bwd_ms": bwd_t,
                "algorithmic_tflops_step": 6.0 * B * T * H * V / (ms / world * 1e-3) / 1e12 if world == 1 else
                6.0 * B * T * H * V / (ms * 1e-3) / 1e12,
                "frac_of_sustained_peak_step": 6.0 * B * T * H * V / (ms * 1e-3) / 1e12 / sustained,
                "executed_flops_factor": "8/6 (backward recomputes the logits tile instead of storing [R,V] logits)",
            },
            "losses": losses,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(1, 256, 3, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
