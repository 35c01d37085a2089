#!/usr/bin/env python
"""bench.py - KD fwd+bwd tokens/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (CUDA, sm_100a kernels)
    python bench.py --impl reference [...]                        # reference arm: the reference's CPU path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # N > 1

Workload (BASELINE.json configs[1]): Qwen3-0.6B student LM head (hidden 1024) fused with the
KD loss, B=8 T=512 per GPU, V=152,936, bf16, dense full-vocab teacher logits, tau=2 alpha=0.5,
forward + backward to dHidden and dWeight.  A "step" = one such pass over one batch of synthetic
input.  N > 1 = token-shard data parallel (weak scaling): global valid-row count + 8-float loss
record + dW all-reduced over NCCL inside the timed step.

What the line holds (DESIGN.md 5 spells out every figure):
  value / ms_per_step   K-step blocks timed by CUDA events back to back for >= KD_BENCH_SECONDS (2.5 s): the MEDIAN
                        block (max over ranks per block), i.e. the power-settled regime; the first (burst) block beside it
  roofline              the largest single launch (fused forward GEMM), timed inside a seconds-long loop of steps ->
                        sustained peak; `roofline_step` the whole step; `roofline_kernels` every kernel class of the step
  kernels               K2 / K3 / teacher-head top-k / sparse K1 / stage-1 numbers for the other BASELINE configs
  gpu_baselines         the reference's eager GPU path and Liger FLCE on the same box (tools/gpu_baselines.py)
  e2e, e2e_topk_cache   the step through the public API with pinned HOST inputs (dense teacher; top-k cache path)
  cpu_baseline          the unmodified reference on the host cores (separate process: own thread pool and peak RSS)
  multi_gpu (N > 1)     N-rank result vs a single-process run on the concatenated batch, bf16 vs fp32 dW all-reduce,
                        and the vocab-parallel mode timed in the same run
"""
import argparse
import json
import math
import os
import statistics
import subprocess
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
MIN_SECONDS = float(os.environ.get("KD_BENCH_SECONDS", "2.5"))
NAMES = {0: "cast", 1: "grad_cached", 2: "dW", 3: "dH", 4: "grad_recompute"}  # kd_fused_bwd_trace classes


def config(n_gpus, extra=None):
    c = {
        "workload": WORKLOAD, "B_per_gpu": B, "T": T, "H": H, "V": V, "tau": TAU, "alpha": ALPHA,
        "tokens_per_step_per_gpu": B * T, "global_batch": B * n_gpus,
        "parallelism": (f"token-shard dp{n_gpus}, dW all-reduce (NCCL) "
                        f"{os.environ.get('KD_BENCH_SYNC', 'overlap')} with the backward") if n_gpus > 1 else "single gpu",
        "l2": "inputs larger than L2 every step (teacher logits 1.25 GB + lm_head 313 MB vs 126 MB L2)",
    }
    if extra:
        c.update(extra)
    return c


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation of the path on the host cores
# (oracle/_ref = the unmodified reference module; the oracle port only if that copy is absent).
# This function is the only place bench.py executes anything under oracle/.
# ----------------------------------------------------------------------------------------------
def _cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _reference_loss_fn():
    from oracle.make_ref import load_reference_module

    mod = load_reference_module()
    if mod is not None:
        return mod.DistillationLoss(temperature=TAU, alpha=ALPHA), "reference"
    from oracle import kd_oracle as O

    def port(student_logits, labels, teacher_logits):
        return O.reference_loss(student_logits, labels, teacher_logits=teacher_logits, temperature=TAU, alpha=ALPHA)

    return port, "port"


def cpu_head_and_loss(sample_b, sample_t, iters, warmup):
    """The same workload as our arm on the host: LM head (F.linear, train.py:54) + the reference DistillationLoss
    (distillation_loss.py:14-128), forward + backward to dH and dW, fp32, every host core, on a sample of the batch."""
    import torch

    loss_fn, kind = _reference_loss_fn()
    g = torch.Generator().manual_seed(1234)
    h = torch.randn(sample_b, sample_t, H, generator=g).bfloat16().float().requires_grad_(True)
    W = (torch.randn(V, H, generator=g) * (2.0 / H ** 0.5)).bfloat16().float().requires_grad_(True)
    y = (torch.randn(sample_b, sample_t, V, generator=g) * 2).bfloat16().float()
    labels = torch.randint(0, V, (sample_b, sample_t), generator=g)
    times, losses = [], None
    for it in range(warmup + iters):
        h.grad = W.grad = None
        t0 = time.perf_counter()
        out = loss_fn(student_logits=torch.nn.functional.linear(h, W), labels=labels, teacher_logits=y)
        out[0].backward()
        dt = time.perf_counter() - t0
        losses = [float(o.detach()) for o in out]
        if it >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    toks = sample_b * sample_t
    return {"value": toks / mean, "best": toks / min(times), "unit": "tokens/s", "kind": kind, "s_per_iter": mean,
            "iters": iters, "warmup": warmup, "sample_B": sample_b, "sample_T": sample_t, "losses": losses,
            "sample": f"B={sample_b} T={sample_t} sample of the workload (H={H}, V={V}): F.linear lm_head + "
                      f"{'the unmodified reference DistillationLoss (oracle/_ref)' if kind == 'reference' else 'oracle port of distillation_loss.py'}"
                      f", fwd+bwd to dH/dW, fp32, mean of {iters} iters after {warmup} warm-up"}


def cpu_loss_only(dtype_name, iters, warmup=1):
    """the unmodified DistillationLoss on [2, 512, 152936] logits (configs[0]), forward + backward to
    student_logits.grad, best of `iters` after one warm-up; tokens/s = B (T-1) / t."""
    import torch

    loss_fn, kind = _reference_loss_fn()
    dt_ = torch.float32 if dtype_name == "fp32" else torch.bfloat16
    g = torch.Generator().manual_seed(1234)
    z = (torch.randn(2, 512, V, generator=g) * 2).to(dt_).requires_grad_(True)
    y = (torch.randn(2, 512, V, generator=g) * 2).to(dt_)
    labels = torch.randint(0, V, (2, 512), generator=g)
    times, losses = [], None
    for it in range(warmup + iters):
        z.grad = None
        t0 = time.perf_counter()
        out = loss_fn(student_logits=z, labels=labels, teacher_logits=y)
        out[0].backward()
        dt = time.perf_counter() - t0
        losses = [float(o.detach()) for o in out]
        if it >= warmup:
            times.append(dt)
    best = min(times)
    return {"dtype": dtype_name, "s_per_iter_best": best, "tokens_per_s": 2 * 511 / best, "iters": iters, "kind": kind,
            "losses": losses, "shape": f"student/teacher logits [2, 512, {V}], labels [2, 512], all rows valid"}


def _pick_sample(steps, warmup, budget_s=150.0):
    """Largest sample of the batch whose (steps + warmup) iterations fit the budget (calibrated on B=1, T=128)."""
    c = cpu_head_and_loss(1, 128, 1, 1)
    per_tok = c["s_per_iter"] / 128
    for sb, st in ((2, 512), (1, 512), (1, 256), (1, 128)):
        if per_tok * sb * st * (steps + warmup) <= budget_s:
            return sb, st
    return 1, 64


def cpu_env():
    import resource

    import torch

    return {"cores": torch.get_num_threads(), "os_cpu_count": os.cpu_count(), "cpu_model": _cpu_model(),
            "torch_threads": torch.get_num_threads(),
            "peak_rss_gb": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 2 ** 20}


def run_reference(args):
    """`--impl reference`: rank 0 times the reference's CPU path; with --cpu-baseline-leg it is the bounded
    cpu_baseline run our arm spawns (sample of about 10-30 s of CPU work, plus BASELINE.md 3's loss-only figures)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    if args.cpu_baseline_leg:
        cb = cpu_head_and_loss(2, 512, 3, 1)
        cb["loss_only_configs0"] = [cpu_loss_only("fp32", 5), cpu_loss_only("bf16", 3)]
        cb.update(cpu_env())
        print(json.dumps(cb), flush=True)
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    sb, st = _pick_sample(steps, warm)
    cb = cpu_head_and_loss(sb, st, steps, warm)
    cb.update(cpu_env())
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": cb["s_per_iter"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config(args.gpus, {"reference_sample": f"each step = B={sb} T={st} of the B={B} T={T} batch "
                                                         f"({sb * st} tokens; per-token throughput is what is compared)"}),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "os_cpu_count", "cpu_model",
                                            "peak_rss_gb")},
        "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "losses": cb["losses"],
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess():
    """Runs the cpu_baseline leg in its own process (own OpenMP pool, clean peak RSS, no CUDA context)."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--cpu-baseline-leg"],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600,
                           env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        cb = json.loads(r.stdout.strip().splitlines()[-1])
        return cb
    except Exception as e:  # reported, never fatal for the GPU line
        return {"error": f"{type(e).__name__}: {e}"[:300]}


# ----------------------------------------------------------------------------------------------
# clocks sampler (NVML; ~1-2 ms period so that a 1.3 ms kernel phase is resolved in aggregate)
# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.samples, self.power, self.reasons, self.max_mhz = index, [], [], set(), None
        self.period = period
        self._halt = threading.Event()
        self.ok = False
        self.t0 = self.t1 = None
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
        self.t0 = time.perf_counter()
        i = 0
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                if i % 8 == 0:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                    self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
            except Exception:
                pass
            i += 1
            time.sleep(self.period)
        self.t1 = time.perf_counter()

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        n = len(self.samples)
        tail = self.samples[n // 2:]
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": n, "sample_period_ms": 1e3 * (self.t1 - self.t0) / n if self.t1 else None,
                "sm_mhz_settled": statistics.median(tail), "sm_mhz_first_100ms": statistics.median(self.samples[:max(1, n // 25)]),
                "sm_mhz_min": min(self.samples), "power_w_median": statistics.median(self.power) if self.power else None}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def _guard(fn, *a, **kw):
    try:
        return fn(*a, **kw)
    except Exception as e:  # a sub-block that fails is reported in place; the headline numbers still print
        return {"error": f"{type(e).__name__}: {e}"[:400]}


def run_ours(args):
    import ctypes

    import torch
    import torch.distributed as dist

    import speech_distill_b200 as K
    from speech_distill_b200 import dist as KD
    from speech_distill_b200 import loss as KL

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
    lib = K.load_library()
    quick = os.environ.get("KD_BENCH_QUICK") == "1"
    steps = max(1, args.steps)
    warm = max(3, args.warmup)

    def gen_inputs(seed_rank, Bn=B):
        g = torch.Generator(device=dev).manual_seed(1234 + seed_rank)
        hh = torch.randn(Bn, T, H, device=dev, generator=g).bfloat16()
        yy = torch.empty(Bn, T, V, device=dev, dtype=torch.bfloat16)
        for b in range(Bn):  # chunked to bound the fp32 temporary
            yy[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
        ll = torch.randint(0, V, (Bn, T), device=dev, generator=g)
        return hh, yy, ll

    gw = torch.Generator(device=dev).manual_seed(99)  # identical weights on every rank
    W = (torch.randn(V, H, device=dev, generator=gw) * (2.0 / H ** 0.5)).bfloat16().requires_grad_(True)
    reduce_fn, count_fn = KD.make_reduce_fns()
    # dW all-reduce: overlapped with the backward, range by range (default), or one call after it
    sync_mode = os.environ.get("KD_BENCH_SYNC", "overlap") if world > 1 else "none"
    sync = None
    if sync_mode == "overlap":
        max_ctas = int(os.environ.get("KD_BENCH_NCCL_CTAS", "32"))
        sync = KD.GradSync(group=KD.GradSync.new_group(max_ctas), n_ranges=int(os.environ.get("KD_BENCH_RANGES", "6")),
                           max_ctas=max_ctas, reserve_sms=os.environ.get("KD_BENCH_SM_LIMIT", "0") == "1",
                           backend=os.environ.get("KD_BENCH_BACKEND", "nccl"),
                           multimem_ctas=int(os.environ.get("KD_BENCH_MM_CTAS", "16")))

    def step(hh, yy, ll, grad_sync=sync, reduce=True):
        out = K.fused_linear_kd_loss(hh, W, ll, teacher_logits=yy, temperature=TAU, alpha=ALPHA,
                                     reduce_fn=reduce_fn if reduce else None,
                                     count_reduce_fn=count_fn if reduce else None, grad_sync=grad_sync)
        out[0].backward()
        if sync_mode == "serial" and reduce:
            KD.allreduce_grad_(W.grad)
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    # ---- N > 1: the N-rank result against a single-process run on the concatenated batch (before any timing) ----
    multi = None
    if world > 1 and not quick:
        multi = {"parity": _guard(multi_gpu_parity, K, KD, dist, torch, dev, rank, world, W, reduce_fn, count_fn, sync)}

    h, y, labels = gen_inputs(rank)
    h.requires_grad_(True)

    def clear():
        h.grad = None
        W.grad = None

    # ---- device-resident throughput ("value"): K-step blocks back to back for >= MIN_SECONDS ----
    for _ in range(warm):
        clear()
        out = step(h, y, labels)
    sync_all()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
        clear()
        out = step(h, y, labels)
    c1.record()
    sync_all()
    est_ms = max_over_ranks([c0.elapsed_time(c1) / 3])[0]
    n_blocks = max(1, int(math.ceil(MIN_SECONDS * 1e3 / (est_ms * steps)))) if not quick else 1
    n_blocks = min(n_blocks, 400)
    sampler = ClockSampler(local)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_blocks + 1)]
    sync_all()
    launches0 = lib.kd_launch_count()
    sampler.start()
    ev[0].record()
    for blk in range(n_blocks):
        for _ in range(steps):
            clear()
            out = step(h, y, labels)
        ev[blk + 1].record()
    sync_all()
    clocks = sampler.stop()
    launches_per_block = (lib.kd_launch_count() - launches0) / n_blocks
    block_ms = max_over_ranks([ev[i].elapsed_time(ev[i + 1]) for i in range(n_blocks)])
    ms = statistics.median(block_ms) / steps
    first_ms = block_ms[0] / steps
    loop_s = sum(block_ms) / 1e3
    losses = [float(o.detach()) for o in out]
    if quick:  # tuning sweeps: device-resident step time only
        if world > 1:
            dist.destroy_process_group()
        if rank == 0:
            print(json.dumps({"quick": True, "n_gpus": world, "ms_per_step": ms,
                              "value": B * T * world / (ms * 1e-3), "sync": sync_mode}), flush=True)
        return

    # ---- per-phase timing inside a seconds-long loop (GPU still hot from the loop above): events around the C-ABI
    #      calls on the launching stream, no host sync inside the loop; then one traced backward for the kernel classes
    h2 = h.detach().reshape(B * T, H)
    y2 = y.reshape(B * T, V)
    Wd = W.detach()
    row_target, n_valid = KL.prepare_rows(labels, None, B, T, -100, dev)
    coef = torch.tensor([ALPHA, 1.0 - ALPHA], dtype=torch.float32, device=dev)
    cache = KL.alloc_logit_cache(B * T, V, 0, dev)

    def phase_step(e=None):
        if e:
            e[0].record()
        sums, row_stats, ws = KL._fused_forward(h2, Wd, y2, row_target, TAU, ALPHA, 0, cache=cache)
        if e:
            e[1].record()
        KL._fused_backward(h2, Wd, y2, row_target, row_stats, n_valid, coef, TAU, 1, 0, 0, torch.bfloat16, True, True,
                           ws, cache=cache)
        if e:
            e[2].record()

    n_ph = max(20, int(1.2e3 / ms))
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_ph)]
    ph_sampler = ClockSampler(local)
    ph_sampler.start()
    for i in range(n_ph):
        phase_step(evs[i])
    torch.cuda.synchronize()
    ph_clocks = ph_sampler.stop()
    half = evs[n_ph // 2:]
    fwd_t = statistics.median(e[0].elapsed_time(e[1]) for e in half)
    bwd_t = statistics.median(e[1].elapsed_time(e[2]) for e in half)
    phase_step_ms = evs[n_ph // 2][0].elapsed_time(evs[-1][2]) / (n_ph - n_ph // 2)
    trace = None
    try:
        for _ in range(5):
            phase_step()
        sums, row_stats, ws = KL._fused_forward(h2, Wd, y2, row_target, TAU, ALPHA, 0, cache=cache)
        lib.kd_fused_bwd_trace_begin()
        KL._fused_backward(h2, Wd, y2, row_target, row_stats, n_valid, coef, TAU, 1, 0, 0, torch.bfloat16, True, True,
                           ws, cache=cache)
        buf = (ctypes.c_float * (4 * 256))()
        n = lib.kd_fused_bwd_trace_read(ctypes.cast(buf, ctypes.c_void_p), 256)
        trace = {}
        t_lo, t_hi = 1e30, -1e30
        for i in range(n):
            c, a, b_ = NAMES.get(int(buf[4 * i]), "other"), buf[4 * i + 2], buf[4 * i + 3]
            trace.setdefault(c, []).append(b_ - a)
            t_lo, t_hi = min(t_lo, a), max(t_hi, b_)
        trace = {"span_ms": t_hi - t_lo, "classes": {k: {"n": len(v), "sum_ms": sum(v)} for k, v in trace.items()}}
    except Exception as e:
        trace = {"error": str(e)[:200]}
    del cache
    torch.cuda.empty_cache()

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
    e2e_steps = max(3, min(steps, 8))
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
    e2e_ms = max_over_ranks([f0.elapsed_time(f1) / e2e_steps])[0]
    e2e_losses = out_host.tolist()
    h2d_bytes = h_host.numel() * 2 + y_host.numel() * 2 + l_host.numel() * 8
    del pf, y_host, host_batch
    fwd_t, bwd_t, phase_step_ms = max_over_ranks([fwd_t, bwd_t, phase_step_ms])

    # ---- N > 1: bf16 vs fp32 all-reduce of dW at full size; the vocab-parallel mode in the same run ----
    if multi is not None:
        multi["dW_allreduce_bf16_vs_fp32"] = _guard(allreduce_precision, dist, torch, step, h, y, labels, W, clear)
        del y, h2, y2
        torch.cuda.empty_cache()
        multi["vocab_parallel"] = _guard(vocab_parallel_block, K, dist, torch, dev, rank, world, steps)
    else:
        del y, h2, y2
    torch.cuda.empty_cache()

    extra = {}
    if rank == 0 and world == 1:
        extra["e2e_topk_cache"] = _guard(e2e_topk_cache, K, torch, dev)
        # library GEMM of the same shape, for context: cuBLAS bf16 [B*T, H] x [H, V] -> bf16 logits
        extra["cublas_tf"] = _guard(cublas_same_shape, torch, h.detach().reshape(B * T, H), Wd)
        if not args.no_kernels:
            from tools import bench_kernels, gpu_baselines

            W.grad = None
            torch.cuda.empty_cache()
            extra["kernels"] = _guard(bench_kernels.run)
            extra["gpu_baselines"] = _guard(gpu_baselines.run, 5, False)
            from tools import read_probe

            # what a read-only stream can reach on this box (K2 forward and K3 are read-only; MEASURED_PEAKS' HBM
            # figure is a copy) - context for the `kernels` fractions, which stay against the measured copy peak
            extra["read_probe"] = _guard(read_probe.probe, 1 << 30)
        if not args.no_cpu_baseline:
            torch.cuda.empty_cache()
            extra["cpu_baseline"] = cpu_baseline_subprocess()

    if rank == 0:
        burst, sustained, hbm, src = peaks()
        tokens = B * T * world
        flops_fwd = 2.0 * B * T * H * V
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("kd_umma_kernel_fwd_dram_bytes_per_launch")
        achieved = flops_fwd / (fwd_t * 1e-3) / 1e12
        cublas_tf = extra.get("cublas_tf") if isinstance(extra.get("cublas_tf"), float) else None
        step_tf = 6.0 * B * T * H * V / (ms * 1e-3) / 1e12
        # kernel classes of the backward (kd_fused_bwd_trace: start / end of every kernel on its stream, one step).
        # The three chains run concurrently and share the SMs, so the class durations overlap (their sum exceeds the
        # span) and a per-class rate is a lower bound of what the kernel does with the SMs it holds.
        rk = {"forward": {"bound": "tensor", "ms": fwd_t, "algorithmic": flops_fwd, "achieved": achieved,
                          "unit": "TFLOP/s", "frac_of_sustained_peak": achieved / sustained,
                          "frac_of_burst_peak": achieved / burst}}
        if trace and "classes" in trace:
            alg = {"dW": (2.0 * B * T * H * V, "tensor"), "dH": (2.0 * B * T * H * V, "tensor"),
                   "grad_cached": (B * T * V * (2.0 + 2.0 + 2.0), "hbm"),  # read cache + read teacher + write G chunk
                   "grad_recompute": (2.0 * B * T * H * V, "tensor"),
                   "cast": (V * H * 4.0 + B * T * H * 4.0, "hbm")}
            for c, d in trace["classes"].items():
                a, bound = alg.get(c, (0.0, "hbm"))
                t_s = d["sum_ms"] * 1e-3
                e = {"bound": bound, "launches": d["n"], "sum_ms": d["sum_ms"], "algorithmic": a}
                if bound == "tensor":
                    e.update({"achieved": a / t_s / 1e12, "unit": "TFLOP/s", "frac_of_sustained_peak": a / t_s / 1e12 / sustained,
                              "frac_of_burst_peak": a / t_s / 1e12 / burst})
                else:
                    e.update({"achieved": a / t_s / 1e9, "unit": "GB/s", "frac_of_hbm_peak": a / t_s / 1e9 / hbm})
                rk[c] = e
            rk["backward_span_ms"] = trace["span_ms"]
            rk["note"] = ("backward kernels run concurrently on three streams (gradient chain, dW chain, dH chain): class "
                          "durations overlap and share SMs; the span and `roofline_step` are the non-overlapping figures")
        line = {
            "metric": METRIC, "value": tokens / (ms * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config(world), "clocks": clocks,
            "timing": {"how": f"{n_blocks} blocks of {steps} steps back to back ({loop_s:.2f} s of device time), CUDA events "
                              "per block, max over ranks per block; ms_per_step = median block / steps (power-settled)",
                       "blocks": n_blocks, "loop_seconds": loop_s, "first_block_ms_per_step": first_ms,
                       "min_block_ms_per_step": min(block_ms) / steps, "max_block_ms_per_step": max(block_ms) / steps,
                       "first_block_value": tokens / (first_ms * 1e-3)},
            "e2e": {"value": tokens / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 16,
                    "h2d_gb_per_s_per_gpu": h2d_bytes / (e2e_ms * 1e-3) / 1e9,
                    "note": "pinned host h, teacher logits and labels copied to the device every step by a "
                            "double-buffered prefetcher (copy of step i+1 overlaps compute of step i; PCIe bound: "
                            "1.25 GB of teacher logits per step); loss 4-tuple read back every step",
                    "losses": e2e_losses},
            "gpu_launches": int(round(launches_per_block)),
            "gpu_launches_how": "kd_launch_count() (the library counts every kernel it launches) over the timed loop, "
                                "per block of `steps` steps",
            "roofline": {
                "kernel": "kd_umma_kernel<FwdEpi> (fused lm_head GEMM + online softmax statistics + logit cache, forward)",
                "bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                "frac": achieved / sustained,
                "peak_source": f"{src} bf16_tflops_sustained: the launch is timed by CUDA events inside a {n_ph}-step "
                               f"({n_ph * phase_step_ms / 1e3:.1f} s) loop of forward+backward steps right after the "
                               f"{loop_s:.1f} s value loop (power-settled regime), not alone",
                "frac_of_burst_peak": achieved / burst, "burst_peak": burst,
                "cublas_same_shape_tflops": cublas_tf,
                "frac_of_cublas_same_shape": (achieved / cublas_tf) if cublas_tf else None,
                "traffic": traffic, "flops_per_launch": flops_fwd, "ms_per_launch": fwd_t,
                "sm_mhz_during_phase_loop": ph_clocks.get("sm_mhz_settled"),
                "note": "launch duration = CUDA events around the kd_fused_linear_fwd C call, no host sync in the loop "
                        "(tcgen05 GEMM kernel + the row-merge and reduce kernels, ~25 us)",
            },
            "roofline_step": {"bound": "tensor", "algorithmic_flops": 6.0 * B * T * H * V, "ms": ms, "achieved": step_tf,
                              "unit": "TFLOP/s", "peak": sustained, "frac": step_tf / sustained,
                              "frac_of_burst_peak": step_tf / burst,
                              "peak_source": "sustained (seconds-long loop)",
                              "first_block_frac_of_burst_peak": 6.0 * B * T * H * V / (first_ms * 1e-3) / 1e12 / burst},
            "roofline_kernels": rk,
            "step_breakdown": {
                "fwd_ms": fwd_t, "bwd_ms": bwd_t, "phase_loop_ms_per_step": phase_step_ms,
                "algorithmic_tflops_step": step_tf, "frac_of_sustained_peak_step": step_tf / sustained,
                "executed_flops_factor": "6/6 while the logit cache (constant budget, default 6 GB; this workload needs 1.23 GB) holds the chunks: "
                                         "the backward reads cached logits instead of recomputing the GEMM; chunks beyond "
                                         "the budget are recomputed (8/6)",
            },
            "losses": losses,
        }
        if multi is not None:
            line["multi_gpu"] = multi
        line["box"] = box_info(torch)
        for k in ("e2e_topk_cache", "kernels", "gpu_baselines", "read_probe"):
            if k in extra:
                line[k] = extra[k]
        if "cpu_baseline" in extra:
            cb = extra["cpu_baseline"]
            line["cpu_baseline"] = cb if "error" in cb else {
                **{k: cb.get(k) for k in ("value", "unit", "cores", "kind", "sample", "best", "os_cpu_count", "cpu_model",
                                          "peak_rss_gb", "losses")},
                "loss_only_configs0": cb.get("loss_only_configs0")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def box_info(torch):
    """Driver / GPU identity of the box that ran this line: the pool's boxes differ (driver 580.159 vs 580.178, settled
    SM clocks 1180 - 1400 MHz under the same 1 kW cap), and run-to-run comparisons need to know which one they got."""
    info = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        info["driver"] = pynvml.nvmlSystemGetDriverVersion()
        info["vbios"] = pynvml.nvmlDeviceGetVbiosVersion(h)
        info["uuid_tail"] = pynvml.nvmlDeviceGetUUID(h)[-8:]
        info["power_limit_w"] = pynvml.nvmlDeviceGetPowerManagementLimit(h) / 1000.0
    except Exception as e:  # noqa: BLE001 - context only
        info["nvml_error"] = str(e)[:80]
    return info


def cublas_same_shape(torch, h2, Wd):
    logits_buf = torch.empty(B * T, V, device=h2.device, dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(h2, Wd.t(), out=logits_buf)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(10):
        torch.matmul(h2, Wd.t(), out=logits_buf)
    c1.record()
    torch.cuda.synchronize()
    return 2.0 * B * T * H * V / (c0.elapsed_time(c1) / 10 * 1e-3) / 1e12


def e2e_topk_cache(K, torch, dev, Bc=16, k=64, steps=20):
    """The dataset-fed step (BASELINE configs[2] student side): the collator's top-k cache batch
    (collate_teacher_topk: fp16 values / int32 indices, pinned) + hidden states + labels go host -> device every step
    through HostPrefetcher, then sparse K1 fwd+bwd; losses read back.  Device-resident step beside it."""
    from speech_distill_b200.io import HostPrefetcher

    g = torch.Generator().manual_seed(5)
    feats = []
    for b in range(Bc):
        idx = torch.stack([torch.randperm(V, generator=g)[:k] for _ in range(8)]).repeat(T // 8, 1).to(torch.int32)
        val = torch.log_softmax(torch.randn(T, k, generator=g) * 3, -1).to(torch.float16)
        feats.append({"teacher_top_k_v": val, "teacher_top_k_i": idx})
    cache = K.collate_teacher_topk(feats, T, pin_memory=True)
    tv_h, ti_h = cache["teacher_top_k_v"], cache["teacher_top_k_i"]
    h_h = torch.randn(Bc, T, H, generator=g).bfloat16().pin_memory()
    l_h = torch.randint(0, V, (Bc, T), generator=g).pin_memory()
    gw = torch.Generator(device=dev).manual_seed(99)
    W = (torch.randn(V, H, device=dev, generator=gw) * (2.0 / H ** 0.5)).bfloat16().requires_grad_(True)
    batch = (h_h, tv_h, ti_h, l_h)
    out_host = torch.empty(4, dtype=torch.float32).pin_memory()
    pf = HostPrefetcher(dev)

    def run(hd, tv, ti, ll):
        W.grad = None
        hh = hd.detach().requires_grad_(True)
        o = K.fused_linear_kd_loss(hh, W, ll, teacher_top_k_v=tv, teacher_top_k_i=ti, temperature=TAU, alpha=ALPHA)
        o[0].backward()
        return o

    def e2e_step():
        cur = pf.next(batch)
        o = run(*cur)
        pf.release()
        out_host.copy_(torch.stack([x.detach().float() for x in o]), non_blocking=True)

    pf.submit(batch)
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / steps
    dev_batch = tuple(t.to(dev) for t in batch)
    for _ in range(3):
        run(*dev_batch)
    e0.record()
    for _ in range(steps):
        run(*dev_batch)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / steps
    toks = Bc * T
    return {"workload": f"sparse K1 fwd+bwd, top-k cache K={k} from collate_teacher_topk (fp16 / int32 wire dtypes), "
                        f"B={Bc} T={T} H={H} V={V} (BASELINE configs[2], student side)",
            "value": toks / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms,
            "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in batch), "d2h_bytes_per_step": 16,
            "device_resident_ms_per_step": dev_ms, "e2e_over_device": e2e_ms / dev_ms, "steps": steps,
            "losses": out_host.tolist()}


def multi_gpu_parity(K, KD, dist, torch, dev, rank, world, W, reduce_fn, count_fn, sync, Bs=1, Ts=256):
    """Token-shard semantics (distillation_loss.py:44-53,68,123: one global N, one mean): every rank scores its own
    shard (25 % of the labels ignored, so the per-rank counts differ) through the product path with the NCCL
    exchanges; rank 0 also runs the single-process fp32-accumulator entry on the concatenated batch."""
    shards = []
    for r in range(world):  # every rank builds every shard (small), deterministic per shard
        g = torch.Generator(device=dev).manual_seed(4321 + r)
        hh = torch.randn(Bs, Ts, H, device=dev, generator=g).bfloat16()
        yy = (torch.randn(Bs, Ts, V, device=dev, generator=g) * 2).bfloat16()
        ll = torch.randint(0, V, (Bs, Ts), device=dev, generator=g)
        ll[:, : (Ts // 8) * (1 + r % 4)] = -100
        shards.append((hh, yy, ll))
    hh, yy, ll = shards[rank]
    hh = hh.clone().requires_grad_(True)
    W.grad = None
    out = K.fused_linear_kd_loss(hh, W, ll, teacher_logits=yy, temperature=TAU, alpha=ALPHA, reduce_fn=reduce_fn,
                                 count_reduce_fn=count_fn, grad_sync=sync)
    out[0].backward()
    if sync is None:
        KD.allreduce_grad_(W.grad)
    torch.cuda.synchronize()
    dW_n = W.grad.float()
    W.grad = None
    l_n = [float(o.detach()) for o in out]
    res = torch.zeros(4, device=dev, dtype=torch.float64)
    # single process, concatenated batch, fp32 accumulators (no bf16 rounding of the gradients)
    hc = torch.cat([s[0] for s in shards], 0)
    yc = torch.cat([s[1] for s in shards], 0)
    lc = torch.cat([s[2] for s in shards], 0)
    l1, dH1, dW1 = K.fused_linear_kd_value_and_grad(hc, W.detach(), lc, teacher_logits=yc, temperature=TAU, alpha=ALPHA)
    dH1 = dH1.reshape(world, Bs, Ts, H)[rank]
    res[0] = max(abs(a - float(b)) / max(abs(float(b)), 1e-30) for a, b in zip(l_n, l1))
    res[1] = float((dW_n - dW1.float()).abs().max() / dW1.float().abs().max())
    res[2] = float((hh.grad.float() - dH1.float()).abs().max() / dH1.float().abs().max())
    res[3] = float(torch.nn.functional.cosine_similarity(dW_n.flatten().double(), dW1.flatten().double(), dim=0))
    mx = res.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    mn = res.clone()
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    del shards, hc, yc, lc, dW1, dW_n
    torch.cuda.empty_cache()
    return {"what": f"{world}-rank token-shard step (B={Bs} T={Ts} per rank, unequal valid-row counts, NCCL count / sums / "
                    "dW all-reduce, bf16 gradients) vs one process on the concatenated batch (fp32-accumulator entry); "
                    "max over ranks of max|a-b| / max|b|",
            "loss_err": float(mx[0]), "dW_err": float(mx[1]), "dH_err": float(mx[2]), "dW_cosine_min": float(mn[3]),
            "ok": bool(mx[0] < 1e-5 and mx[1] < 8e-3 and mx[2] < 8e-3)}


def allreduce_precision(dist, torch, step, h, y, labels, W, clear):
    """What the bf16 SUM all-reduce of dW costs at full size: local bf16 dW of this step reduced in bf16 (what the
    product does) against the same values reduced in fp32."""
    clear()
    # local gradient without the data-parallel all-reduce of dW (count / sums still global)
    import speech_distill_b200 as K
    from speech_distill_b200 import dist as KD

    rf, cf = KD.make_reduce_fns()
    o = K.fused_linear_kd_loss(h, W, labels, teacher_logits=y, temperature=TAU, alpha=ALPHA, reduce_fn=rf,
                               count_reduce_fn=cf)
    o[0].backward()
    local = W.grad
    a = local.float()
    dist.all_reduce(a, op=dist.ReduceOp.SUM)
    b = local.clone()
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    scale = float(a.abs().max())
    err = float((b.float() - a).abs().max()) / scale
    ref_round = float((a.bfloat16().float() - a).abs().max()) / scale
    clear()
    del a, b
    torch.cuda.empty_cache()
    return {"max_err_over_max": err, "one_bf16_rounding_of_the_fp32_sum": ref_round,
            "what": "max|allreduce_bf16(dW) - allreduce_fp32(dW)| / max|dW| at the full configs[1] size on this rank"}


def vocab_parallel_block(K, dist, torch, dev, rank, world, steps):
    """BASELINE configs[4] second half: every rank sees the same B*world sequences and holds V/world rows of the LM
    head + the matching teacher columns; parity against the unsharded kernels (small batch), then the step time."""
    slices = K.vocab_slices(V, world)
    v0, v1 = slices[rank]
    gw = torch.Generator(device=dev).manual_seed(99)
    Wf = (torch.randn(V, H, device=dev, generator=gw) * (2.0 / H ** 0.5)).bfloat16()

    def tokens(Bn, seed):
        g = torch.Generator(device=dev).manual_seed(seed)  # same seed on every rank: identical tokens
        return (torch.randn(Bn, T, H, device=dev, generator=g).bfloat16(),
                torch.randint(0, V, (Bn, T), device=dev, generator=g))

    def teacher_cols(Bn, c0, c1, seed):
        yy = torch.empty(Bn, T, c1 - c0, device=dev, dtype=torch.bfloat16)
        for b in range(Bn):
            g = torch.Generator(device=dev).manual_seed(seed * 1000 + b)
            yy[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()[:, c0:c1]
        return yy

    hp, lp = tokens(2, 7)
    y_full = teacher_cols(2, 0, V, 5)
    hs, Ws = hp.clone().requires_grad_(True), Wf[v0:v1].clone().requires_grad_(True)
    out = K.fused_linear_kd_loss_vocab_parallel(hs, Ws, lp, v0, teacher_logits_slice=y_full[..., v0:v1].contiguous())
    out[0].backward()
    hu, Wu = hp.clone().requires_grad_(True), Wf.clone().requires_grad_(True)
    ref = K.fused_linear_kd_loss(hu, Wu, lp, teacher_logits=y_full)
    ref[0].backward()
    torch.cuda.synchronize()
    errs = torch.tensor([
        max(abs(float(a.detach()) - float(b.detach())) / max(1.0, abs(float(b.detach()))) for a, b in zip(out, ref)),
        float((hs.grad.float() - hu.grad.float()).abs().max() / hu.grad.float().abs().max()),
        float((Ws.grad.float() - Wu.grad[v0:v1].float()).abs().max() / Wu.grad.float().abs().max())], device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    del y_full, hu, Wu, hs, Ws, ref, out
    torch.cuda.empty_cache()
    Bg = B * world
    hg, lg = tokens(Bg, 11)
    hg.requires_grad_(True)
    Wl = Wf[v0:v1].clone().requires_grad_(True)
    del Wf
    yl = teacher_cols(Bg, v0, v1, 6)

    def vstep():
        hg.grad = None
        Wl.grad = None
        o = K.fused_linear_kd_loss_vocab_parallel(hg, Wl, lg, v0, teacher_logits_slice=yl)
        o[0].backward()
        return o

    for _ in range(3):
        o = vstep()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    n = max(10, min(steps, 40))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        o = vstep()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    return {"mode": "vocab-parallel (W and teacher columns sharded over V, LSE max/sum records all-gathered, dH all-reduced)",
            "tokens_per_step": Bg * T, "ms_per_step": ms, "tokens_per_s": Bg * T / (ms * 1e-3), "steps": n,
            "slice_rows": v1 - v0, "parity_vs_unsharded": {"loss_err": float(errs[0]), "dH_err": float(errs[1]),
                                                           "dW_err": float(errs[2]),
                                                           "ok": bool(errs[0] < 1e-5 and errs[1] < 8e-3 and errs[2] < 8e-3)},
            "losses": [float(x.detach()) for x in o]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernels", action="store_true", help="skip the `kernels` and `gpu_baselines` blocks")
    ap.add_argument("--cpu-baseline-leg", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
