"""Two seconds of back-to-back K1 steps (and of cuBLAS bf16 GEMMs for comparison) with SM clock, power and throttle
reasons sampled by NVML every 10 ms: is the step power-capped?"""
import os, sys, threading, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
import pynvml
pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
B, T, H, V = 8, 512, 1024, 152936
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
h = torch.randn(B, T, H, device=dev, generator=g).bfloat16().requires_grad_(True)
W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16().requires_grad_(True)
y = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
for b in range(B):
    y[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
labels = torch.randint(0, V, (B, T), device=dev, generator=g)
A8 = torch.randn(8192, 8192, device=dev).bfloat16()
B8 = torch.randn(8192, 8192, device=dev).bfloat16()


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop_ = [], False

    def run(self):
        while not self.stop_:
            self.rows.append((pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM),
                              pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0,
                              pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd)))
            time.sleep(0.01)


def run(name, fn, flops, seconds=2.0):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s = Sampler(); s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    s.stop_ = True; s.join()
    t = e0.elapsed_time(e1) / n * 1e-3
    rows = s.rows[len(s.rows) // 2:]  # second half: settled
    clk = sorted(r[0] for r in rows)[len(rows) // 2]
    pw = sorted(r[1] for r in rows)[len(rows) // 2]
    reasons = 0
    for r in rows:
        reasons |= r[2]
    print(f"{name:28s} {t*1e3:7.3f} ms/iter  {flops/t/1e12:7.0f} TF/s   settled: SM clock {clk} MHz, power {pw:.0f} W, "
          f"reasons mask 0x{reasons:x} (0x4 = sw_power_cap), {n} iters")


def step():
    h.grad = None
    W.grad = None
    out = K.fused_linear_kd_loss(h, W, labels, teacher_logits=y)
    out[0].backward()


from speech_distill_b200 import loss as KL  # noqa: E402

h2, y2, Wd = h.detach().reshape(B * T, H), y.reshape(B * T, V), W.detach()
row_target, n_valid = KL.prepare_rows(labels, None, B, T, -100, h.device)
coef = torch.tensor([0.5, 0.5], dtype=torch.float32, device=h.device)
cache = KL.alloc_logit_cache(B * T, V, 0, h.device)
sums, row_stats, ws = KL._fused_forward(h2, Wd, y2, row_target, 2.0, 0.5, 0, cache=cache)
logits = torch.empty(B * T, V, device=h.device, dtype=torch.bfloat16)
G16 = torch.randn(B * T, 18944, device=h.device).half()
W16 = Wd[:18944].half()
h16 = h2.half()
dWc = torch.empty(18944, H, device=h.device, dtype=torch.half)
dHc = torch.empty(B * T, H, device=h.device, dtype=torch.half)


def fwd_only():
    KL._fused_forward(h2, Wd, y2, row_target, 2.0, 0.5, 0, cache=cache)


def bwd_only():
    KL._fused_backward(h2, Wd, y2, row_target, row_stats, n_valid, coef, 2.0, 1, 0, 0, torch.bfloat16, True, True, ws,
                       cache=cache)


def cublas_bwd_chunk():  # the library's version of one chunk's dW and dH GEMMs (fp16 operands, same shapes)
    torch.matmul(G16.t(), h16, out=dWc)
    torch.matmul(G16, W16, out=dHc)


F = 2.0 * B * T * H * V
run("K1 step (6 R H V algorithmic)", step, 3 * F)
run("K1 forward only", fwd_only, F)
run("K1 backward only", bwd_only, 2 * F)
run("cuBLAS bf16 8192^3", lambda: torch.matmul(A8, B8), 2.0 * 8192 ** 3)
run("cuBLAS lm_head shape", lambda: torch.matmul(h2, Wd.t(), out=logits), F)
run("cuBLAS dW+dH of one chunk", cublas_bwd_chunk, 2 * 2.0 * B * T * H * 18944)
run("K1 step again", step, 3 * F)
