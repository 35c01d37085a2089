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


run("K1 step (8 R H V executed)", step, 8.0 * B * T * H * V)
run("cuBLAS bf16 8192^3", lambda: torch.matmul(A8, B8), 2.0 * 8192 ** 3)
run("K1 step again", step, 8.0 * B * T * H * V)
