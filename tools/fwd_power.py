"""Sustained forward-kernel rate (dense teacher / no teacher / epilogue math skipped) and cuBLAS bf16 in the same
process, 60 launches back to back each, with SM clock and power sampled by NVML during every loop."""
import os, sys, threading, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
from speech_distill_b200 import loss as KL
import pynvml
pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
B, T, H, V = 8, 512, 1024, 152936
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
h = torch.randn(B * T, H, device=dev, generator=g).bfloat16()
W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16()
y = torch.empty(B * T, V, device=dev, dtype=torch.bfloat16)
for b in range(B):
    y[b * T:(b + 1) * T] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
labels = torch.randint(0, V, (B, T), device=dev, generator=g)
row_target, n_valid = KL.prepare_rows(labels, None, B, T, -100, dev)
A8 = torch.randn(8192, 8192, device=dev).bfloat16()
B8 = torch.randn(8192, 8192, device=dev).bfloat16()


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.clk, self.pw, self.stop_ = [], [], False

    def run(self):
        while not self.stop_:
            self.clk.append(pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM))
            self.pw.append(pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0)
            time.sleep(0.005)


def run(name, fn, flops, n=60):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s = Sampler(); s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    s.stop_ = True; s.join()
    t = e0.elapsed_time(e1) / n * 1e-3
    clk = sorted(s.clk)[len(s.clk) // 2] if s.clk else 0
    pw = sorted(s.pw)[len(s.pw) // 2] if s.pw else 0
    print(f"{name:38s} {t*1e6:7.0f} us  {flops/t/1e12:7.0f} TF/s   median SM clock {clk} MHz, power {pw:.0f} W ({len(s.clk)} samples)")


fl = 2.0 * B * T * H * V
mode = os.environ.get("KD_DEBUG_SKIP_MATH", "0")
run(f"fwd dense teacher (skip_math={mode})", lambda: KL._fused_forward(h, W, y, row_target, 2.0, 0.5, 0), fl)
run(f"fwd no teacher    (skip_math={mode})", lambda: KL._fused_forward(h, W, None, row_target, 1.0, 1.0, 0), fl)
run("cuBLAS bf16 8192^3", lambda: torch.matmul(A8, B8), 2.0 * 8192 ** 3, n=100)
logits = torch.empty(B * T, V, device=dev, dtype=torch.bfloat16)
run("cuBLAS bf16 lm_head shape (logits out)", lambda: torch.matmul(h, W.t(), out=logits), fl)
run("kd_linear_bf16 lm_head shape", lambda: K.linear_bf16(h, W, logits), fl)
C = torch.empty(8192, 8192, device=dev, dtype=torch.float32)
lib = K.load_library()
st = torch.cuda.current_stream().cuda_stream
run("kd_gemm_bf16 8192^3 (fp32 out)", lambda: lib.kd_gemm_bf16(A8.data_ptr(), 8192, 0, B8.data_ptr(), 8192, 0, C.data_ptr(), 8192, 8192, 8192, 8192, st), 2.0 * 8192 ** 3, n=100)
