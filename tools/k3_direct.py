"""Same-process comparison of the K3 entry points at the configs[2] shape (diagnosis)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
from speech_distill_b200 import _lib
from speech_distill_b200._lib import check, dtype_code, stream_ptr

V, R, k = 152936, int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 64
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(7)
x = torch.empty(R, V, device=dev, dtype=torch.bfloat16)
for r0 in range(0, R, 1024):
    x[r0:r0 + 1024] = (torch.randn(min(1024, R - r0), V, device=dev, generator=g) * 2).bfloat16()
lib = _lib.load()
out_v = torch.empty((R, k), dtype=torch.float16, device=dev)
out_i = torch.empty((R, k), dtype=torch.int32, device=dev)
ws_bytes = int(lib.kd_topk_workspace_bytes(R, V))
ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
ws_ptr = (ws.data_ptr() + 255) & ~255


def direct():
    check(lib.kd_topk_logprobs(x.data_ptr(), dtype_code(x.dtype), R, V, x.stride(0), k, out_v.data_ptr(), out_i.data_ptr(),
                               stream_ptr(dev)), "direct")


def with_ws():
    check(lib.kd_topk_logprobs_ws(x.data_ptr(), dtype_code(x.dtype), R, V, x.stride(0), k, out_v.data_ptr(),
                                  out_i.data_ptr(), ws_ptr, ws_bytes, stream_ptr(dev)), "ws")


def api():
    K.teacher_topk_logprobs(x, k)


import pynvml
pynvml.nvmlInit()
nv = pynvml.nvmlDeviceGetHandleByIndex(0)
clk = {}


def t(fn, n=40, name=""):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        # sampled while the queued launches run
        clk[name] = (pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_MEM),
                     round(pynvml.nvmlDeviceGetPowerUsage(nv) / 1000))
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return round(best)


res = {"env": {a: b for a, b in os.environ.items() if a.startswith("KD_TOPK")}}
for name, fn in (("direct", direct), ("with_ws", with_ws), ("api", api), ("direct_again", direct)):
    res[name + "_us"] = t(fn, name=name)
res["sm_mem_mhz_power_w"] = clk
res["gpu"] = {"uuid": pynvml.nvmlDeviceGetUUID(nv)[-8:], "vbios": pynvml.nvmlDeviceGetVbiosVersion(nv), "driver": pynvml.nvmlSystemGetDriverVersion(),
              "temp": pynvml.nvmlDeviceGetTemperature(nv, 0)}
print(json.dumps(res))
