"""Mainloop-only throughput of the tcgen05 GEMM building block (plain fp32-store epilogue)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_distill_b200 import _lib
lib = _lib.load()
def run(M, N, K, a_mn, b_mn, iters=20):
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    Ain = A.t().contiguous() if a_mn else A; Bin = B.t().contiguous() if b_mn else B
    C = torch.empty(M, N, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    f = lambda: lib.kd_gemm_bf16(Ain.data_ptr(), Ain.stride(0), a_mn, Bin.data_ptr(), Bin.stride(0), b_mn, C.data_ptr(), C.stride(0), M, N, K, s)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): torch.matmul(A, B.t())
    torch.cuda.synchronize(); t0.record()
    for _ in range(iters): torch.matmul(A, B.t())
    t1.record(); torch.cuda.synchronize()
    ms2 = t0.elapsed_time(t1) / iters
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: ours {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF/s | cuBLAS(bf16 out) {ms2*1e3:8.1f} us {fl/ms2/1e9:7.1f} TF/s", flush=True)
print("KD_UMMA_CTA_GROUP =", os.environ.get("KD_UMMA_CTA_GROUP", "2 (default)"))
run(4096, 9472, 1024, 0, 0)
run(4096, 37888, 1024, 0, 0)
run(9472, 1024, 4096, 1, 1)
run(4096, 1024, 9472, 0, 1)
run(8192, 8192, 8192, 0, 0, iters=5)
