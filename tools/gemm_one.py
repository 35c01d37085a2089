import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_distill_b200 import _lib
lib = _lib.load()
M = N = K = 4096
A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
C = torch.empty(M, N, device="cuda")
for _ in range(3):
    lib.kd_gemm_bf16(A.data_ptr(), A.stride(0), 0, B.data_ptr(), B.stride(0), 0, C.data_ptr(), C.stride(0), M, N, K, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok")
