"""Bring-up diagnostics for the tcgen05 GEMM building block (run on the GPU box, prints numbers
instead of asserting).  Each case runs in a subprocess with a timeout so that a hung kernel cannot
take the whole call down."""
import subprocess
import sys

CASES = [
    # M, N, K, a_mn, b_mn
    (128, 256, 64, 0, 0), (128, 256, 128, 0, 0), (128, 256, 1024, 0, 0), (256, 512, 256, 0, 0), (300, 700, 136, 0, 0),
    (128, 256, 64, 0, 1), (256, 512, 256, 0, 1), (128, 256, 64, 1, 1), (256, 512, 256, 1, 1), (4096, 2048, 1024, 0, 0),
]

CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from speech_distill_b200 import _lib
M, N, K, a_mn, b_mn = map(int, sys.argv[1:6])
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
ref = A.float() @ B.float().t()
Ain = A.t().contiguous() if a_mn else A
Bin = B.t().contiguous() if b_mn else B
C = torch.full((M, N), float("nan"), device="cuda")
rc = lib.kd_gemm_bf16(Ain.data_ptr(), Ain.stride(0), a_mn, Bin.data_ptr(), Bin.stride(0), b_mn, C.data_ptr(), C.stride(0), M, N, K, torch.cuda.current_stream().cuda_stream)
print("rc", rc, lib.kd_last_error())
torch.cuda.synchronize()
err = (C - ref).abs()
print("nan", int(torch.isnan(C).sum()), "max_err", float(err.nan_to_num(1e9).max()), "ref_max", float(ref.abs().max()))
bad = (err.nan_to_num(1e9) > 1e-2)
print("bad_frac", float(bad.float().mean()))
if bad.any():
    rows = bad.any(1).nonzero().flatten()[:8].tolist(); cols = bad.any(0).nonzero().flatten()[:8].tolist()
    print("bad rows", rows, "bad cols", cols)
    print("C[0,:8]", C[0,:8].tolist()); print("R[0,:8]", ref[0,:8].tolist())
'''

for c in CASES:
    print("=== gemm", c, flush=True)
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, *map(str, c)], capture_output=True, text=True, timeout=90)
        print(r.stdout[-1500:], r.stderr[-1500:], flush=True)
    except subprocess.TimeoutExpired:
        print("TIMEOUT (hang)", flush=True)
