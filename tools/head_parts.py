"""Parts of the fused teacher-head top-k at the configs[2] shape: stats GEMM per row block alone, selection alone."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
from speech_distill_b200 import _lib, topk as KT
from speech_distill_b200._lib import check, stream_ptr
V, Ht, R, k = 152936, 2048, 8192, 64
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
ht = torch.randn(R, Ht, device="cuda", generator=g).bfloat16()
Wt = (torch.randn(V, Ht, device="cuda", generator=g) * (2.5 / Ht ** 0.5)).bfloat16()
lib = _lib.load()
ps, qs = KT.head_topk_layout(V)


def timeit(fn, n=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rb in (1024, 2048, 8192):
    ld = -(-V // 8) * 8
    scratch = torch.empty((rb, ld), dtype=torch.bfloat16, device=dev)
    pmax = torch.empty((rb, ps), dtype=torch.bfloat16, device=dev)
    part = torch.empty((rb, qs, 2), dtype=torch.float32, device=dev)
    out_v = torch.empty((R, k), dtype=torch.float16, device=dev)
    out_i = torch.empty((R, k), dtype=torch.int32, device=dev)
    n_part = ctypes.c_int(0)

    def gemm_stats():
        for r0 in range(0, R, rb):
            hb = ht[r0:r0 + rb]
            check(lib.kd_head_logits_stats(hb.data_ptr(), hb.stride(0), Wt.data_ptr(), Wt.stride(0), scratch.data_ptr(),
                                           ld, pmax.data_ptr(), ps, part.data_ptr(), qs, ctypes.byref(n_part), rb, Ht, V,
                                           stream_ptr(dev)), "stats")

    def gemm_plain():
        for r0 in range(0, R, rb):
            K.linear_bf16(ht[r0:r0 + rb], Wt, scratch[:, :V])

    def select():
        for r0 in range(0, R, rb):
            check(lib.kd_head_topk_select(scratch.data_ptr(), ld, pmax.data_ptr(), ps, part.data_ptr(), qs, n_part.value,
                                          rb, V, k, out_v[r0:r0 + rb].data_ptr(), out_i[r0:r0 + rb].data_ptr(),
                                          stream_ptr(dev)), "select")

    def k3():
        for r0 in range(0, R, rb):
            K.teacher_topk_logprobs(scratch[:, :V], k)

    t1, t2 = timeit(gemm_stats), timeit(gemm_plain)
    t3, t4 = timeit(select), timeit(k3)
    print(f"row_block {rb}: stats GEMM x{R // rb} {t1:.3f} ms (n_part {n_part.value}), plain GEMM {t2:.3f} ms, "
          f"selection {t3:.3f} ms, full-row compaction {t4:.3f} ms")
for fused in (True, False):
    t = timeit(lambda: K.teacher_head_topk(ht, Wt, k, fused=fused))
    print(f"teacher_head_topk fused={fused}: {t:.3f} ms")
