"""teacher_head_topk at the configs[2] shape as a function of the row block (scratch = 2 x row_block x V bf16),
fused (GEMM epilogue statistics + piece-wise selection) against the unfused pipeline, the head GEMM alone beside them.
Round-robin over the variants so that clock drift hits all of them alike."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
V, Ht, R = 152936, 2048, 8192
g = torch.Generator(device="cuda").manual_seed(0)
ht = torch.randn(R, Ht, device="cuda", generator=g).bfloat16()
Wt = (torch.randn(V, Ht, device="cuda", generator=g) * (2.5 / Ht ** 0.5)).bfloat16()
full = torch.empty(R, V, device="cuda", dtype=torch.bfloat16)


def timeit(fn, n=6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


variants = {"head GEMM alone (one launch, full [R,V] buffer)": lambda: K.linear_bf16(ht, Wt, full)}
for rb in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]:
    variants[f"fused   row_block {rb:5d}"] = (lambda rb=rb: K.teacher_head_topk(ht, Wt, 64, row_block=rb, fused=True))
    variants[f"unfused row_block {rb:5d}"] = (lambda rb=rb: K.teacher_head_topk(ht, Wt, 64, row_block=rb, fused=False))
for fn in variants.values():
    fn()
    fn()
best = {k: [] for k in variants}
for rnd in range(3):
    for k, fn in variants.items():
        best[k].append(timeit(fn))
for k, ts in best.items():
    t = sorted(ts)[1]
    print(f"{k}: {t*1e3:6.2f} ms = {2.0*R*Ht*V/t/1e12:5.0f} TFLOP/s   (runs: {', '.join(f'{x*1e3:.2f}' for x in ts)})")
