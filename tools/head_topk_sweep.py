"""teacher_head_topk at the configs[2] shape as a function of the row block (scratch = 2 x row_block x V bf16)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
V, Ht, R = 152936, 2048, 8192
g = torch.Generator(device="cuda").manual_seed(0)
ht = torch.randn(R, Ht, device="cuda", generator=g).bfloat16()
Wt = (torch.randn(V, Ht, device="cuda", generator=g) * (2.5 / Ht ** 0.5)).bfloat16()
for rb in [int(a) for a in sys.argv[1:]] or [1024, 1480, 2048, 2960, 4096, 8192]:
    for _ in range(2):
        K.teacher_head_topk(ht, Wt, 64, row_block=rb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(6):
        K.teacher_head_topk(ht, Wt, 64, row_block=rb)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 6 * 1e-3
    print(f"row_block {rb:5d}: {t*1e3:6.2f} ms = {2.0*R*Ht*V/t/1e12:5.0f} TFLOP/s (scratch 2 x {rb*V*2/1e6:.0f} MB)")
