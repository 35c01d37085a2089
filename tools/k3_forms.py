"""K3 forms across row counts (one process per form: the library reads KD_TOPK_FORM once).  Prints `R form us`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
V, k = 152936, 64
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(7)
Rmax = 8192
x = torch.empty(Rmax, V, device=dev, dtype=torch.bfloat16)
for r0 in range(0, Rmax, 1024):
    x[r0:r0 + 1024] = (torch.randn(1024, V, device=dev, generator=g) * 2).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = []
for R in [int(a) for a in sys.argv[1:]] or [256, 1024, 2048, 4096, 8192]:
    xs = x[:R]
    for _ in range(3):
        K.teacher_topk_logprobs(xs, k)
    ts = []
    for _ in range(12):
        flush.zero_()          # small launches would otherwise re-read their rows from L2
        e0.record()
        K.teacher_topk_logprobs(xs, k)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    out.append(f"R={R}: {ts[len(ts) // 2]:.0f}")
print(os.environ.get("KD_TOPK_FORM", "auto"), " ".join(out))
