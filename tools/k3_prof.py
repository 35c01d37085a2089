"""A few K3 launches at V = 152,936, R = 8192 bf16 rows (configs[2] shape) for ncu captures:
KD_TOPK_BLOCKS=1 makes one sweep launch cover all rows."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.empty(R, 152936, device="cuda", dtype=torch.bfloat16)
for r0 in range(0, R, 1024):
    x[r0:r0 + 1024] = (torch.randn(min(1024, R - r0), 152936, device="cuda", generator=g) * 2).bfloat16()
for _ in range(2):
    v, i = K.teacher_topk_logprobs(x, 64)
torch.cuda.synchronize()
print(v.shape, int(i[0, 0]))
