"""One K3 launch at V = 152,936, R = 2048 bf16 rows (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
g = torch.Generator(device="cuda").manual_seed(1)
x = (torch.randn(2048, 152936, device="cuda", generator=g) * 2).bfloat16()
for _ in range(2):
    v, i = K.teacher_topk_logprobs(x, 64)
torch.cuda.synchronize()
print(v.shape, int(i[0, 0]))
