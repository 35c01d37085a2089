"""One GPU: what does splitting the backward into vocabulary ranges cost by itself (no communication)?  A stand-in
GradSync hands every range to a side stream that does nothing; with and without the dw_ready_stream hand-over."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K
from speech_distill_b200.dist import plan_ranges

B, T, H, V = 8, 512, 1024, 152936
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
h = torch.randn(B, T, H, device=dev, generator=g).bfloat16().requires_grad_(True)
W = (torch.randn(V, H, device=dev, generator=g) * (2.0 / H ** 0.5)).bfloat16().requires_grad_(True)
y = torch.empty(B, T, V, device=dev, dtype=torch.bfloat16)
for b in range(B):
    y[b] = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
labels = torch.randint(0, V, (B, T), device=dev, generator=g)


class Stub:
    def __init__(self, n, ready):
        self.n, self.side, self.ready = n, torch.cuda.Stream(), ready
        if ready:
            self.ready_stream_ptr = lambda device: self.side.cuda_stream

    def ranges(self, V, row_begin, v_chunk):
        return plan_ranges(V, row_begin, v_chunk, self.n)

    def sm_limit(self):
        return 0

    def reduce_rows(self, grad, r0, r1, last=True):
        if last or not self.ready:
            self.side.wait_stream(torch.cuda.current_stream())

    def finish(self):
        torch.cuda.current_stream().wait_stream(self.side)


def step(sync):
    h.grad = W.grad = None
    out = K.fused_linear_kd_loss(h, W, labels, teacher_logits=y, grad_sync=sync)
    out[0].backward()


for name, sync in (("one call", None), ("6 ranges, join per range", Stub(6, False)), ("6 ranges, ready stream", Stub(6, True)),
                   ("9 ranges, ready stream", Stub(9, True)), ("one call again", None)):
    for _ in range(10):
        step(sync)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        step(sync)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / 100:.3f} ms/step")
