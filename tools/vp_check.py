"""Vocab-parallel mode on real GPUs (torchrun, NCCL): parity against the unsharded kernels on rank 0's view, then
step time at BASELINE configs[4] scale (every rank sees the same B_global x 512 tokens and holds V / world rows of
the LM head plus the matching teacher columns).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/vp_check.py [B_per_gpu]
"""
import json, os, sys, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_distill_b200 as K

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H, V, T = 1024, 152936, 512
slices = K.vocab_slices(V, world)
v0, v1 = slices[rank]


def make(B, seed):
    g = torch.Generator(device=dev).manual_seed(seed)  # same seed on every rank: identical tokens
    h = torch.randn(B, T, H, device=dev, generator=g).bfloat16()
    labels = torch.randint(0, V, (B, T), device=dev, generator=g)
    gw = torch.Generator(device=dev).manual_seed(99)
    W = (torch.randn(V, H, device=dev, generator=gw) * (2.0 / H ** 0.5)).bfloat16()
    return h, labels, W


def teacher_cols(B, c0, c1, seed):
    # column block [c0, c1) of a teacher tensor defined per (row-block, column) so that every rank can build its slice
    y = torch.empty(B, T, c1 - c0, device=dev, dtype=torch.bfloat16)
    for b in range(B):
        g = torch.Generator(device=dev).manual_seed(seed * 1000 + b)
        full = (torch.randn(T, V, device=dev, generator=g) * 2).bfloat16()
        y[b] = full[:, c0:c1]
    return y


# ---- parity (small batch): vocab-parallel over NCCL vs the unsharded kernels ----
B = 2
h, labels, W = make(B, 7)
y_full = teacher_cols(B, 0, V, 5)
hs = h.clone().requires_grad_(True)
Ws = W[v0:v1].clone().requires_grad_(True)
out = K.fused_linear_kd_loss_vocab_parallel(hs, Ws, labels, v0, teacher_logits_slice=y_full[..., v0:v1].contiguous())
out[0].backward()
hu = h.clone().requires_grad_(True)
Wu = W.clone().requires_grad_(True)
ref = K.fused_linear_kd_loss(hu, Wu, labels, teacher_logits=y_full)
ref[0].backward()
torch.cuda.synchronize()
l_err = max(abs(float(a) - float(b)) / max(1.0, abs(float(b))) for a, b in zip(out, ref))
dh_err = float((hs.grad.float() - hu.grad.float()).abs().max() / hu.grad.float().abs().max())
dw_err = float((Ws.grad.float() - Wu.grad[v0:v1].float()).abs().max() / Wu.grad.float().abs().max())
errs = torch.tensor([l_err, dh_err, dw_err], device=dev)
dist.all_reduce(errs, op=dist.ReduceOp.MAX)
ok = bool(errs[0] < 1e-5 and errs[1] < 8e-3 and errs[2] < 8e-3)
del y_full, hu, Wu, hs, Ws, ref, out
torch.cuda.empty_cache()

# ---- timing: global batch = B_per_gpu * world sequences on every rank ----
Bg = (int(sys.argv[1]) if len(sys.argv) > 1 else 8) * world
h, labels, _ = make(Bg, 11)
h.requires_grad_(True)
Wl = W[v0:v1].clone().requires_grad_(True)
yl = teacher_cols(Bg, v0, v1, 6)


def step():
    h.grad = None
    Wl.grad = None
    o = K.fused_linear_kd_loss_vocab_parallel(h, Wl, labels, v0, teacher_logits_slice=yl)
    o[0].backward()
    return o


for _ in range(3):
    o = step()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    o = step()
e1.record()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"mode": "vocab-parallel", "n_gpus": world, "parity_ok": ok, "loss_err": float(errs[0]),
                      "dH_err_vs_unsharded": float(errs[1]), "dW_err_vs_unsharded": float(errs[2]),
                      "tokens_per_step": Bg * T, "ms_per_step": float(ms[0]),
                      "tokens_per_s": Bg * T / (float(ms[0]) * 1e-3), "slice_rows": v1 - v0,
                      "losses": [float(x) for x in o]}), flush=True)
dist.destroy_process_group()
