"""Probe (torchrun, N >= 2): does torch's symmetric memory rendezvous work on this box, is NVLS multicast available, and
what do the multimem / two-shot all-reduces cost for the 313 MB bf16 LM-head gradient next to NCCL's all_reduce?"""
import json, os, sys, time
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
V, H = 152936, 1024
out = {"world": world}
try:
    import torch.distributed._symmetric_memory as symm_mem

    gname = dist.group.WORLD.group_name
    t = symm_mem.empty((V, H), dtype=torch.bfloat16, device=dev)
    hdl = symm_mem.rendezvous(t, gname)
    out["multicast_ptr_nonzero"] = bool(getattr(hdl, "multicast_ptr", 0))
    out["handle_attrs"] = [a for a in dir(hdl) if not a.startswith("_")][:40]
    x = torch.empty((V, H), dtype=torch.bfloat16, device=dev)

    def fill():
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        x.copy_(torch.randn(V, H, device=dev, generator=g).bfloat16() * 0.01)
        t.copy_(x)

    def timeit(fn, n=10):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    fill()
    ref = x.float()
    dist.all_reduce(ref)  # fp32 reference sum
    for name, fn in (("nccl_all_reduce", lambda: dist.all_reduce(x)),
                     ("multimem_all_reduce_", lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname)),
                     ("two_shot_all_reduce_", lambda: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", gname))):
        try:
            fill()
            fn()
            torch.cuda.synchronize()
            got = (t if "shot" in name or "multimem" in name else x).float()
            err = float((got - ref).abs().max() / ref.abs().max())
            out[name] = {"ms": timeit(fn), "err_vs_fp32_sum": err}
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
except Exception as e:  # noqa: BLE001
    out["error"] = f"{type(e).__name__}: {e}"[:500]
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
