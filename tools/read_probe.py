"""Read-only streaming bandwidth of this GPU by load mechanism (kd_probe_read_bandwidth) next to torch's copy figure:
the roofline a read-only stream (K2 forward, K3) is bounded by.  Prints one JSON line."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_distill_b200 import _lib
from speech_distill_b200._lib import check, stream_ptr


def probe(nbytes=2 << 30, dev="cuda"):
    lib = _lib.load()
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    buf.random_(0, 255)
    scratch = torch.zeros(4, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}

    def t(fn, n=8):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e-3

    for mode, name in ((0, "ldg"), (1, "ldg_evict_first"), (2, "tma_bulk")):
        for ctas, unroll in ((4, 4), (8, 4), (4, 8), (8, 8), (4, 16)) if mode < 2 else ((1, 12), (2, 6), (3, 4), (4, 3)):
            sec = t(lambda: check(lib.kd_probe_read_bandwidth(buf.data_ptr(), nbytes, mode, ctas, unroll, scratch.data_ptr(),
                                                              stream_ptr(torch.device(dev))), "probe"))
            out[f"{name}_ctas{ctas}_u{unroll}"] = round(nbytes / sec / 1e9)
    half = buf[: nbytes // 2]
    dst = torch.empty_like(half)
    sec = t(lambda: dst.copy_(half))
    out["torch_copy_read_plus_write"] = round(2 * half.numel() / sec / 1e9)
    out["best_read"] = max(v for k, v in out.items() if k != "torch_copy_read_plus_write")
    return out


if __name__ == "__main__":
    print(json.dumps(probe()))
