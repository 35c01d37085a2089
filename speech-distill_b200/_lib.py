"""ctypes binding of libkd_b200.so (the C ABI declared in include/kd_b200.h).

There is no CPU fallback: if the library is missing, or a call is made without a CUDA device,
the caller gets an exception, never a silently slower path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# KD_B200_LIB selects another build of the same library (A/B experiments); default: the in-tree one
LIB_PATH = os.environ.get("KD_B200_LIB") or os.path.join(_HERE, "libkd_b200.so")

KD_DTYPE_F32, KD_DTYPE_BF16, KD_DTYPE_F16 = 0, 1, 2
KD_TEACHER_NONE, KD_TEACHER_DENSE, KD_TEACHER_SPARSE = 0, 1, 2
ABI_VERSION = 5  # KD_ABI_VERSION in include/kd_b200.h
KD_RANGE_FIRST, KD_RANGE_LAST = 1, 2
KD_GRAD_DH_F32 = 0x100

_c = ctypes
_vp, _i32, _i64, _f32, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); mirrors include/kd_b200.h one to one
SIGNATURES = {
    "kd_version": (_i32, []),
    "kd_last_error": (_c.c_char_p, []),
    "kd_launch_count": (_c.c_ulonglong, []),
    "kd_fused_bwd_trace_begin": (_i32, []),
    "kd_fused_bwd_trace_read": (_i32, [_vp, _i32]),
    "kd_device_info": (_i32, [_c.POINTER(_i32)] * 3),
    "kd_prepare_rows": (_i32, [_vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp]),
    "kd_finalize_losses": (_i32, [_vp, _f32, _f32, _i32, _vp, _vp]),
    "kd_stream_workspace_bytes": (_sz, []),
    "kd_dense_fwd_bwd": (_i32, [_vp, _i32, _i64, _i64, _vp, _i32, _i64, _i64, _vp, _i32, _i32, _i32, _f32, _f32,
                                _vp, _f32, _vp, _vp, _vp, _sz, _vp]),
    "kd_sparse_fwd_bwd": (_i32, [_vp, _i32, _i64, _i64, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _f32, _f32,
                                 _vp, _f32, _vp, _vp, _vp, _sz, _vp]),
    "kd_scale_inplace": (_i32, [_vp, _i32, _i64, _vp, _vp]),
    "kd_topk_logprobs": (_i32, [_vp, _i32, _i64, _i32, _i64, _i32, _vp, _vp, _vp]),
    "kd_topk_workspace_bytes": (_sz, [_i64, _i32]),
    "kd_topk_logprobs_ws": (_i32, [_vp, _i32, _i64, _i32, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "kd_mask_rows": (_i32, [_vp, _i32, _i64, _i64, _vp]),
    "kd_probe_read_bandwidth": (_i32, [_vp, _sz, _i32, _i32, _i32, _vp, _vp]),
    "kd_multimem_allreduce": (_i32, [_vp, _sz, _sz, _i32, _i32, _i32, _i32, _vp]),
    "kd_fused_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32]),
    "kd_fused_logit_cache_bytes": (_sz, [_i32, _i32, _i32, _sz]),
    "kd_compact_rows": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "kd_gather_rows": (_i32, [_vp, _i64, _vp, _i32, _vp, _i64, _i64, _i32, _vp]),
    "kd_zero_if_empty": (_i32, [_vp, _i64, _vp, _vp]),
    "kd_fused_linear_fwd": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _i32,
                                   _i32, _f32, _f32, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "kd_fused_linear_bwd": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _i32,
                                   _i32, _i32, _f32, _vp, _vp, _i32, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _sz,
                                   _vp, _sz, _vp]),
    "kd_fused_linear_bwd_range": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _vp,
                                         _i32, _i32, _i32, _f32, _vp, _vp, _i32, _vp, _i64, _vp, _i64, _i64, _i32,
                                         _i32, _i32, _i32, _i32, _i32, _vp, _sz, _vp, _sz, _vp, _vp]),
    "kd_fused_linear_fwd_partial": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp,
                                           _i32, _i32, _i32, _i32, _f32, _vp, _vp, _sz, _vp, _sz, _vp]),
    "kd_fused_merge_workspace_bytes": (_sz, []),
    "kd_fused_merge_ranks": (_i32, [_vp, _i32, _vp, _i32, _i32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "kd_ce_fused_linear_fwd": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp, _sz,
                                      _vp]),
    "kd_ce_fused_linear_bwd": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _i64,
                                      _vp, _i64, _i64, _i32, _vp, _sz, _vp, _sz, _vp]),
    "kd_linear_bf16": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp]),
    "kd_head_topk_layout": (_i32, [_i32, _c.POINTER(_i32), _c.POINTER(_i32)]),
    "kd_head_logits_stats": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _i32, _c.POINTER(_i32), _i32, _i32,
                                    _i32, _vp]),
    "kd_head_topk_select": (_i32, [_vp, _i64, _vp, _i32, _vp, _i32, _i32, _i64, _i32, _i32, _vp, _vp, _vp]),
    "kd_gemm_bf16": (_i32, [_vp, _i64, _i32, _vp, _i64, _i32, _vp, _i64, _i32, _i32, _i32, _vp]),
}

_lib = None


class KdError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KdError(
            f"{LIB_PATH} not found: build it with `python speech-distill_b200/build.py` "
            "(the KD kernels have no CPU or PyTorch fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.kd_version() != ABI_VERSION:
        raise KdError(f"libkd_b200 ABI version {lib.kd_version()} != {ABI_VERSION}")
    if os.environ.get("KD_NVTX", "1") != "0":
        _add_nvtx_ranges(lib)
    _lib = lib
    return lib


# entry points that only query (no kernels): no range
_NO_RANGE = {"kd_head_topk_layout", "kd_version", "kd_last_error", "kd_launch_count", "kd_device_info", "kd_stream_workspace_bytes",
             "kd_fused_workspace_bytes", "kd_fused_logit_cache_bytes", "kd_fused_merge_workspace_bytes",
             "kd_fused_bwd_trace_begin", "kd_fused_bwd_trace_read", "kd_topk_workspace_bytes"}


def _add_nvtx_ranges(lib):
    """NVTX range around every launching C-ABI call (tracing row): a timeline tool shows the path's calls
    by name.  A push / pop pair costs well under a microsecond; KD_NVTX=0 leaves the raw ctypes functions."""
    try:
        import torch

        if not torch.cuda.is_available():
            return
        push, pop = torch.cuda.nvtx.range_push, torch.cuda.nvtx.range_pop
        push("kd_b200:init")
        pop()
    except Exception:  # NVTX is optional tooling, never a reason to fail
        return

    def wrap(name, fn):
        def call(*args):
            push(name)
            try:
                return fn(*args)
            finally:
                pop()

        call.__name__ = name
        return call

    for name in SIGNATURES:
        if name not in _NO_RANGE:
            setattr(lib, name, wrap(name, getattr(lib, name)))


def check(rc, what):
    if rc != 0:
        msg = load().kd_last_error().decode("utf-8", "replace")
        raise KdError(f"{what} failed (code {rc}): {msg}")


def dtype_code(t):
    import torch

    if t == torch.float32:
        return KD_DTYPE_F32
    if t == torch.bfloat16:
        return KD_DTYPE_BF16
    if t == torch.float16:
        return KD_DTYPE_F16
    raise TypeError(f"unsupported dtype {t}; the KD kernels take float32, bfloat16 or float16")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise KdError(
                "speech_distill_b200 runs on CUDA tensors only (sm_100a kernels, no CPU fallback); "
                f"got a tensor on {t.device}"
            )


def stream_ptr(device):
    import torch

    return torch.cuda.current_stream(device).cuda_stream
