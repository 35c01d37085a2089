"""CPU-side checks of the C-ABI boundary: the library builds, loads and exports every symbol that
include/kd_b200.h declares; the Python shim refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "kd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kd_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    names = _declared()
    for must in ("kd_dense_fwd_bwd", "kd_sparse_fwd_bwd", "kd_topk_logprobs", "kd_fused_linear_fwd",
                 "kd_fused_linear_bwd", "kd_mask_rows", "kd_prepare_rows", "kd_finalize_losses", "kd_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import speech_distill_b200 as K
    from speech_distill_b200 import _lib

    if not os.path.exists(K.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = ctypes.CDLL(K.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in kd_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(_declared())
    header = open(os.path.join(ROOT, "include", "kd_b200.h")).read()
    abi = int(re.search(r"#define KD_ABI_VERSION (\d+)", header).group(1))
    assert K.load_library().kd_version() == abi == _lib.ABI_VERSION
    assert lib.kd_stream_workspace_bytes() > 0


def test_no_cpu_fallback():
    import speech_distill_b200 as K

    z = torch.randn(1, 4, 16)
    lab = torch.randint(0, 16, (1, 4))
    with pytest.raises(K.KdError):
        K.DistillationLoss()(z, lab, teacher_logits=z)
    with pytest.raises(K.KdError):
        K.teacher_topk_logprobs(z, 4)
    with pytest.raises(K.KdError):
        K.fused_linear_kd_loss(torch.randn(1, 4, 8).bfloat16(), torch.randn(16, 8).bfloat16(), lab, teacher_logits=z)


def test_value_error_matches_reference():
    import speech_distill_b200 as K

    if torch.cuda.is_available():
        z = torch.randn(1, 4, 16, device="cuda")
        with pytest.raises(ValueError, match="Either teacher_logits or top_k must be provided"):
            K.DistillationLoss()(z, torch.zeros(1, 4, dtype=torch.long, device="cuda"))


def test_dropin_module_name():
    import distillation_loss
    import speech_distill_b200 as K

    assert distillation_loss.DistillationLoss is K.DistillationLoss
    m = distillation_loss.DistillationLoss(temperature=3.0, alpha=0.25)
    assert m.temperature == 3.0 and m.alpha == 0.25 and m.ignore_index == -100
    assert list(m.parameters()) == [] and list(m.buffers()) == []  # stateless like the reference


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "speech-distill_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read(), fn
    assert "oracle" not in open(os.path.join(ROOT, "distillation_loss.py")).read()


def test_freeze_model_weights_host_logic():
    """stage1.py:29-93 semantics on CPU tensors (hook logic only; the CUDA memset is covered on GPU)."""
    import numpy as np
    import torch.nn as nn
    import torch.nn.functional as F
    import speech_distill_b200 as K

    class Toy(nn.Module):
        def __init__(self, V, H):
            super().__init__()
            self.emb = nn.Embedding(V, H)
            self.mid = nn.Linear(H, H)
            self.head = nn.Linear(H, V, bias=False)
            self.head.weight = self.emb.weight

        def get_input_embeddings(self):
            return self.emb

        def get_output_embeddings(self):
            return self.head

        def forward(self, ids):
            return self.head(torch.tanh(self.mid(self.emb(ids))))

    d = np.load(os.path.join(ROOT, "tests", "golden", "stage1_mask.npz"))
    torch.manual_seed(7)
    V, H, new = 50, 16, 6
    toy = Toy(V, H).double()
    K.freeze_model_weights(toy, new)
    ids = torch.randint(0, V, (2, 9))
    loss = F.cross_entropy(toy(ids)[:, :-1].reshape(-1, V), ids[:, 1:].reshape(-1))
    loss.backward()
    np.testing.assert_allclose(toy.emb.weight.grad.numpy(), d["masked"], rtol=1e-12, atol=1e-15)
    assert [n for n, p in toy.named_parameters() if p.requires_grad] == ["emb.weight"]


def test_header_is_plain_c():
    """include/kd_b200.h is the drop-in boundary: it must compile as C (no C++ / CUDA / torch types) and every
    prototype must be callable from a C translation unit that links against the library."""
    import shutil
    import subprocess
    import tempfile

    import speech_distill_b200 as K

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = r'''
#include "kd_b200.h"
#include <stdio.h>
int main(void) {
  /* take the address of every entry point: unresolved or mis-declared symbols fail at link time */
  void* fns[] = {(void*)kd_version, (void*)kd_last_error, (void*)kd_launch_count, (void*)kd_fused_bwd_trace_begin,
                 (void*)kd_fused_bwd_trace_read, (void*)kd_device_info, (void*)kd_prepare_rows,
                 (void*)kd_finalize_losses, (void*)kd_stream_workspace_bytes, (void*)kd_dense_fwd_bwd,
                 (void*)kd_sparse_fwd_bwd, (void*)kd_scale_inplace, (void*)kd_topk_logprobs, (void*)kd_topk_workspace_bytes,
                 (void*)kd_topk_logprobs_ws, (void*)kd_probe_read_bandwidth, (void*)kd_multimem_allreduce,
                 (void*)kd_mask_rows,
                 (void*)kd_compact_rows, (void*)kd_gather_rows, (void*)kd_zero_if_empty,
                 (void*)kd_fused_workspace_bytes, (void*)kd_fused_logit_cache_bytes, (void*)kd_fused_linear_fwd,
                 (void*)kd_fused_linear_bwd,
                 (void*)kd_fused_linear_bwd_range, (void*)kd_fused_linear_fwd_partial,
                 (void*)kd_fused_merge_workspace_bytes, (void*)kd_fused_merge_ranks, (void*)kd_ce_fused_linear_fwd,
                 (void*)kd_ce_fused_linear_bwd, (void*)kd_linear_bf16, (void*)kd_head_topk_layout,
                 (void*)kd_head_logits_stats, (void*)kd_head_topk_select, (void*)kd_gemm_bf16};
  printf("%d %d\n", kd_version(), (int)(sizeof(fns) / sizeof(fns[0])));
  return kd_version() == KD_ABI_VERSION ? 0 : 1;
}
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "abi.c")
        exe = os.path.join(td, "abi")
        open(c, "w").write(src)
        libdir = os.path.dirname(K.LIB_PATH)
        subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o", exe,
                        "-L", libdir, "-l:libkd_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
        assert int(out[0]) >= 2 and int(out[1]) == len(_declared())


def test_ctypes_signatures_match_header():
    """Every prototype in include/kd_b200.h has the argument kinds (pointer / int / int64 / size_t / float) the
    ctypes table of speech_distill_b200._lib declares, in the same order: a mismatch is a host-side crash."""
    import ctypes as c

    from speech_distill_b200 import _lib as L

    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "kd_b200.h")).read(), flags=re.S)
    kinds_of = {c.c_void_p: "ptr", c.c_char_p: "ptr", c.c_int: "int", c.c_int64: "int64", c.c_size_t: "size",
                c.c_float: "float", c.c_ulonglong: "size"}  # 64-bit unsigned either way
    for name, (_, args) in L.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", header, flags=re.S)
        assert m, f"{name} is not declared in kd_b200.h"
        want = []
        for p in (x.strip() for x in m.group(1).split(",")):
            if not p or p == "void":
                continue
            if "*" in p:
                want.append("ptr")
            elif p.startswith("int64_t"):
                want.append("int64")
            elif p.startswith("size_t"):
                want.append("size")
            elif p.startswith("float"):
                want.append("float")
            elif p.startswith("unsigned long long"):
                want.append("size")
            else:
                want.append("int")
        got = [kinds_of.get(a, "ptr") for a in args]
        assert got == want, f"{name}: ctypes {got} != header {want}"


def test_custom_ops_registered_with_fake_kernels():
    """speech_distill_b200.ops registers the loss entry points with torch.library; their fake (meta) kernels give the
    output shapes a compiler traces with - checked here on fake CUDA tensors, no GPU needed."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode

    import speech_distill_b200  # noqa: F401
    from speech_distill_b200 import ops  # noqa: F401

    ns = torch.ops.speech_distill_b200
    for name in ("kd_loss", "fused_linear_kd", "fused_linear_kd_bwd"):
        assert hasattr(ns, name)
    with FakeTensorMode():
        h = torch.empty(2, 8, 64, dtype=torch.bfloat16, device="cuda")
        W = torch.empty(1000, 64, dtype=torch.bfloat16, device="cuda")
        lab = torch.empty(2, 8, dtype=torch.int64, device="cuda")
        y = torch.empty(2, 8, 1000, dtype=torch.bfloat16, device="cuda")
        losses, row_stats, row_target, n_valid, cache = ns.fused_linear_kd(h, W, lab, y, None, None, None, 2.0, 0.5, -100,
                                                                           1536.0, 0)
        assert tuple(losses.shape) == (4,) and losses.dtype == torch.float32
        assert tuple(row_stats.shape) == (16, 4) and tuple(row_target.shape) == (16,) and cache.dtype == torch.uint8
        dH, dW = ns.fused_linear_kd_bwd(h, W, y, None, None, row_stats, row_target, n_valid, cache, losses[0], 2.0, 0.5, 0)
        assert dH.shape == h.shape and dW.shape == W.shape
        z = torch.empty(2, 8, 1000, dtype=torch.bfloat16, device="cuda")
        l4, dz = ns.kd_loss(z, lab, y, None, None, None, 2.0, 0.5, -100)
        assert tuple(l4.shape) == (4,) and dz.shape == z.shape and dz.dtype == z.dtype
