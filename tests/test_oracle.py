"""Pins oracle/kd_oracle.py to the golden vectors produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import kd_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOSS_FILES = sorted(glob.glob(os.path.join(GOLDEN, "loss_*.npz")))


def _kw(d):
    kw = dict(temperature=float(d["tau"]), alpha=float(d["alpha"]))
    if "speech" in d.files:
        kw["speech_token_mask"] = d["speech"]
    if str(d["mode"]) == "dense":
        kw["teacher_logits"] = d["y"]
    else:
        kw["teacher_top_k_v"] = d["v"]
        kw["teacher_top_k_i"] = d["i"]
    return kw


def _torch_kw(kw):
    return {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()}


def test_golden_files_present():
    assert len(LOSS_FILES) >= 11
    for f in ("fused_dense_f64.npz", "topk_f32.npz", "stage1_mask.npz"):
        assert os.path.exists(os.path.join(GOLDEN, f))


@pytest.mark.parametrize("path", LOSS_FILES, ids=[os.path.basename(p)[:-4] for p in LOSS_FILES])
def test_torch_restatement_matches_reference(path):
    d = np.load(path)
    z = torch.from_numpy(d["z"])
    (losses, grad) = O.reference_loss_and_grad(z, torch.from_numpy(d["labels"]), **_torch_kw(_kw(d)))
    got = np.array([float(x.detach()) for x in losses])
    # same op sequence, same dtype -> bitwise equal scalars and gradient
    np.testing.assert_array_equal(got, d["losses"])
    np.testing.assert_array_equal(grad.numpy(), d["grad"])


@pytest.mark.parametrize("path", LOSS_FILES, ids=[os.path.basename(p)[:-4] for p in LOSS_FILES])
def test_closed_form_matches_reference(path):
    d = np.load(path)
    r = O.closed_form(d["z"], d["labels"], **_kw(d))
    f64 = d["z"].dtype == np.float64
    # sparse: the reference does its K-side softmax in fp32 (distillation_loss.py:82-95)
    rtol = (1e-12 if str(d["mode"]) == "dense" else 2e-6) if f64 else 2e-5
    np.testing.assert_allclose(np.array(r["losses"]), d["losses"], rtol=rtol, atol=1e-12)
    scale = max(np.abs(d["grad"]).max(), 1e-30)
    assert np.abs(r["grad"] - d["grad"]).max() / scale < (1e-9 if f64 and str(d["mode"]) == "dense" else 5e-5)


def test_closed_form_row_bookkeeping():
    d = np.load(os.path.join(GOLDEN, "loss_dense_f64_mixed.npz"))
    r = O.closed_form(d["z"], d["labels"], **_kw(d))
    lab = d["labels"]
    ok = (lab[:, 1:] != -100) & (d["speech"][:, 1:] != 0)
    assert r["sums"][3] == ok.sum()
    # invalid rows and the last position of every sequence get exactly zero gradient
    g = r["grad"]
    assert np.all(g[:, -1, :] == 0)
    assert np.all(g[:, :-1][~ok] == 0)


def test_fused_linear_reference():
    d = np.load(os.path.join(GOLDEN, "fused_dense_f64.npz"))
    losses, gh, gw = O.fused_linear_reference(
        torch.from_numpy(d["h"]), torch.from_numpy(d["W"]), torch.from_numpy(d["labels"]),
        teacher_logits=torch.from_numpy(d["y"]), temperature=float(d["tau"]), alpha=float(d["alpha"]),
    )
    np.testing.assert_allclose([float(x.detach()) for x in losses], d["losses"], rtol=1e-13)
    np.testing.assert_allclose(gh.numpy(), d["dh"], rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose(gw.numpy(), d["dW"], rtol=1e-10, atol=1e-15)
    # dH = G W, dW = G^T h with the closed-form G (math sheet)
    z = d["h"] @ d["W"].T
    r = O.closed_form(z, d["labels"], teacher_logits=d["y"], temperature=float(d["tau"]), alpha=float(d["alpha"]))
    G = r["grad"].reshape(-1, z.shape[-1])
    np.testing.assert_allclose(G @ d["W"], d["dh"].reshape(-1, d["h"].shape[-1]), rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(G.T @ d["h"].reshape(-1, d["h"].shape[-1]), d["dW"], rtol=1e-9, atol=1e-14)


def test_topk_oracles_match_reference():
    d = np.load(os.path.join(GOLDEN, "topk_f32.npz"))
    logits = torch.from_numpy(d["logits"])
    k = int(d["k"])
    v, i = O.topk_logprobs_reference(logits, k)
    np.testing.assert_array_equal(i.numpy(), d["i"])
    np.testing.assert_array_equal(v.numpy(), d["v"])
    # the kernel's deterministic spec is index-exact on tie-free inputs; values within 1 fp16 ulp
    v2, i2 = O.topk_spec(logits, k)
    np.testing.assert_array_equal(i2.numpy(), d["i"])
    assert np.abs(v2.float().numpy() - d["v"].astype(np.float32)).max() <= 2 ** -7


def test_stage1_mask_oracle():
    d = np.load(os.path.join(GOLDEN, "stage1_mask.npz"))
    old = int(d["old_vocab"])
    got = O.mask_old_rows(torch.from_numpy(d["unmasked"]), old).numpy()
    np.testing.assert_array_equal(got, d["masked"])
    assert np.all(d["masked"][:old] == 0) and np.abs(d["masked"][old:]).max() > 0
    assert int(d["n_trainable"]) == 1


def test_stage1_ce_reference_masks_rows():
    g = torch.Generator().manual_seed(3)
    h = torch.randn(2, 6, 8, generator=g, dtype=torch.float64)
    w = torch.randn(20, 8, generator=g, dtype=torch.float64)
    lab = torch.randint(0, 20, (2, 6), generator=g)
    lab[0, :2] = -100
    loss, gh, gw = O.stage1_ce_reference(h, w, lab, old_vocab_size=15)
    assert torch.all(gw[:15] == 0) and gw[15:].abs().max() > 0 and gh.abs().max() > 0
    r = O.closed_form((h @ w.t()).numpy(), lab.numpy(), teacher_logits=(h @ w.t()).numpy(), temperature=1.0, alpha=1.0)
    assert abs(r["losses"][1] - float(loss)) < 1e-6  # HF upcasts logits to fp32


# ---- the oracle against the reference's CALLER (unmodified compute_loss, oracle/make_golden_flow.py) ---------------
import ast  # noqa: E402


def _load_flow_model(d, tag):
    from transformers import Qwen3Config, Qwen3ForCausalLM

    cfg = ast.literal_eval(str(d[f"{tag}_cfg"]))
    model = Qwen3ForCausalLM(Qwen3Config(**cfg)).float()
    sd = {k[len(tag) + 1:]: torch.from_numpy(d[k]).view(torch.bfloat16).float() for k in d.files if k.startswith(tag + "/")}
    model.load_state_dict(sd)
    return model.eval()


@pytest.mark.parametrize("name", ["onthefly_topk16", "dense_teacher"])
def test_oracle_reproduces_reference_compute_loss(name):
    """tests/golden/flow_*.npz hold what reference train.py:43-116 returned and logged for a tiny fp32 Qwen3 pair.
    The oracle, fed with the same models' logits (and the top-k cache built as train.py:82-91 does), gives the same
    four numbers and the same gradient of the student's LM head: the oracle is pinned on the caller as well."""
    d = np.load(os.path.join(GOLDEN, f"flow_{name}.npz"))
    student, teacher = _load_flow_model(d, "student"), _load_flow_model(d, "teacher")
    ids = torch.from_numpy(d["input_ids"])
    labels = torch.from_numpy(d["labels"])
    mask = torch.from_numpy(d["speech_token_mask"])
    top_k = int(d["top_k"])
    with torch.no_grad():
        y = teacher(input_ids=ids).logits
    kw = {}
    z = student(input_ids=ids).logits
    if top_k > 0:
        lp = torch.log_softmax(y[..., : z.size(-1)], dim=-1)
        v, i = torch.topk(lp, top_k, dim=-1)
        kw = dict(teacher_top_k_v=v.to(torch.float16), teacher_top_k_i=i.to(torch.int32))
    else:
        kw = dict(teacher_logits=y)
    out = O.reference_loss(z, labels, speech_token_mask=mask, temperature=2.0, alpha=0.5, **kw)
    out[0].backward()
    want = [float(d["loss"]), float(d["student_loss"]), float(d["distill_loss"]), float(d["teacher_loss"])]
    np.testing.assert_allclose([float(o.detach()) for o in out], want, rtol=2e-6)
    np.testing.assert_allclose(student.lm_head.weight.grad.numpy(), d["grad_lm_head"], rtol=1e-4, atol=1e-8)
