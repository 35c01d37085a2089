"""CPU oracle for the speech-distill KD hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path
(``speech-distill_b200/``) never does and fails loudly without its CUDA library.

Parity status: **pinned against the reference itself, run in the build container.**
The reference (``/root/reference``) ships no tests, golden vectors or known-answer
fixtures, so the pins are outputs of the unmodified reference
``distillation_loss.DistillationLoss`` / ``stage1.freeze_model_weights`` hooks and the
three torch lines of ``extract_teacher_logits.py`` executed here by
``oracle/make_golden.py``; the resulting vectors live in ``tests/golden/*.npz`` and
``tests/test_oracle.py`` checks every function below against them.

Two independent restatements are kept on purpose:

* ``reference_loss`` - the reference's op sequence in torch (same softmax /
  log_softmax / kl_div / cross_entropy calls, so the same rounding behaviour in every
  dtype).  This is also what ``bench.py`` times as the CPU baseline.
* ``closed_form`` - the single-pass math sheet in numpy fp64,
  i.e. the formulas the CUDA kernels implement, including the analytic gradient.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

IGNORE_INDEX = -100


# --------------------------------------------------------------------------------------
# 1. torch restatement of distillation_loss.py:14-128
# --------------------------------------------------------------------------------------
def valid_rows(labels, speech_token_mask=None, ignore_index=IGNORE_INDEX):
    """Row predicate of distillation_loss.py:34-41 (causal shift: row t scores label t+1)."""
    nxt = labels[..., 1:].reshape(-1)
    ok = nxt != ignore_index
    if speech_token_mask is not None:
        ok = speech_token_mask[..., 1:].reshape(-1).bool() & ok
    return ok, nxt


def reference_loss(
    student_logits,
    labels,
    teacher_logits=None,
    teacher_top_k_v=None,
    teacher_top_k_i=None,
    speech_token_mask=None,
    temperature=2.0,
    alpha=0.5,
    ignore_index=IGNORE_INDEX,
):
    """(total, task, distill, teacher_task) exactly as distillation_loss.py:14-128 forms them."""
    tau = temperature
    V = student_logits.size(-1)
    ok, nxt = valid_rows(labels, speech_token_mask, ignore_index)
    s = student_logits[..., :-1, :].reshape(-1, V)[ok]  # :31-33,44
    tgt = nxt[ok]  # :45
    dev = student_logits.device
    if s.size(0) == 0:  # :47-53
        z = lambda: torch.tensor(0.0, device=dev)
        return z(), z(), z(), z()

    if teacher_logits is not None:  # dense, :56-71
        t = teacher_logits[..., :-1, :].reshape(-1, teacher_logits.size(-1)).detach()[ok]
        p = F.softmax(t / tau, dim=-1)
        logq = F.log_softmax(s / tau, dim=-1)
        distill = F.kl_div(logq, p, reduction="batchmean") * (tau**2)
        teacher_task = F.cross_entropy(t, tgt)
    elif teacher_top_k_v is not None and teacher_top_k_i is not None:  # sparse, :73-118
        K = teacher_top_k_v.size(-1)
        v = teacher_top_k_v[..., :-1, :].reshape(-1, K)[ok].to(dev, dtype=torch.float32)
        idx = teacher_top_k_i[..., :-1, :].reshape(-1, K)[ok].to(dev).long()
        p = F.softmax(v / tau, dim=-1)
        logp = F.log_softmax(v / tau, dim=-1)
        logq = F.log_softmax(s / tau, dim=-1).gather(-1, idx)
        distill = (p * (logp - logq)).sum(-1).mean() * (tau**2)
        hit = (idx == tgt.unsqueeze(-1)).nonzero(as_tuple=True)
        if hit[0].numel() > 0:
            teacher_task = -v[hit[0], hit[1]].mean()
        else:
            teacher_task = torch.tensor(0.0, device=dev)
    else:
        raise ValueError("Either teacher_logits or top_k must be provided")  # :120

    task = F.cross_entropy(s, tgt)  # :123
    total = alpha * task + (1 - alpha) * distill  # :126
    return total, task, distill, teacher_task


def reference_loss_and_grad(student_logits, labels, **kw):
    """Forward + backward to d(total)/d(student_logits) through torch autograd."""
    z = student_logits.detach().clone().requires_grad_(True)
    out = reference_loss(z, labels, **kw)
    if out[0].requires_grad:
        out[0].backward()
        g = z.grad
    else:  # N == 0: the reference returns constants without a graph
        g = torch.zeros_like(z)
    return tuple(o.detach() for o in out), g


def fused_linear_reference(hidden, weight, labels, **kw):
    """LM head (logits = hidden @ weight^T, HF nn.Linear without bias) followed by the loss.

    Returns the 4 scalars and (dHidden, dWeight); dtype of the arithmetic = dtype of the inputs
    (tests pass fp32/fp64 copies of bf16-rounded tensors, SURVEY.md 8d "parity protocol").
    """
    h = hidden.detach().clone().requires_grad_(True)
    w = weight.detach().clone().requires_grad_(True)
    out = reference_loss(h @ w.t(), labels, **kw)
    if out[0].requires_grad:
        out[0].backward()
        gh, gw = h.grad, w.grad
    else:
        gh, gw = torch.zeros_like(h), torch.zeros_like(w)
    return tuple(o.detach() for o in out), gh, gw


# --------------------------------------------------------------------------------------
# 2. closed form (what the kernels compute), numpy fp64
# --------------------------------------------------------------------------------------
def _lse(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    return (m + np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True))).squeeze(axis)


def closed_form(
    student_logits,
    labels,
    teacher_logits=None,
    teacher_top_k_v=None,
    teacher_top_k_i=None,
    speech_token_mask=None,
    temperature=2.0,
    alpha=0.5,
    ignore_index=IGNORE_INDEX,
):
    """Single-pass formulas + analytic gradient in fp64.

    Returns dict(sums=[sum_ce, sum_kl, sum_teacher, n_valid, n_hits], losses=(total, task,
    distill, teacher_task), grad=[B,T,V] d(total)/d(student_logits)).  ``sum_teacher`` is the
    teacher-CE row sum (dense) or the sum of hit log-probs (sparse).
    """
    z = np.asarray(student_logits, dtype=np.float64)
    lab = np.asarray(labels)
    tau = float(temperature)
    lead, V = z.shape[:-1], z.shape[-1]
    T = lead[-1]
    z2 = z.reshape(-1, T, V)
    lab2 = lab.reshape(-1, T)
    B = z2.shape[0]
    ok = np.zeros((B, T), dtype=bool)
    ok[:, :-1] = lab2[:, 1:] != ignore_index
    if speech_token_mask is not None:
        m = np.asarray(speech_token_mask).reshape(-1, T)
        ok[:, :-1] &= m[:, 1:] != 0
    tgt = np.zeros((B, T), dtype=np.int64)
    tgt[:, :-1] = lab2[:, 1:]
    n = int(ok.sum())
    grad = np.zeros_like(z2)
    if n == 0:
        return dict(sums=np.zeros(5), losses=(0.0, 0.0, 0.0, 0.0), grad=grad.reshape(z.shape))
    rows = np.nonzero(ok)
    zr = z2[rows]  # [n, V]
    lr = tgt[rows]
    ar = np.arange(n)
    lse1 = _lse(zr)
    lset = _lse(zr / tau)
    ce = lse1 - zr[ar, lr]
    q1 = np.exp(zr - lse1[:, None])
    qt = np.exp(zr / tau - lset[:, None])
    n_hits = 0.0
    if teacher_logits is not None:
        y = np.asarray(teacher_logits, dtype=np.float64).reshape(-1, T, V)[rows]
        lsett = _lse(y / tau)
        P = np.exp(y / tau - lsett[:, None])
        with np.errstate(invalid="ignore", divide="ignore"):
            cross = np.where(P > 0, P * (y - zr) / tau, 0.0).sum(-1)
        kl = cross - lsett + lset
        tsum = float((_lse(y) - y[ar, lr]).sum())
        teacher_task = tsum / n
    elif teacher_top_k_v is not None and teacher_top_k_i is not None:
        K = np.asarray(teacher_top_k_v).shape[-1]
        # the reference computes the K-side softmax in fp32 (distillation_loss.py:82-95)
        v = np.asarray(teacher_top_k_v, dtype=np.float32).astype(np.float64).reshape(-1, T, K)[rows]
        idx = np.asarray(teacher_top_k_i).reshape(-1, T, K)[rows].astype(np.int64)
        lk = _lse(v / tau)
        pk = np.exp(v / tau - lk[:, None])
        logpk = v / tau - lk[:, None]
        zg = np.take_along_axis(zr, idx, axis=-1)
        kl = (pk * (logpk - (zg / tau - lset[:, None]))).sum(-1)
        P = np.zeros_like(zr)
        np.add.at(P, (ar[:, None].repeat(K, 1), idx), pk)  # duplicates accumulate (autograd scatter-add)
        hit = idx == lr[:, None]
        n_hits = float(hit.sum())
        tsum = float(v[hit].sum())
        teacher_task = -tsum / n_hits if n_hits > 0 else 0.0
    else:
        raise ValueError("Either teacher_logits or top_k must be provided")
    onehot = np.zeros_like(zr)
    onehot[ar, lr] = 1.0
    g = (alpha * (q1 - onehot) + (1.0 - alpha) * tau * (qt - P)) / n
    grad[rows] = g
    task = float(ce.sum()) / n
    distill = tau * tau * float(kl.sum()) / n
    total = alpha * task + (1 - alpha) * distill
    return dict(
        sums=np.array([ce.sum(), kl.sum(), tsum, float(n), n_hits]),
        losses=(total, task, distill, teacher_task),
        grad=grad.reshape(z.shape),
    )


# --------------------------------------------------------------------------------------
# 3. teacher top-k compaction (train.py:82-91, extract_teacher_logits.py:114-129)
# --------------------------------------------------------------------------------------
def topk_logprobs_reference(logits, k):
    """log_softmax -> topk -> fp16 values / int32 indices, the reference's three torch calls."""
    lp = F.log_softmax(logits, dim=-1)
    v, i = torch.topk(lp, k=k, dim=-1)
    return v.to(torch.float16), i.to(torch.int32)


def topk_spec(logits, k):
    """Deterministic spec of the CUDA kernel: select/sort by (raw logit desc, index asc);
    value = round_to_logits_dtype((x - max) - log(sum exp(x - max))) in fp32, then fp16.
    Equals ``topk_logprobs_reference`` index-for-index whenever the top-k logits are distinct
    (hard part 4)."""
    x = logits.detach().to(torch.float32).cpu().numpy()
    R = x.reshape(-1, x.shape[-1])
    idx = np.empty((R.shape[0], k), dtype=np.int32)
    for r in range(R.shape[0]):
        order = np.lexsort((np.arange(R.shape[1]), -R[r].astype(np.float64)))
        idx[r] = order[:k]
    m = R.max(-1, keepdims=True)
    lse = np.log(np.exp((R - m).astype(np.float64)).sum(-1, keepdims=True))
    lp = ((R - m).astype(np.float64) - lse).astype(np.float32)
    vals = torch.from_numpy(np.take_along_axis(lp, idx.astype(np.int64), -1))
    vals = vals.to(logits.dtype).to(torch.float16)
    shp = tuple(logits.shape[:-1]) + (k,)
    return vals.reshape(shp), torch.from_numpy(idx).reshape(shp)


# --------------------------------------------------------------------------------------
# 4. stage1 frozen-vocab gradient row mask (stage1.py:53-57, 67-71)
# --------------------------------------------------------------------------------------
def mask_old_rows(grad, old_vocab_size):
    g = grad.clone()
    g[:old_vocab_size] = 0.0
    return g


def stage1_ce_reference(hidden, weight, labels, old_vocab_size, ignore_index=IGNORE_INDEX):
    """Causal-LM CE (transformers loss_utils.ForCausalLMLoss semantics: shift, ignore_index,
    mean over valid) through an LM head, with stage1's row mask applied to dWeight."""
    h = hidden.detach().clone().requires_grad_(True)
    w = weight.detach().clone().requires_grad_(True)
    V = w.shape[0]
    logits = (h @ w.t()).float()
    s = logits[..., :-1, :].reshape(-1, V)
    t = labels[..., 1:].reshape(-1)
    n = int((t != ignore_index).sum())
    if n == 0:
        return torch.zeros(()), torch.zeros_like(h), torch.zeros_like(w)
    loss = F.cross_entropy(s, t, ignore_index=ignore_index, reduction="mean")
    loss.backward()
    return loss.detach(), h.grad, mask_old_rows(w.grad, old_vocab_size)
