"""Logits that are never computed: the glue that lets the reference's ``train.py`` run unchanged on the fused path
(SURVEY.md 8f rank 1 and 2).

``enable_lazy_logits(model)`` replaces a causal LM's forward (reference call sites: ``train.py:54`` student,
``train.py:60-72`` teacher) by one that runs the transformer body only and returns ``outputs.logits`` as a
``LazyLogits`` = (last hidden states, lm_head weight).  Nothing of size [B,T,V] is allocated, and HF's own
redundant cross-entropy on ``labels`` is skipped.  ``compute_loss`` then does, unmodified:

* ``student_logits.size(-1)``                         -> answered from the weight's shape
* ``teacher_logits[..., :vocab_size]``                -> another LazyLogits over the first ``vocab_size`` weight rows
* ``F.log_softmax(.., dim=-1)``, ``torch.topk(.., k)`` -> ``teacher_head_topk`` (head GEMM + compaction kernels)
* ``DistillationLoss(student_logits=LazyLogits, ...)`` -> ``fused_linear_kd_loss`` (K1), dense or sparse teacher

Anything else asked of a LazyLogits (evaluation code reading ``outputs.logits``, arithmetic) materialises it
through ``kd_linear_bf16``.  All arithmetic stays in libkd_b200.so; this file is plumbing.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ._lib import KdError


class LazyLogits:
    """``hidden @ weight[:vocab].T`` as a value, not as a tensor.  ``log_probs=True`` marks the result of
    ``log_softmax`` over the last dimension (only ``topk`` is defined on it without materialising)."""

    def __init__(self, hidden, weight, vocab_size=None, log_probs=False):
        if hidden.shape[-1] != weight.shape[-1]:
            raise ValueError(f"hidden size {hidden.shape[-1]} != weight columns {weight.shape[-1]}")
        self.hidden = hidden
        self.weight = weight
        self.vocab_size = int(weight.shape[0] if vocab_size is None else min(vocab_size, weight.shape[0]))
        self.log_probs = bool(log_probs)

    # ---- tensor facade: metadata only ----
    @property
    def shape(self):
        return torch.Size(tuple(self.hidden.shape[:-1]) + (self.vocab_size,))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return self.hidden.dim()

    ndim = property(dim)

    @property
    def dtype(self):
        return self.hidden.dtype

    @property
    def device(self):
        return self.hidden.device

    @property
    def requires_grad(self):
        return bool(self.hidden.requires_grad or self.weight.requires_grad)

    def head_weight(self):
        """The weight rows that define these logits (a view)."""
        return self.weight if self.vocab_size == self.weight.shape[0] else self.weight[: self.vocab_size]

    def __repr__(self):
        return f"LazyLogits(shape={tuple(self.shape)}, dtype={self.dtype}, device={self.device}, log_probs={self.log_probs})"

    def __getitem__(self, idx):
        # the one indexing pattern of the path: logits[..., :vocab_size] (train.py:82-83)
        if isinstance(idx, tuple) and len(idx) >= 1 and isinstance(idx[-1], slice) and all(
                i is Ellipsis or (isinstance(i, slice) and i == slice(None)) for i in idx[:-1]):
            sl = idx[-1]
            if sl.start in (None, 0) and sl.step in (None, 1):
                stop = self.vocab_size if sl.stop is None else sl.stop
                if stop < 0:
                    stop += self.vocab_size
                return LazyLogits(self.hidden, self.weight, min(stop, self.vocab_size), self.log_probs)
        return self.materialize()[idx]

    # ---- the operations the path applies to logits ----
    def log_softmax(self, dim=-1, **_):
        if dim not in (-1, self.dim() - 1) or self.log_probs:
            return F.log_softmax(self.materialize(), dim=dim)
        return LazyLogits(self.hidden, self.weight, self.vocab_size, log_probs=True)

    def topk(self, k, dim=-1, largest=True, sorted=True):  # noqa: A002 - torch's keyword
        """``torch.topk(F.log_softmax(logits, -1), k)`` (train.py:85-88): values are the log-probs rounded to the
        logits dtype then to fp16 (exact for the reference's next line, ``.to(torch.float16)``), indices int64."""
        if not self.log_probs or dim not in (-1, self.dim() - 1) or not largest:
            return torch.topk(self.materialize(), k, dim=dim, largest=largest, sorted=sorted)
        from .topk import teacher_head_topk

        v, i = teacher_head_topk(self.hidden, self.head_weight(), int(k))
        return torch.return_types.topk((v, i.long()))

    def materialize(self):
        """The actual tensor (kd_linear_bf16; log_softmax applied by torch if this is a log-prob view)."""
        if not self.hidden.is_cuda:
            raise KdError("LazyLogits can only be materialised on a CUDA device (no CPU fallback)")
        from .topk import linear_bf16

        H = self.hidden.shape[-1]
        h2 = self.hidden.detach().reshape(-1, H).to(torch.bfloat16)
        out = linear_bf16(h2.contiguous(), self.head_weight().detach().to(torch.bfloat16))
        out = out.reshape(*self.hidden.shape[:-1], self.vocab_size)
        return F.log_softmax(out, dim=-1) if self.log_probs else out

    def detach(self):
        return self.materialize()

    def float(self):
        return self.materialize().float()

    def to(self, *a, **kw):
        return self.materialize().to(*a, **kw)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name == "log_softmax" and args and isinstance(args[0], LazyLogits):
            dim = kwargs.get("dim", args[1] if len(args) > 1 else -1)
            return args[0].log_softmax(dim=dim)
        if name == "topk" and args and isinstance(args[0], LazyLogits):
            k = kwargs.get("k", args[1] if len(args) > 1 else None)
            dim = kwargs.get("dim", args[2] if len(args) > 2 else -1)
            return args[0].topk(k, dim=dim, largest=kwargs.get("largest", True), sorted=kwargs.get("sorted", True))
        # anything else: fall back to the real tensor
        conv = lambda x: x.materialize() if isinstance(x, LazyLogits) else x  # noqa: E731
        return func(*[conv(a) for a in args], **{k: conv(v) for k, v in kwargs.items()})


def enable_lazy_logits(model):
    """Patch ``model.forward`` (a transformers ``*ForCausalLM``) to return ``outputs.logits`` as LazyLogits.

    The body (``model.model``) runs as before; the LM head and HF's internal loss are skipped (``loss`` is None:
    ``DistillationTrainer.compute_loss`` computes its own).  Returns the model; ``model._kd_original_forward``
    restores the stock behaviour."""
    body = getattr(model, "model", None) or model.base_model
    head = model.get_output_embeddings()
    if head is None or getattr(head, "bias", None) is not None:
        raise ValueError("enable_lazy_logits needs a bias-free output embedding (lm_head)")
    from transformers.modeling_outputs import CausalLMOutputWithPast

    def forward(input_ids=None, attention_mask=None, labels=None, **kw):
        for drop in ("logits_to_keep", "num_logits_to_keep", "num_items_in_batch"):
            kw.pop(drop, None)
        out = body(input_ids=input_ids, attention_mask=attention_mask, **kw)
        hidden = out.last_hidden_state if hasattr(out, "last_hidden_state") else out[0]
        return CausalLMOutputWithPast(loss=None, logits=LazyLogits(hidden, head.weight),
                                      past_key_values=getattr(out, "past_key_values", None),
                                      hidden_states=getattr(out, "hidden_states", None),
                                      attentions=getattr(out, "attentions", None))

    model._kd_original_forward = model.forward
    model.forward = forward
    return model


def enable_fused_ce(model, num_new_tokens=0):
    """Stage-1 counterpart of ``enable_lazy_logits`` (reference ``stage1.py:298-340``: TRL ``SFTTrainer`` calls the
    model with ``labels`` and takes ``outputs.loss``, computed by transformers' ``ForCausalLMLoss`` or by Liger's
    fused-linear-CE when ``use_liger_kernel=True``, ``stage1.py:315``).

    The patched forward runs the transformer body, then - when ``labels`` (or ``shift_labels``) are given - the fused
    LM-head cross-entropy kernels with ``dw_row_begin = V - num_new_tokens`` (the rows ``freeze_model_weights``
    masks, ``stage1.py:46-57``: they are never computed and stay exactly zero), and returns the loss with
    ``logits=LazyLogits``.  Same reduction as ``ForCausalLMLoss``: mean over the scored tokens, or
    sum / ``num_items_in_batch`` when the trainer passes it (gradient accumulation)."""
    from .loss import IGNORE_INDEX, fused_linear_kd_loss, prepare_rows  # noqa: F401

    body = getattr(model, "model", None) or model.base_model
    head = model.get_output_embeddings()
    if head is None or getattr(head, "bias", None) is not None:
        raise ValueError("enable_fused_ce needs a bias-free output embedding (lm_head)")
    from transformers.modeling_outputs import CausalLMOutputWithPast

    def forward(input_ids=None, attention_mask=None, labels=None, shift_labels=None, num_items_in_batch=None, **kw):
        for drop in ("logits_to_keep", "num_logits_to_keep"):
            kw.pop(drop, None)
        out = body(input_ids=input_ids, attention_mask=attention_mask, **kw)
        hidden = out.last_hidden_state if hasattr(out, "last_hidden_state") else out[0]
        loss = None
        if labels is not None or shift_labels is not None:
            if shift_labels is not None:
                # already shifted (TRL with padding-free batches): undo the shift the kernels apply (row t scores
                # labels[t + 1]) by prepending one ignored position
                lab = torch.nn.functional.pad(shift_labels, (1, 0), value=IGNORE_INDEX)[..., : shift_labels.size(-1)]
            else:
                lab = labels
            old_vocab = head.weight.size(0) - int(num_new_tokens) if num_new_tokens else 0
            total, _, _, _ = fused_linear_kd_loss(hidden, head.weight, lab, teacher_logits=None, temperature=1.0,
                                                  alpha=1.0, ignore_index=IGNORE_INDEX, dw_row_begin=old_vocab)
            loss = total
            if num_items_in_batch is not None:  # ForCausalLMLoss: reduction "sum" / num_items_in_batch
                B, T = lab.shape[0], lab.shape[-1]
                _, n_valid = prepare_rows(lab, None, B, T, IGNORE_INDEX, hidden.device)
                n_items = num_items_in_batch.to(hidden.device) if torch.is_tensor(num_items_in_batch) else num_items_in_batch
                loss = total * (n_valid.to(torch.float32).reshape(()) / n_items)
        return CausalLMOutputWithPast(loss=loss, logits=LazyLogits(hidden, head.weight),
                                      past_key_values=getattr(out, "past_key_values", None),
                                      hidden_states=getattr(out, "hidden_states", None),
                                      attentions=getattr(out, "attentions", None))

    model._kd_original_forward = model.forward
    model.forward = forward
    return model
