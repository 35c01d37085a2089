"""Stage-1 speech-token warm-up pieces (reference ``stage1.py:29-93`` and the causal-LM CE its
trainer runs, transformers ``loss_utils.ForCausalLMLoss``)."""
from __future__ import annotations

from . import _lib
from ._lib import check, dtype_code, require_cuda, stream_ptr
from .loss import fused_linear_kd_loss


def mask_old_rows_(grad, old_vocab_size):
    """In-place ``grad[:old_vocab_size] = 0`` on a [V, H] CUDA gradient (stage1.py:53-57 without the clone)."""
    require_cuda(grad)
    if not grad.is_contiguous():
        raise ValueError("mask_old_rows_ needs a contiguous [V, H] gradient")
    lib = _lib.load()
    n = max(0, min(int(old_vocab_size), grad.size(0)))
    check(lib.kd_mask_rows(grad.data_ptr(), dtype_code(grad.dtype), n, grad[0].numel(), stream_ptr(grad.device)),
          "kd_mask_rows")
    return grad


def freeze_model_weights(model, num_new_tokens, verbose=False):
    """Same contract as the reference ``freeze_model_weights`` (stage1.py:29-93): everything frozen,
    input/output embedding weights trainable, gradient rows of the old vocabulary forced to zero.
    The hook zeroes rows with one memset instead of clone + slice-assign; with the fused CE below
    (``dw_row_begin = old_vocab``) those rows are never computed and the hook is a no-op safety net."""
    for _, p in model.named_parameters():
        p.requires_grad = False
    if num_new_tokens > 0:
        emb = model.get_input_embeddings()
        old_vocab = emb.weight.size(0) - num_new_tokens
        seen = set()
        for layer in (emb, model.get_output_embeddings()):
            if layer is None:
                continue
            layer.weight.requires_grad_(True)
            if id(layer.weight) in seen:  # tied embeddings: one hook is enough (idempotent anyway)
                continue
            seen.add(id(layer.weight))

            def _hook(grad, _old=old_vocab):
                if grad is None:
                    return grad
                g = grad if grad.is_contiguous() else grad.contiguous()
                if g.is_cuda:
                    return mask_old_rows_(g.clone() if g is grad else g, _old)
                g = g.clone()  # CPU tensors (unit tests of the host logic): plain torch, no kernel involved
                g[:_old] = 0.0
                return g

            layer.weight.register_hook(_hook)
    if verbose:
        total = sum(p.numel() for p in model.parameters())
        new = num_new_tokens * model.get_input_embeddings().weight.size(1)
        print(f"trainable (new-token rows only): {new:,} of {total:,} parameters")
    return model


def fused_linear_cross_entropy(hidden, lm_head_weight, labels, old_vocab_size=0, ignore_index=-100, v_chunk=0):
    """Causal-LM cross-entropy through the LM head with the stage1 row mask folded into the dW GEMM.

    Equivalent to ``F.cross_entropy((hidden @ W.T)[..., :-1, :], labels[..., 1:], ignore_index)`` with
    ``dW[:old_vocab_size] = 0`` - but the logits are never materialised and the masked rows of dW are
    never computed.  Returns the mean loss (0-dim fp32)."""
    total, task, _, _ = fused_linear_kd_loss(hidden, lm_head_weight, labels, teacher_logits=None, temperature=1.0,
                                             alpha=1.0, ignore_index=ignore_index, dw_row_begin=old_vocab_size,
                                             v_chunk=v_chunk)
    return total
