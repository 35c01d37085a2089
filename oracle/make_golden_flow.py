"""Golden vectors for the CALLER of the loss: the UNMODIFIED ``DistillationTrainer.compute_loss``
(reference ``train.py:43-116``) run in the build container on tiny random Qwen3 models, fp32 on the CPU.

    python oracle/make_golden_flow.py        # needs /root/reference; writes tests/golden/flow_*.npz

``compute_loss`` is called unbound with a SimpleNamespace standing in for the Trainer (transformers' Trainer cannot
be constructed here: no accelerate), exactly as SURVEY.md 8c describes.  Each fixture holds the model configs and
weights (bf16-representable values), the batch, and the reference's loss and gradients of the student's LM head
and embedding; tests/test_gpu_lazy.py::test_flow_matches_reference_run replays the batch through models patched by
enable_lazy_logits.  Nothing under tests/, bench.py or the package reads /root/reference at run time.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("KD_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_train():
    from transformers import Trainer  # noqa: F401  (must precede the peft stub)

    for name in ("peft", "trl", "s3tokenizer", "torchaudio", "librosa", "onnxruntime", "bitsandbytes", "wandb"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__getattr__ = lambda attr, _n=name: type(attr, (), {})
                sys.modules[name] = m
    sys.path.insert(0, REF)
    import train as ref_train

    assert os.path.abspath(ref_train.__file__).startswith(os.path.abspath(REF))
    return ref_train


def tiny_qwen3(vocab, hidden, seed, head_scale=2.0):
    from transformers import Qwen3Config, Qwen3ForCausalLM

    torch.manual_seed(seed)
    cfg = dict(vocab_size=vocab, hidden_size=hidden, intermediate_size=2 * hidden, num_hidden_layers=2,
               num_attention_heads=4, num_key_value_heads=2, head_dim=hidden // 4, max_position_embeddings=128,
               tie_word_embeddings=False)
    model = Qwen3ForCausalLM(Qwen3Config(**cfg)).float()
    with torch.no_grad():
        model.lm_head.weight.mul_(head_scale / (0.02 * hidden ** 0.5))
        for p in model.parameters():  # bf16-representable values: the GPU replay can hold them in bf16 exactly
            p.copy_(p.bfloat16().float())
    return model, cfg


def bf16_head_input(model):
    """Test double of a bf16 LM head on an fp32 body: the final norm's output is rounded to bf16 in the forward
    (straight-through in the backward), so `lm_head(hidden)` multiplies bf16-representable operands exactly as a
    tensor-core head does.  The model is a fixture; compute_loss / DistillationLoss stay the unmodified reference."""
    model.model.norm.register_forward_hook(lambda mod, args, out: out + (out.bfloat16().float() - out).detach())
    return model


def run(ref_train, name, top_k, with_cache, bf16_head=False):
    V, Vt, B, T = 512, 544, 2, 40
    student, scfg = tiny_qwen3(V, 32, 1)
    teacher, tcfg = tiny_qwen3(Vt if top_k > 0 else V, 64, 2)
    if bf16_head:
        bf16_head_input(student)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, V, (B, T), generator=g)
    labels = ids.clone()
    labels[:, :9] = -100
    mask = torch.ones(B, T)
    mask[1, 30:] = 0
    inputs = {"input_ids": ids.clone(), "attention_mask": torch.ones_like(ids), "labels": labels.clone(),
              "speech_token_mask": mask.clone()}
    extra = {}
    if with_cache:  # pre-computed cache as extract_teacher_logits.py:109-129 writes it
        with torch.no_grad():
            if bf16_head:  # a bf16 teacher (what the extractor runs on a GPU): bf16 hidden x bf16 W -> bf16 logits
                hid = teacher.model(input_ids=ids).last_hidden_state.bfloat16().float()
                logits = (hid @ teacher.lm_head.weight[:V].float().t()).bfloat16()
                lp = torch.log_softmax(logits, dim=-1)  # bf16 log-probs, as log_softmax of a bf16 tensor returns
            else:
                lp = torch.log_softmax(teacher(input_ids=ids).logits[..., :V], dim=-1)
            v, i = torch.topk(lp.float(), top_k, dim=-1)
        inputs["teacher_top_k_v"] = v.to(torch.float16)
        inputs["teacher_top_k_i"] = i.to(torch.int32)
        extra = {"teacher_top_k_v": inputs["teacher_top_k_v"].numpy(), "teacher_top_k_i": inputs["teacher_top_k_i"].numpy()}
    logged = {}
    self_ns = types.SimpleNamespace(
        teacher_model=teacher, top_k=top_k, is_quantized_teacher=False,
        distill_loss_fn=ref_train.DistillationLoss(temperature=2.0, alpha=0.5),
        state=types.SimpleNamespace(global_step=0), args=types.SimpleNamespace(logging_steps=1),
        log=lambda d: logged.update(d))
    loss = ref_train.DistillationTrainer.compute_loss(self_ns, student, inputs)
    loss.backward()
    out = {
        "student_cfg": np.array(repr(scfg)), "teacher_cfg": np.array(repr(tcfg)), "top_k": top_k,
        "input_ids": ids.numpy(), "labels": labels.numpy(), "speech_token_mask": mask.numpy(),
        "bf16_head_input": int(bf16_head), "loss": float(loss), "student_loss": logged["student_loss"], "teacher_loss": logged["teacher_loss"],
        "distill_loss": logged["distill_loss"],
        "grad_lm_head": student.lm_head.weight.grad.numpy(), "grad_embed": student.model.embed_tokens.weight.grad.numpy(),
    }
    for tag, m in (("student", student), ("teacher", teacher)):
        for k, v in m.state_dict().items():
            out[f"{tag}/{k}"] = v.bfloat16().view(torch.int16).numpy()  # exact: values are bf16-representable
    out.update(extra)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"flow_{name}.npz"), **out)
    print(name, "loss", float(loss), logged)


if __name__ == "__main__":
    rt = _import_train()
    run(rt, "onthefly_topk16", 16, False)
    run(rt, "dense_teacher", 0, False)
    # (a pre-computed cache, run(rt, "cached_topk16", 16, True), gives the same numbers as the on-the-fly run:
    #  the GPU test derives that case from the first fixture instead of storing the weights twice)
    # bf16 heads on both sides: the cache comes from a bf16 teacher head and the student's head input is
    # bf16-representable, so the replay through the tensor-core head has nothing left to differ by (1e-3 bar)
    run(rt, "bf16head_cached_topk16", 16, True, bf16_head=True)
