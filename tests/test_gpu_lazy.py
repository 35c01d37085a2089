"""train.py's compute_loss, restated line by line, on models patched by enable_lazy_logits: neither the student's
nor the teacher's [B,T,V] logits exist, results equal the stock models + the reference loss."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import kd_oracle as O

pytestmark = pytest.mark.gpu


def _tiny_qwen3(vocab, hidden, seed):
    from transformers import Qwen3Config, Qwen3ForCausalLM

    torch.manual_seed(seed)
    cfg = Qwen3Config(vocab_size=vocab, hidden_size=hidden, intermediate_size=2 * hidden, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, head_dim=hidden // 4, max_position_embeddings=128,
                      tie_word_embeddings=False)
    model = Qwen3ForCausalLM(cfg)
    with torch.no_grad():  # logits with a spread of ~2 (a random-init head gives near-uniform, tie-ridden rows)
        model.lm_head.weight.mul_(2.0 / (0.02 * hidden ** 0.5))
    return model.to(device="cuda", dtype=torch.bfloat16)


def _compute_loss(student, teacher, loss_fn, inputs, top_k):
    """reference train.py:43-116 (DistillationTrainer.compute_loss) without the Trainer bookkeeping"""
    inputs = dict(inputs)
    speech_mask = inputs.pop("speech_token_mask", None)
    teacher_top_k_v = inputs.pop("teacher_top_k_v", None)
    teacher_top_k_i = inputs.pop("teacher_top_k_i", None)
    outputs = student(**inputs)                                     # :54
    student_logits = outputs.logits                                  # :55
    labels = inputs.pop("labels", None)                              # :56
    teacher_logits = None
    if teacher_top_k_v is None and teacher is not None:              # :60
        with torch.no_grad():
            teacher_logits = teacher(**inputs).logits                # :69-70
    if teacher_logits is not None and teacher_top_k_v is None and top_k > 0:   # :75-80
        with torch.no_grad():
            vocab_size = student_logits.size(-1)                     # :83
            teacher_logits_truncated = teacher_logits[..., :vocab_size]
            teacher_logprobs = F.log_softmax(teacher_logits_truncated, dim=-1)
            teacher_top_k_v, teacher_top_k_i = torch.topk(teacher_logprobs, k=top_k, dim=-1)
            teacher_top_k_v = teacher_top_k_v.to(torch.float16)      # :90
            teacher_top_k_i = teacher_top_k_i.to(torch.int32)        # :91
        teacher_logits = None
    return loss_fn(student_logits=student_logits, labels=labels, teacher_logits=teacher_logits,
                   teacher_top_k_v=teacher_top_k_v, teacher_top_k_i=teacher_top_k_i, speech_token_mask=speech_mask)


@pytest.mark.parametrize("top_k", [16, 0])
def test_compute_loss_flow_on_lazy_models(top_k):
    import speech_distill_b200 as K

    V, Vt, B, T = 1000, 1040, 2, 48
    student, teacher = _tiny_qwen3(V, 64, 1), _tiny_qwen3(Vt if top_k > 0 else V, 128, 2)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, V, (B, T), generator=g).cuda()
    labels = ids.clone()
    labels[:, :10] = -100
    mask = torch.ones(B, T)
    mask[1, 30:] = 0
    inputs = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels,
              "speech_token_mask": mask.cuda()}
    loss_fn = K.DistillationLoss(temperature=2.0, alpha=0.5)

    # stock models (logits materialised, K2) first, then the patched ones (K1 + head top-k)
    out_stock = _compute_loss(student, teacher, loss_fn, inputs, top_k)
    out_stock[0].backward()
    g_head = student.lm_head.weight.grad.clone()
    g_emb = student.model.embed_tokens.weight.grad.clone()
    student.zero_grad(set_to_none=True)

    K.enable_lazy_logits(student)
    K.enable_lazy_logits(teacher)
    probe = student(input_ids=ids, attention_mask=torch.ones_like(ids), labels=labels)
    assert isinstance(probe.logits, K.LazyLogits) and probe.loss is None and probe.logits.shape == (B, T, V)
    out_lazy = _compute_loss(student, teacher, loss_fn, inputs, top_k)
    out_lazy[0].backward()

    a = [float(x.detach()) for x in out_lazy]
    b = [float(x.detach()) for x in out_stock]
    np.testing.assert_allclose(a, b, rtol=2e-2, atol=1e-3)  # bf16 scalars (reference dtypes); bf16 vs fp32 logits
    gh = student.lm_head.weight.grad.float()
    assert float((gh - g_head.float()).abs().max() / g_head.float().abs().max()) < 3e-2
    ge = student.model.embed_tokens.weight.grad.float()
    assert float((ge - g_emb.float()).abs().max() / g_emb.float().abs().max()) < 5e-2  # dH flows into the body

    # and against the oracle on the stock student's logits (fp32 reference loss)
    student.forward = student._kd_original_forward
    teacher.forward = teacher._kd_original_forward
    with torch.no_grad():
        z = student(input_ids=ids).logits.float().cpu()
        y = teacher(input_ids=ids).logits[..., :V].float().cpu()
    if top_k > 0:
        tv, ti = O.topk_logprobs_reference(y.bfloat16().cuda(), top_k)  # bf16 log-softmax as the reference runs it
        ref = O.reference_loss(z, labels.cpu(), teacher_top_k_v=tv.cpu(), teacher_top_k_i=ti.cpu(),
                               speech_token_mask=mask)
    else:
        ref = O.reference_loss(z, labels.cpu(), teacher_logits=y, speech_token_mask=mask)
    np.testing.assert_allclose(a[:3], [float(x.detach()) for x in ref][:3], rtol=3e-2, atol=2e-3)


@pytest.mark.parametrize("with_num_items", [False, True])
def test_stage1_fused_ce_model_patch(with_num_items):
    """stage1.py:143 + :298-340 on a tiny Qwen3: freeze_model_weights + the trainer's ``model(**batch).loss``.
    Patched (fused CE, dW rows of the old vocabulary never computed) vs stock transformers ForCausalLMLoss + the
    reference hook: same loss, same gradients on the new-token rows, exact zeros on the old ones."""
    import speech_distill_b200 as K

    V, new, B, T = 1000, 40, 2, 48
    model = _tiny_qwen3(V, 64, 5)
    g = torch.Generator().manual_seed(8)
    ids = torch.randint(0, V, (B, T), generator=g).cuda()
    labels = ids.clone()
    labels[:, :7] = -100
    labels[1, 40:] = -100
    kw = {"num_items_in_batch": torch.tensor(61)} if with_num_items else {}

    K.freeze_model_weights(model, new)
    out_ref = model(input_ids=ids, attention_mask=torch.ones_like(ids), labels=labels, **kw)
    out_ref.loss.backward()
    g_head = model.lm_head.weight.grad.clone()
    g_emb = model.model.embed_tokens.weight.grad.clone()
    model.zero_grad(set_to_none=True)

    K.enable_fused_ce(model, num_new_tokens=new)
    out = model(input_ids=ids, attention_mask=torch.ones_like(ids), labels=labels, **kw)
    assert isinstance(out.logits, K.LazyLogits)
    out.loss.backward()
    assert abs(float(out.loss) - float(out_ref.loss)) <= 2e-3 * abs(float(out_ref.loss))
    gh = model.lm_head.weight.grad
    assert float(gh[: V - new].abs().max()) == 0.0 and float(g_head[: V - new].abs().max()) == 0.0  # stage1.py:53-57
    ref_new, got_new = g_head[V - new:].float(), gh[V - new:].float()
    assert float((got_new - ref_new).abs().max() / ref_new.abs().max()) < 2e-2
    ge, ge_ref = model.model.embed_tokens.weight.grad.float(), g_emb.float()
    assert float(ge[: V - new].abs().max()) == 0.0
    assert float((ge[V - new:] - ge_ref[V - new:]).abs().max() / ge_ref[V - new:].abs().max().clamp_min(1e-12)) < 5e-2


# ---- replay of the UNMODIFIED reference compute_loss (oracle/make_golden_flow.py, fp32 CPU run) ---------------------
import ast  # noqa: E402
import os  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load_model(d, tag):
    from transformers import Qwen3Config, Qwen3ForCausalLM

    cfg = ast.literal_eval(str(d[f"{tag}_cfg"]))
    model = Qwen3ForCausalLM(Qwen3Config(**cfg)).float()
    sd = {k[len(tag) + 1:]: torch.from_numpy(d[k]).view(torch.bfloat16).float() for k in d.files if k.startswith(tag + "/")}
    model.load_state_dict(sd)
    if tag == "student" and "bf16_head_input" in d.files and int(d["bf16_head_input"]):
        # same test double as oracle/make_golden_flow.py::bf16_head_input: bf16-representable head input, straight-through
        model.model.norm.register_forward_hook(lambda mod, args, out: out + (out.bfloat16().float() - out).detach())
    return model.cuda()  # fp32 body as in the reference run; the fused head casts its operands to bf16


@pytest.mark.parametrize("name,cached", [("onthefly_topk16", False), ("dense_teacher", False), ("onthefly_topk16", True)])
def test_flow_matches_reference_run(name, cached):
    """Fixtures = outputs of reference train.py:43-116 itself (tiny fp32 Qwen3 pair).  The same batch through
    models patched by enable_lazy_logits gives the reference's loss, its logged components and its gradients on
    the student's LM head and embedding, to the bf16 rounding of the head's operands."""
    import speech_distill_b200 as K

    d = np.load(os.path.join(GOLDEN, f"flow_{name}.npz"))
    student, teacher = _load_model(d, "student"), _load_model(d, "teacher")
    K.enable_lazy_logits(student)
    K.enable_lazy_logits(teacher)
    top_k = int(d["top_k"])
    ids = torch.from_numpy(d["input_ids"]).cuda()
    inputs = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": torch.from_numpy(d["labels"]).cuda(),
              "speech_token_mask": torch.from_numpy(d["speech_token_mask"]).cuda()}
    if cached:  # the pre-computed cache of extract_teacher_logits.py, produced by our own head + compaction kernels
        with torch.no_grad():
            hid = teacher(input_ids=ids).logits.hidden
        v, i = K.teacher_head_topk(hid, teacher.lm_head.weight, top_k, vocab_size=student.lm_head.weight.size(0))
        inputs["teacher_top_k_v"], inputs["teacher_top_k_i"] = v, i
        teacher = None
    loss_fn = K.DistillationLoss(temperature=2.0, alpha=0.5)
    total, task, distill, teach = _compute_loss_parts(student, teacher, loss_fn, inputs, top_k)
    total.backward()
    assert abs(float(total) - float(d["loss"])) <= 3e-3 * float(d["loss"])
    assert abs(float(task) - float(d["student_loss"])) <= 3e-3 * float(d["student_loss"])
    assert abs(float(distill) - float(d["distill_loss"])) <= 5e-3 * float(d["distill_loss"])
    assert abs(float(teach) - float(d["teacher_loss"])) <= 1e-2 * float(d["teacher_loss"])
    for got, want in ((student.lm_head.weight.grad, d["grad_lm_head"]), (student.model.embed_tokens.weight.grad, d["grad_embed"])):
        want = torch.from_numpy(want).cuda()
        assert got.dtype == torch.float32  # fp32 parameters receive fp32 gradients through the bf16 head
        err = float((got - want).abs().max() / want.abs().max())
        cos = float(torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0))
        print(f"{name} cached={cached}: grad err {err:.3e} cos {cos:.6f}")
        # the reference ran its teacher in fp32; here both heads run in bf16 (as train.py does on a GPU), so the
        # top-k log-probs carry bf16 rounding (2^-9 relative on |log p| ~ 5) and p_k moves by a few per cent
        assert err < (0.12 if top_k > 0 else 0.015) and cos > (0.999 if top_k > 0 else 0.9999)


def _compute_loss_parts(student, teacher, loss_fn, inputs, top_k):
    """_compute_loss above, returning the 4-tuple (train.py logs elements 1-3, :106-112)"""
    out = {}

    def spy(**kw):
        out["r"] = loss_fn(**kw)
        return out["r"]

    _compute_loss(student, teacher, spy, inputs, top_k)
    return out["r"]


def test_flow_bf16_heads_matches_reference_run_1e3():
    """The reference's compute_loss with bf16 heads on both sides (fixture flow_bf16head_cached_topk16: the top-k cache
    was written from a bf16 teacher head as extract_teacher_logits.py does on a GPU, and the student's head input is
    bf16-representable): nothing is left for the tensor-core head to round differently, so the replay meets north_star's
    1e-3 on the losses AND on the gradients of the fp32 LM head / embedding (they come from the fp32 accumulators)."""
    import speech_distill_b200 as K

    d = np.load(os.path.join(GOLDEN, "flow_bf16head_cached_topk16.npz"))
    student, teacher = _load_model(d, "student"), _load_model(d, "teacher")
    K.enable_lazy_logits(student)
    K.enable_lazy_logits(teacher)
    top_k = int(d["top_k"])
    ids = torch.from_numpy(d["input_ids"]).cuda()
    inputs = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": torch.from_numpy(d["labels"]).cuda(),
              "speech_token_mask": torch.from_numpy(d["speech_token_mask"]).cuda(),
              "teacher_top_k_v": torch.from_numpy(d["teacher_top_k_v"]).cuda(),
              "teacher_top_k_i": torch.from_numpy(d["teacher_top_k_i"]).cuda()}
    loss_fn = K.DistillationLoss(temperature=2.0, alpha=0.5)
    total, task, distill, teach = _compute_loss_parts(student, None, loss_fn, dict(inputs), top_k)
    total.backward()
    for got, key in ((total, "loss"), (task, "student_loss"), (distill, "distill_loss"), (teach, "teacher_loss")):
        assert abs(float(got) - float(d[key])) <= 1e-3 * abs(float(d[key])), (key, float(got), float(d[key]))
    for got, want in ((student.lm_head.weight.grad, d["grad_lm_head"]), (student.model.embed_tokens.weight.grad, d["grad_embed"])):
        want = torch.from_numpy(want).cuda()
        assert got.dtype == torch.float32
        err = float((got - want).abs().max() / want.abs().max())
        cos = float(torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0))
        print(f"bf16 heads: grad err {err:.3e} cos {cos:.7f}")
        assert err < 1e-3 and cos > 0.999999
    # the cache itself: our bf16 teacher head + compaction kernels on the replayed teacher reproduce the fixture's
    # (torch, CPU) cache - same indices wherever the bf16 log-probs are tie-free, values to one bf16 ulp
    with torch.no_grad():
        hid = teacher(input_ids=ids).logits.hidden
    v, i = K.teacher_head_topk(hid, teacher.lm_head.weight, top_k, vocab_size=student.lm_head.weight.size(0))
    v_ref = inputs["teacher_top_k_v"].float()
    assert float((v.float() - v_ref).abs().max()) <= 2.0 ** -7 * float(v_ref.abs().max())
    same = (torch.sort(i.long(), -1).values == torch.sort(inputs["teacher_top_k_i"].long(), -1).values).float().mean()
    assert float(same) > 0.85  # bf16 logits tie at the k-th place; a flipped hidden-state rounding swaps boundary entries
