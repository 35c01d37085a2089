"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    python oracle/make_golden.py            # needs /root/reference (not present on the GPU box)

The reference holds no fixtures of its own, so these vectors - outputs of
``/root/reference/distillation_loss.py::DistillationLoss``, of the three torch calls at
``extract_teacher_logits.py:114-129`` and of the hooks that ``stage1.py::freeze_model_weights``
installs - are the pins for ``oracle/kd_oracle.py`` and, through it, for the CUDA kernels.
Nothing under tests/, bench.py or the package reads /root/reference at run time.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("KD_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    sys.path.insert(0, REF)
    import distillation_loss as ref_loss  # noqa: E402  (imports torch only)

    assert os.path.abspath(ref_loss.__file__).startswith(os.path.abspath(REF))
    return ref_loss


def _import_stage1_freeze():
    """stage1.py needs trl / s3tokenizer / datasets at import time; stub what is missing."""
    from transformers import Trainer  # noqa: F401  (must precede the peft stub)

    for name in ("peft", "trl", "s3tokenizer", "torchaudio", "librosa", "onnxruntime"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__getattr__ = lambda attr, _n=name: type(attr, (), {})
                sys.modules[name] = m
    import stage1 as ref_stage1

    return ref_stage1.freeze_model_weights


def make_inputs(seed, B, T, V, K=0, dtype=torch.float64, mask_kind="mixed", fill=False):
    g = torch.Generator().manual_seed(seed)
    z = (torch.randn(B, T, V, generator=g, dtype=torch.float64) * 2).to(dtype)
    y = (torch.randn(B, T, V, generator=g, dtype=torch.float64) * 2).to(dtype)
    if fill:  # cosyvoice2/teacher_wrapper.py:139-162 fills unused head columns with -10000.0
        y[..., V // 2 : V // 2 + V // 5] = -10000.0
    labels = torch.randint(0, V, (B, T), generator=g)
    speech = None
    if mask_kind == "mixed":
        labels[:, : max(1, T // 4)] = -100
        labels[0, -2:] = -100
        speech = torch.ones(B, T, dtype=torch.float32)
        speech[B - 1, T // 2] = 0.0
    elif mask_kind == "empty":
        labels[:] = -100
    out = dict(z=z, y=y, labels=labels, speech=speech)
    if K:
        lp = F.log_softmax(y.float(), dim=-1)
        v, i = torch.topk(lp, K, dim=-1)
        out["v"] = v.to(torch.float16)
        out["i"] = i.to(torch.int32)
        # make about half of the scored labels land inside the teacher's top-k (monitor hits)
        for b in range(B):
            for t in range(1, T):
                if labels[b, t] != -100 and (b + t) % 2 == 0:
                    labels[b, t] = int(i[b, t - 1, (b + 3 * t) % K])
    return out


def run_loss(ref_loss, inp, mode, tau, alpha):
    fn = ref_loss.DistillationLoss(temperature=tau, alpha=alpha)
    z = inp["z"].clone().requires_grad_(True)
    kw = dict(speech_token_mask=inp["speech"])
    if mode == "dense":
        kw["teacher_logits"] = inp["y"]
    else:
        kw["teacher_top_k_v"] = inp["v"]
        kw["teacher_top_k_i"] = inp["i"]
    out = fn(z, inp["labels"], **kw)
    if out[0].requires_grad:
        out[0].backward()
        grad = z.grad
    else:
        grad = torch.zeros_like(z)
    return [float(o) for o in out], grad


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_loss = _import_reference()
    torch.manual_seed(0)
    cases = [
        # name, seed, B, T, V, K, dtype, mask, fill, mode, tau, alpha
        ("dense_f64_mixed", 11, 2, 9, 257, 0, torch.float64, "mixed", False, "dense", 2.0, 0.5),
        ("dense_f64_fill", 12, 2, 11, 517, 0, torch.float64, "mixed", True, "dense", 2.0, 0.5),
        ("dense_f64_tau3", 13, 3, 7, 130, 0, torch.float64, "all", False, "dense", 3.0, 0.3),
        ("dense_f32_tau1", 14, 2, 8, 1000, 0, torch.float32, "mixed", False, "dense", 1.0, 0.7),
        ("dense_empty", 15, 2, 5, 64, 0, torch.float32, "empty", False, "dense", 2.0, 0.5),
        ("sparse_f64_k8", 21, 2, 9, 257, 8, torch.float64, "mixed", False, "sparse", 2.0, 0.5),
        ("sparse_f32_k64", 22, 2, 6, 2051, 64, torch.float32, "all", False, "sparse", 2.0, 0.5),
        ("sparse_f64_tau15", 23, 1, 12, 300, 16, torch.float64, "mixed", False, "sparse", 1.5, 0.25),
    ]
    for name, seed, B, T, V, K, dt, mk, fill, mode, tau, alpha in cases:
        inp = make_inputs(seed, B, T, V, K, dt, mk, fill)
        losses, grad = run_loss(ref_loss, inp, mode, tau, alpha)
        rec = dict(
            z=inp["z"].numpy(),
            labels=inp["labels"].numpy(),
            losses=np.array(losses, dtype=np.float64),
            grad=grad.numpy(),
            tau=tau,
            alpha=alpha,
            mode=mode,
        )
        if inp["speech"] is not None:
            rec["speech"] = inp["speech"].numpy()
        if mode == "dense":
            rec["y"] = inp["y"].numpy()
        else:
            rec["v"] = inp["v"].numpy()
            rec["i"] = inp["i"].numpy()
        np.savez_compressed(os.path.join(OUT, f"loss_{name}.npz"), **rec)
        print(name, losses)

    # sparse case where no label is inside the top-k -> teacher monitor must be exactly 0.0
    inp = make_inputs(31, 1, 6, 200, 4, torch.float64, "all", False)
    for b in range(1):
        for t in range(1, 6):
            taken = set(inp["i"][b, t - 1].tolist())
            lab = next(c for c in range(200) if c not in taken)
            inp["labels"][b, t] = lab
    losses, grad = run_loss(ref_loss, inp, "sparse", 2.0, 0.5)
    assert losses[3] == 0.0
    np.savez_compressed(
        os.path.join(OUT, "loss_sparse_nohit.npz"),
        z=inp["z"].numpy(), labels=inp["labels"].numpy(), v=inp["v"].numpy(), i=inp["i"].numpy(),
        losses=np.array(losses), grad=grad.numpy(), tau=2.0, alpha=0.5, mode="sparse",
    )
    print("sparse_nohit", losses)

    # bf16-stored inputs, reference evaluated in fp32 on the rounded values (parity protocol)
    inp = make_inputs(41, 2, 8, 1031, 16, torch.float32, "mixed", False)
    inp["z"] = inp["z"].bfloat16().float()
    inp["y"] = inp["y"].bfloat16().float()
    for mode in ("dense", "sparse"):
        losses, grad = run_loss(ref_loss, inp, mode, 2.0, 0.5)
        rec = dict(z=inp["z"].numpy(), labels=inp["labels"].numpy(), speech=inp["speech"].numpy(),
                   losses=np.array(losses), grad=grad.numpy(), tau=2.0, alpha=0.5, mode=mode)
        if mode == "dense":
            rec["y"] = inp["y"].numpy()
        else:
            rec["v"] = inp["v"].numpy()
            rec["i"] = inp["i"].numpy()
        np.savez_compressed(os.path.join(OUT, f"loss_{mode}_bf16vals.npz"), **rec)
        print(mode + "_bf16vals", losses)

    # fused LM-head form: logits = hidden @ W^T (HF nn.Linear), grads w.r.t. hidden and W
    g = torch.Generator().manual_seed(51)
    B, T, H, V = 2, 10, 32, 389
    h = (torch.randn(B, T, H, generator=g, dtype=torch.float64)).bfloat16().double().requires_grad_(True)
    W = (torch.randn(V, H, generator=g, dtype=torch.float64) * 0.3).bfloat16().double().requires_grad_(True)
    y = (torch.randn(B, T, V, generator=g, dtype=torch.float64) * 2).bfloat16().double()
    labels = torch.randint(0, V, (B, T), generator=g)
    labels[:, :2] = -100
    fn = ref_loss.DistillationLoss(temperature=2.0, alpha=0.5)
    out = fn(F.linear(h, W), labels, teacher_logits=y)
    out[0].backward()
    np.savez_compressed(
        os.path.join(OUT, "fused_dense_f64.npz"),
        h=h.detach().numpy(), W=W.detach().numpy(), y=y.numpy(), labels=labels.numpy(),
        losses=np.array([float(o) for o in out]), dh=h.grad.numpy(), dW=W.grad.numpy(),
        tau=2.0, alpha=0.5,
    )
    print("fused_dense", [float(o) for o in out])

    # top-k extraction, tie-free fp32 logits (extract_teacher_logits.py:114-129)
    g = torch.Generator().manual_seed(61)
    logits = torch.randn(3, 5, 4099, generator=g) * 3
    lp = F.log_softmax(logits, dim=-1)
    tv, ti = torch.topk(lp, k=64, dim=-1)
    np.savez_compressed(
        os.path.join(OUT, "topk_f32.npz"), logits=logits.numpy(),
        v=tv.to(torch.float16).numpy(), i=ti.to(torch.int32).numpy(), k=64,
    )

    # stage1 hooks (stage1.py:29-73) on a tied-embedding toy LM
    try:
        freeze = _import_stage1_freeze()

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

        torch.manual_seed(7)
        V, H, new = 50, 16, 6
        toy = Toy(V, H).double()
        import contextlib, io

        with contextlib.redirect_stdout(io.StringIO()):
            freeze(toy, new)
        ids = torch.randint(0, V, (2, 9))
        logits = toy(ids)
        loss = F.cross_entropy(logits[:, :-1].reshape(-1, V), ids[:, 1:].reshape(-1))
        loss.backward()
        trainable = [n for n, p in toy.named_parameters() if p.requires_grad]
        # unmasked gradient for comparison
        toy2 = Toy(V, H).double()
        toy2.load_state_dict(toy.state_dict())
        l2 = F.cross_entropy(toy2(ids)[:, :-1].reshape(-1, V), ids[:, 1:].reshape(-1))
        l2.backward()
        np.savez_compressed(
            os.path.join(OUT, "stage1_mask.npz"),
            masked=toy.emb.weight.grad.numpy(), unmasked=toy2.emb.weight.grad.numpy(),
            old_vocab=V - new, n_trainable=len(trainable),
        )
        print("stage1 mask ok; trainable:", trainable)
    except Exception as e:  # pragma: no cover - environment dependent
        print("stage1 fixture skipped:", repr(e))


if __name__ == "__main__":
    main()
