"""Collator fast path for the top-k teacher cache against a restatement of the reference's
``_pad_logits`` (data.py:330-348)."""
import numpy as np
import torch


def _reference_pad_logits(logit_list, max_length, padding_value=0.0):
    # data.py:330-348, restated
    batch = []
    for l in logit_list:
        if not isinstance(l, torch.Tensor):
            l = torch.tensor(l)
        p_len = max_length - l.size(0)
        if p_len > 0:
            l = torch.cat([l, torch.full((p_len, l.size(1)), padding_value, dtype=l.dtype)], dim=0)
        elif p_len < 0:
            l = l[:max_length]
        batch.append(l)
    return torch.stack(batch)


def test_pad_logits_matches_reference_contract():
    from speech_distill_b200.cache import WIRE_DTYPES, collate_teacher_topk, pad_logits

    g = np.random.default_rng(0)
    lens, K, T = [5, 9, 12, 1], 8, 9
    v = [(-np.abs(g.standard_normal((n, K))) * 3).astype(np.float16) for n in lens]   # wire format of the extractor
    i = [g.integers(0, 152936, (n, K)).astype(np.int32) for n in lens]
    # reference path as data.py feeds it: datasets decodes the columns to nested Python lists
    ref_v = _reference_pad_logits([x.tolist() for x in v], T, 0.0)
    ref_i = _reference_pad_logits([x.tolist() for x in i], T, 0)
    assert ref_v.dtype == torch.float32 and ref_i.dtype == torch.int64             # what the reference ends up with
    got_v = pad_logits(v, T, 0.0, WIRE_DTYPES["teacher_top_k_v"])
    got_i = pad_logits(i, T, 0, WIRE_DTYPES["teacher_top_k_i"])
    assert got_v.dtype == torch.float16 and got_i.dtype == torch.int32 and got_v.shape == (4, T, K)
    assert torch.equal(got_v.float(), ref_v) and torch.equal(got_i.long(), ref_i)    # fp16 -> fp32 is exact
    # nested lists and tensors are accepted too; default dtype follows the input
    assert torch.equal(pad_logits([x.tolist() for x in v], T), ref_v)
    assert pad_logits([torch.from_numpy(x) for x in i], T, 0).dtype == torch.int32
    feats = [{"teacher_top_k_v": a, "teacher_top_k_i": b} for a, b in zip(v, i)]
    out = collate_teacher_topk(feats, T)
    assert torch.equal(out["teacher_top_k_v"], got_v) and torch.equal(out["teacher_top_k_i"], got_i)
    assert collate_teacher_topk([{"input_ids": [1, 2]}], T) == {}
