"""K1 / K2 / K3 at BASELINE.json's full sizes.  The oracle (reference loss restated in torch) is run on the GPU in
fp32 as the checker - it needs seconds there - plus size-independent properties: linearity in the upstream
gradient (bit-exact for a power of two), row independence of dH, sums over vocabulary slices."""
import numpy as np
import pytest
import torch

from oracle import kd_oracle as O

pytestmark = pytest.mark.gpu
V_FULL, H_STUDENT = 152936, 1024


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _inputs(B, T, seed, masked=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    h = torch.randn(B, T, H_STUDENT, device="cuda", generator=g).bfloat16()
    W = (torch.randn(V_FULL, H_STUDENT, device="cuda", generator=g) * (2.0 / H_STUDENT ** 0.5)).bfloat16()
    y = torch.empty(B, T, V_FULL, device="cuda", dtype=torch.bfloat16)
    for b in range(B):
        y[b] = (torch.randn(T, V_FULL, device="cuda", generator=g) * 2).bfloat16()
    labels = torch.randint(0, V_FULL, (B, T), device="cuda", generator=g)
    if masked:  # text prefix + padded tail, the collator's pattern (data.py:246-251)
        labels[:, : T // 4] = -100
        labels[0, -T // 10:] = -100
    return h, W, y, labels


def test_k1_dense_configs1_full_size_vs_oracle():
    """BASELINE configs[1]: B=8, T=512, H=1024, V=152,936, dense bf16 teacher, tau=2, alpha=0.5."""
    import speech_distill_b200 as K

    h, W, y, labels = _inputs(8, 512, 2024)
    ref, gh_ref, gw_ref = O.fused_linear_reference(h.float(), W.float(), labels, teacher_logits=y.float(),
                                                   temperature=2.0, alpha=0.5)
    hc, Wc = h.clone().requires_grad_(True), W.clone().requires_grad_(True)
    out = K.fused_linear_kd_loss(hc, Wc, labels, teacher_logits=y, temperature=2.0, alpha=0.5)
    out[0].backward()
    for got, want in zip(out, ref):
        got, want = float(got.detach()), float(want.detach())
        assert abs(got - want) <= 1e-3 * max(1.0, abs(want))  # north_star tolerance
        assert abs(got - want) <= 2e-5 * max(1.0, abs(want))  # what fp32 statistics deliver
    _, gh32, gw32 = K.fused_linear_kd_value_and_grad(h, W, labels, teacher_logits=y, temperature=2.0, alpha=0.5)
    assert rel_err(gh32, gh_ref) < 1e-3 and rel_err(gw32, gw_ref) < 1e-3    # north_star tolerance (fp32 accumulate)
    # bf16 outputs: one rounding of the leaf gradients on top (half an ulp = 2e-3 of the largest entry)
    assert rel_err(hc.grad.float(), gh_ref) < 4e-3 and rel_err(Wc.grad.float(), gw_ref) < 4e-3
    cos = torch.nn.functional.cosine_similarity(Wc.grad.float().flatten(), gw_ref.flatten(), dim=0)
    assert float(cos) > 0.99999

    # linearity in the upstream gradient: d(4 * total) = 4 * d(total), bit for bit (power of two)
    h4, W4 = h.clone().requires_grad_(True), W.clone().requires_grad_(True)
    out4 = K.fused_linear_kd_loss(h4, W4, labels, teacher_logits=y, temperature=2.0, alpha=0.5)
    (out4[0] * 4.0).backward()
    assert torch.equal(h4.grad, hc.grad * 4) and torch.equal(W4.grad, Wc.grad * 4)

    # row independence: dH of the first 3 sequences computed alone (other tile schedule, other N) = same rows
    # up to the 1/N factor; compare after rescaling by the valid-row counts.  The gradient tile is rounded (to
    # scaled fp16) after the 1/N scaling, so the two runs round different numbers: equal to that rounding level
    n_all = int((labels[:, 1:] != -100).sum())
    n_sub = int((labels[:3, 1:] != -100).sum())
    _, gh_sub, _ = K.fused_linear_kd_value_and_grad(h[:3].contiguous(), W, labels[:3].contiguous(),
                                                    teacher_logits=y[:3].contiguous(), temperature=2.0, alpha=0.5)
    assert rel_err(gh_sub * (n_sub / n_all), gh32[:3]) < 1e-3


def test_k1_sparse_configs2_full_size_vs_oracle():
    """BASELINE configs[2] (loss part): B=16, T=512, top-k=64 cache from a teacher, sparse KD, V=152,936."""
    import speech_distill_b200 as K

    h, W, y, labels = _inputs(16, 512, 77)
    tv, ti = K.teacher_topk_logprobs(y, 64)
    del y
    ref, gh_ref, gw_ref = O.fused_linear_reference(h.float(), W.float(), labels, teacher_top_k_v=tv,
                                                   teacher_top_k_i=ti, temperature=2.0, alpha=0.5)
    losses, gh32, gw32 = K.fused_linear_kd_value_and_grad(h, W, labels, temperature=2.0, alpha=0.5,
                                                          teacher_top_k_v=tv, teacher_top_k_i=ti)
    for got, want in zip(losses, ref):
        assert abs(float(got) - float(want)) <= 2e-5 * max(1.0, abs(float(want)))
    assert rel_err(gh32, gh_ref) < 1e-3 and rel_err(gw32, gw_ref) < 1e-3


def test_stage1_configs3_full_size_masked_rows():
    """BASELINE configs[3]: CE with the frozen-vocabulary mask, B=8, T=2048, 1,000 new rows: rows < V_old of dW are
    exactly zero and never computed; the live rows and dH match the oracle."""
    import speech_distill_b200 as K

    g = torch.Generator(device="cuda").manual_seed(5)
    B, T, old = 8, 2048, V_FULL - 1000
    h = torch.randn(B, T, H_STUDENT, device="cuda", generator=g).bfloat16()
    W = (torch.randn(V_FULL, H_STUDENT, device="cuda", generator=g) * (2.0 / H_STUDENT ** 0.5)).bfloat16()
    labels = torch.randint(old - 500, V_FULL, (B, T), device="cuda", generator=g)  # mostly speech tokens
    loss_ref, gh_ref, gw_ref = O.stage1_ce_reference(h.float(), W.float(), labels, old)
    hc, Wc = h.clone().requires_grad_(True), W.clone().requires_grad_(True)
    loss = K.fused_linear_cross_entropy(hc, Wc, labels, old_vocab_size=old)
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) <= 2e-5 * float(loss_ref)
    assert float(Wc.grad[:old].abs().max()) == 0.0
    assert rel_err(Wc.grad[old:].float(), gw_ref[old:]) < 4e-3
    assert rel_err(hc.grad.float(), gh_ref) < 4e-3


def test_teacher_head_topk_configs2_full_size():
    """BASELINE configs[2] (teacher part): SoulX-1.7B head, hidden 2048, B=16, T=512 -> top-64; the row-block
    pipeline (GEMM epilogue statistics + piece-wise selection) picks the same indices as the compaction of the
    materialised logits of the same GEMM, bit for bit."""
    import speech_distill_b200 as K

    g = torch.Generator(device="cuda").manual_seed(9)
    Ht, R = 2048, 16 * 512
    h = torch.randn(R, Ht, device="cuda", generator=g).bfloat16()
    W = (torch.randn(V_FULL, Ht, device="cuda", generator=g) * (2.5 / Ht ** 0.5)).bfloat16()
    v, i = K.teacher_head_topk(h, W, 64)
    logits = K.linear_bf16(h, W)
    v2, i2 = K.teacher_topk_logprobs(logits, 64)
    assert torch.equal(i, i2)  # indices: bit for bit
    diff = (v.float() - v2.float()).abs()  # values: the fp32 log-sum-exp is summed in another order (GEMM epilogue)
    assert bool((diff <= 2.0 ** -7 * v2.float().abs()).all()) and float((diff > 0).float().mean()) < 1e-3
    vu, iu = K.teacher_head_topk(h, W, 64, fused=False)
    assert torch.equal(iu, i2) and torch.equal(vu, v2)  # the unfused row-block pipeline: bit for bit
    sel = torch.gather(logits.float(), -1, i.long())
    assert torch.equal(sel, torch.topk(logits.float(), 64, -1).values)  # a valid top-k of those logits


class _FullSizeRangeStub:
    """dist.GradSync's interface without a process group (one GPU): the real range plan; `reduce_rows` does what an
    all-reduce over identical ranks would leave behind on a side stream that waits for the range - it reads and
    rewrites the finished row block while the next range's GEMMs run."""

    def __init__(self, n_ranges, ready=False):
        self.n_ranges, self.blocks = n_ranges, []
        self.side = torch.cuda.Stream()
        self.done = []
        self.ready = ready
        if ready:  # GradSync.ready_stream_ptr: finished rows are handed to the side stream, the caller's stream never
            self.ready_stream_ptr = lambda device: self.side.cuda_stream  # waits for a range's dW chain

    def ranges(self, V, row_begin, v_chunk):
        from speech_distill_b200.dist import plan_ranges

        return plan_ranges(V, row_begin, v_chunk, self.n_ranges)

    def sm_limit(self):
        return 0

    def reduce_rows(self, grad, r0, r1, last=True):
        self.blocks.append((r0, r1))
        if last or not self.ready:  # otherwise the library already made the side stream wait for the rows
            self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            grad[r0:r1].mul_(2.0)  # world_size = 2 ranks holding the same gradient: SUM = 2 x (exact in bf16)
            ev = torch.cuda.Event()
            ev.record()
        self.done.append(ev)

    def finish(self):
        for ev in self.done:
            torch.cuda.current_stream().wait_event(ev)
        self.done = []


@pytest.mark.parametrize("ready", [False, True])
def test_gradsync_ranges_configs1_full_size(ready):
    """The overlapped dW exchange at BASELINE configs[1] size: 6 vocabulary ranges handed to a stand-in all-reduce on
    a side stream while the following ranges run = the one-call backward (dH bit for bit, dW exactly 2 x); with and
    without the dw_ready_stream hand-over of kd_fused_linear_bwd_range."""
    import speech_distill_b200 as K

    h, W, y, labels = _inputs(8, 512, 77)

    def run(sync):
        hc, Wc = h.clone().requires_grad_(True), W.clone().requires_grad_(True)
        out = K.fused_linear_kd_loss(hc, Wc, labels, teacher_logits=y, temperature=2.0, alpha=0.5, grad_sync=sync)
        out[0].backward()
        torch.cuda.synchronize()
        return hc.grad, Wc.grad

    gh0, gw0 = run(None)
    stub = _FullSizeRangeStub(6, ready)
    gh1, gw1 = run(stub)
    assert len(stub.blocks) == 6 and stub.blocks[0][0] == 0 and stub.blocks[-1][1] == V_FULL
    assert all(a[1] == b[0] for a, b in zip(stub.blocks, stub.blocks[1:]))
    assert torch.equal(gh0, gh1)
    assert torch.equal(gw0 * 2, gw1)


@pytest.mark.parametrize("cache_mb", [0, 1024])
def test_k1_peak_activation_memory_independent_of_V(cache_mb):
    """north_star: peak activation memory of the fused step does not depend on V.  Measured with the caching
    allocator's high-water mark around forward + backward at V = 152,936 and 2 V = 305,872 (R = 4096): after
    subtracting what is resident before the step (inputs, W) and the gradients it returns (dW [V,H], dH [R,H]) the
    two peaks are equal - gradient chunk buffers, fp32 dH accumulator, fp16 operand copies, row records and a logit
    cache whose size is a caller-chosen constant (here 0 and 1 GB: six whole 18,944-column chunks at either V)."""
    import speech_distill_b200 as K

    g = torch.Generator(device="cuda").manual_seed(3)
    B, T = 8, 512
    peaks = {}
    for mult in (1, 2):
        V = V_FULL * mult
        h = torch.randn(B, T, H_STUDENT, device="cuda", generator=g).bfloat16().requires_grad_(True)
        W = (torch.randn(V, H_STUDENT, device="cuda", generator=g) * (2.0 / H_STUDENT ** 0.5)).bfloat16().requires_grad_(True)
        y = torch.empty(B, T, V, device="cuda", dtype=torch.bfloat16)
        for b in range(B):
            y[b] = (torch.randn(T, V, device="cuda", generator=g) * 2).bfloat16()
        labels = torch.randint(0, V, (B, T), device="cuda", generator=g)
        for _ in range(2):  # the first call also sizes the library's per-device pools
            h.grad = W.grad = None
            out = K.fused_linear_kd_loss(h, W, labels, teacher_logits=y, temperature=2.0, alpha=0.5,
                                         logit_cache_mb=cache_mb)
            out[0].backward()
            del out
        h.grad = W.grad = None
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        out = K.fused_linear_kd_loss(h, W, labels, teacher_logits=y, temperature=2.0, alpha=0.5,
                                     logit_cache_mb=cache_mb)
        out[0].backward()
        torch.cuda.synchronize()
        grads = W.grad.numel() * 2 + h.grad.numel() * 2
        peaks[mult] = torch.cuda.max_memory_allocated() - base - grads
        assert float(out[0]) > 0
        del h, W, y, labels, out
        torch.cuda.empty_cache()
    print(f"peak activation bytes (cache budget {cache_mb} MB): V {peaks[1] / 2**20:.1f} MiB, 2V {peaks[2] / 2**20:.1f} MiB")
    assert abs(peaks[2] - peaks[1]) <= 4 * 2 ** 20, peaks   # allocator rounding only (2 MiB blocks)
    assert peaks[1] < (cache_mb + 600) * 2 ** 20, peaks        # 2 x 155 MB chunks + 16 MB dH + records (+ cache)
