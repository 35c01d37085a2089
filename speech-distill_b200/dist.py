"""Token-shard data parallelism for the KD hot path.

Every rank holds the full LM-head weight and its own sequences.  The only exchanges are
  (1) all-reduce(SUM) of the valid-row count  -> global N used inside the gradient kernels,
  (2) all-reduce(SUM) of the 8-float sums record -> identical loss scalars on every rank,
  (3) all-reduce(SUM) of dW (done by the caller / DDP; helper below).
With the global N the result equals the single-process reference evaluated on the concatenated
batch (HF-DDP would average per-rank means instead, which differs when valid counts differ).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def make_reduce_fns(group=None):
    """Returns (reduce_fn, count_reduce_fn) for kd_loss_on_logits / fused_linear_kd_loss."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None, None

    def reduce_fn(sums):
        out = sums.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    def count_reduce_fn(n_valid):
        out = n_valid.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    return reduce_fn, count_reduce_fn


def losses_from_sums(sums, tau, alpha, sparse=False):
    """Host/torch mirror of kd_finalize_losses (distillation_loss.py:68,116-118,123,126) for logging
    and for the CPU tests of the reduction logic; works on any device."""
    n = sums[3]
    if float(n) <= 0:
        z = torch.zeros((), dtype=sums.dtype)
        return z, z.clone(), z.clone(), z.clone()
    task = sums[0] / n
    distill = (tau * tau) * sums[1] / n
    if sparse:
        teacher = -sums[2] / sums[4] if float(sums[4]) > 0 else torch.zeros((), dtype=sums.dtype)
    else:
        teacher = sums[2] / n
    return alpha * task + (1 - alpha) * distill, task, distill, teacher


def allreduce_grad_(grad, group=None, bucket_rows=0):
    """In-place SUM all-reduce of a [V, H] LM-head gradient; ``bucket_rows`` > 0 splits it into row
    blocks so that NCCL can start on finished vocabulary chunks while later ones are computed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grad
    if bucket_rows <= 0:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
        return grad
    handles = []
    for r0 in range(0, grad.size(0), bucket_rows):
        handles.append(dist.all_reduce(grad[r0 : r0 + bucket_rows], op=dist.ReduceOp.SUM, group=group, async_op=True))
    for h in handles:
        h.wait()
    return grad


DEFAULT_V_CHUNK = 18944  # kDefaultVChunk of csrc/kd_fused.cu


def plan_ranges(V, row_begin, v_chunk, n_ranges):
    """Vocabulary ranges [v0, v1) for the overlapped dW all-reduce: whole backward chunks per range (so the
    chunking inside the library is unchanged), at most ``n_ranges`` of them, the frozen rows below
    ``row_begin`` (stage1) all in the first range.  Pure host logic (CPU-tested)."""
    vc = int(v_chunk) if v_chunk and v_chunk > 0 else DEFAULT_V_CHUNK
    vc = -(-vc // 256) * 256
    n_chunks = -(-V // vc)
    first_live = min(max(int(row_begin), 0) // vc, n_chunks - 1)  # chunk holding the first row with a gradient
    live = n_chunks - first_live
    n = max(1, min(int(n_ranges), live))
    base, extra = divmod(live, n)
    bounds, c = [0], first_live
    for r in range(n):
        c += base + (1 if r < extra else 0)
        bounds.append(min(c * vc, V))
    bounds[-1] = V
    return [(bounds[i], bounds[i + 1]) for i in range(n)]


class GradSync:
    """SUM all-reduce of the LM-head gradient, row block by row block, overlapped with the K1 backward.

    ``reduce_rows(dW, r0, r1)`` is called by the backward after each vocabulary range: the rows are final, the
    all-reduce is enqueued asynchronously (NCCL's stream waits for the current stream at that point) and the
    next range's GEMMs run beside it.  ``finish()`` makes the current stream wait for every all-reduce.
    ``max_ctas`` bounds the SMs NCCL may take (process-group config where torch exposes it).  The GEMM kernels
    draw their work units from an atomic counter, so the CTAs that have to wait for an SM NCCL occupies simply
    draw fewer units; ``reserve_sms=True`` instead launches the overlapped ranges on ``sm_count - max_ctas``
    SMs (kd_fused_linear_bwd_range sm_limit; useful with KD_SCHED=static).  Measured on 2 B200 (tools/n2_sweep.sh):
    32 CTAs without a reservation is the fastest setting.
    """

    def __init__(self, group=None, n_ranges=6, max_ctas=32, sm_count=None, reserve_sms=False, backend="nccl",
                 multimem_ctas=16):
        self.group = group
        self.n_ranges = int(n_ranges)
        self.max_ctas = int(max_ctas)
        self.reserve_sms = bool(reserve_sms)
        self._sm_count = sm_count
        self._pending = []
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        # backend "multimem": the gradient lives in a symmetric buffer and every range is reduced in the NVSwitch by
        # kd_multimem_allreduce (csrc/kd_multimem.cu) - 2/G of the bytes through this GPU's SMs instead of a ring's
        # 2 (G-1)/G, a few CTAs for ~0.1 ms per range.  Needs NVLS multicast (torch symmetric memory reports it); falls
        # back to NCCL otherwise.
        self.backend = backend if self.active else "nccl"
        self.multimem_ctas = int(multimem_ctas)
        self._symm = None       # (buffer tensor [V, H], handle, key)
        self._side = None
        self._mm_ranges = 0
        self._out = None        # this backward's result tensor (the reduced rows are copied out range by range)

    def _new_out(self, V, H, dtype, device, zero_rows):
        self._out = (torch.zeros if zero_rows > 0 else torch.empty)((int(V), int(H)), dtype=dtype, device=device)
        self._out.record_stream(self._side)  # written on the side stream, range by range

    def result(self):
        """The reduced gradient of the backward that just finished (multimem backend): an ordinary tensor of its own -
        the symmetric buffer is overwritten by the next backward."""
        out, self._out = self._out, None
        return out

    def grad_buffer(self, V, H, dtype, device, zero_rows=0):
        """The [V, H] tensor the backward should write dW into: a symmetric (multicast-mapped) buffer for the multimem
        backend, None for NCCL (the backward allocates as usual).  Allocated and exchanged once per shape."""
        if self.backend != "multimem":
            return None
        key = (int(V), int(H), dtype, torch.device(device))
        if self._symm is not None and self._symm[2] == key:
            self._new_out(V, H, dtype, device, zero_rows)
            return self._symm[0]
        try:
            import torch.distributed._symmetric_memory as symm_mem

            buf = symm_mem.empty((int(V), int(H)), dtype=dtype, device=device)
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)  # every rank of the job holds a replica
            if not int(getattr(hdl, "multicast_ptr", 0)):
                raise RuntimeError("no NVLS multicast support on this system")
        except Exception as e:  # noqa: BLE001 - any failure means: use NCCL
            import warnings

            warnings.warn(f"GradSync: multimem backend unavailable ({type(e).__name__}: {e}); using NCCL")
            self.backend = "nccl"
            return None
        self._symm = (buf, hdl, key)
        if self._side is None:
            self._side = torch.cuda.Stream(device=device)
        self._new_out(V, H, dtype, device, zero_rows)
        return buf

    @staticmethod
    def new_group(max_ctas=32, **kw):
        """A NCCL process group whose collectives use at most ``max_ctas`` CTAs (falls back to the default
        group configuration when this torch build does not expose the option)."""
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = int(max_ctas)
            opts.config.min_ctas = min(int(max_ctas), 4)
            return dist.new_group(backend="nccl", pg_options=opts, **kw)
        except Exception:  # pragma: no cover - depends on the torch / NCCL build
            return dist.new_group(**kw)

    def ranges(self, V, row_begin, v_chunk):
        if not self.active:
            return [(0, V)]
        return plan_ranges(V, row_begin, v_chunk, self.n_ranges)

    def sm_limit(self):
        if not self.active or self.max_ctas <= 0 or not self.reserve_sms:
            return 0
        if self._sm_count is None:
            self._sm_count = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        lim = self._sm_count - self.max_ctas
        return lim - (lim % 2) if lim >= 2 else 0  # CTA pairs

    def ready_stream_ptr(self, device):
        """Raw handle of the stream the finished rows of a range are handed to (kd_fused_linear_bwd_range's
        dw_ready_stream): the all-reduce of every range but the last is enqueued there, so the caller's stream - from
        which the next range's kernels fork - never waits for a range's dW chain.  0 when inactive."""
        if not self.active:
            return 0
        if self._side is None:
            self._side = torch.cuda.Stream(device=device)
        return self._side.cuda_stream

    def reduce_rows(self, grad, r0, r1, last=True):
        """``last=False``: the rows' completion was handed to the ready stream (see ready_stream_ptr), the collective is
        enqueued behind it there; ``last=True`` (or no ready stream): behind the current stream."""
        if not (self.active and r1 > r0):
            return
        on_side = (not last) and self._side is not None
        if self.backend == "multimem" and self._symm is not None and grad.data_ptr() == self._symm[0].data_ptr():
            from . import _lib

            buf, hdl, _ = self._symm
            row_bytes = buf.stride(0) * buf.element_size()
            if not on_side:  # the range's dW rows are final on the current stream
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(grad.device))
            with torch.cuda.stream(self._side):
                if not on_side:
                    self._side.wait_event(ready)
                hdl.barrier(channel=0)  # every rank has written its rows of the range
                _lib.check(_lib.load().kd_multimem_allreduce(int(hdl.multicast_ptr), int(r0) * row_bytes,
                                                             int(r1 - r0) * row_bytes, _lib.dtype_code(buf.dtype),
                                                             int(hdl.rank), int(hdl.world_size), self.multimem_ctas,
                                                             self._side.cuda_stream), "kd_multimem_allreduce")
                hdl.barrier(channel=1)  # every rank's sums have landed in every replica
                self._out[r0:r1].copy_(buf[r0:r1], non_blocking=True)
            self._mm_ranges += 1
            return
        if on_side:  # NCCL's stream waits for the stream that is current at the call: make it the ready stream
            with torch.cuda.stream(self._side):
                self._pending.append(dist.all_reduce(grad[r0:r1], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            return
        self._pending.append(dist.all_reduce(grad[r0:r1], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        for w in self._pending:
            w.wait()
        self._pending = []
        if self._mm_ranges:
            torch.cuda.current_stream(self._symm[0].device).wait_stream(self._side)
            self._mm_ranges = 0
