"""Host -> device input pipeline for the KD step (plumbing; no arithmetic).

The step's inputs live in pinned host memory (hidden states, teacher logits or top-k cache, labels: what a
DataLoader with ``pin_memory=True`` hands to the trainer).  ``HostPrefetcher`` copies batch i + 1 on its own
stream into the second of two device staging sets while batch i is being computed, so the PCIe transfer
(1.26 GB per step for a dense bf16 teacher at B=8, T=512, V=152,936) overlaps the kernels instead of preceding them.
Ordering is by CUDA events only; the host never blocks.
"""
from __future__ import annotations

import torch


class HostPrefetcher:
    """Double-buffered asynchronous staging of a tuple of pinned host tensors.

        pf = HostPrefetcher(device)
        pf.submit(batch0)                    # starts copying batch0
        for nxt in batches[1:] + [None]:
            cur = pf.next(nxt)               # device tensors of the batch submitted before; starts copying `nxt`
            step(*cur)                       # enqueue the compute that reads `cur` BEFORE the next pf.next()

    A staging set is overwritten only after the work that read it: ``next()`` marks the set it handed out on the
    previous call as free at that point of the current stream (everything enqueued so far has to finish first),
    so the loop above needs no explicit ``release()``; calling ``release()`` right after the step is enqueued frees
    the set a little earlier.  ``submit()`` refuses to overwrite a batch that was never taken.
    """

    _EMPTY, _FILLED, _HANDED = 0, 1, 2

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self._sets = [None] * depth          # device staging tensors
        self._ready = [torch.cuda.Event() for _ in range(depth)]   # copy finished
        self._free = [None] * depth          # compute that read the set finished
        self._state = [self._EMPTY] * depth
        self._n_submit = 0
        self._n_take = 0
        self.bytes_per_batch = 0

    def _alloc(self, slot, batch):
        cur = self._sets[slot]
        if cur is not None and all(c.shape == b.shape and c.dtype == b.dtype for c, b in zip(cur, batch)):
            return cur
        self._sets[slot] = tuple(torch.empty(b.shape, dtype=b.dtype, device=self.device) for b in batch)
        return self._sets[slot]

    def _mark_free(self, slot):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[slot] = ev
        self._state[slot] = self._EMPTY

    def submit(self, batch):
        """Enqueue the H2D copies of ``batch`` (tuple of pinned CPU tensors) on the copy stream."""
        for b in batch:
            if not b.is_pinned():
                raise ValueError("HostPrefetcher needs pinned host tensors (pin_memory=True)")
        slot = self._n_submit % self.depth
        if self._state[slot] == self._FILLED:
            raise RuntimeError(f"HostPrefetcher.submit(): all {self.depth} staging sets hold batches that were never "
                               "taken with next(); take one first")
        if self._state[slot] == self._HANDED:
            # handed out and not released: whatever reads it has been enqueued on the current stream by now
            self._mark_free(slot)
        dst = self._alloc(slot, batch)
        with torch.cuda.stream(self.stream):
            if self._free[slot] is not None:
                self.stream.wait_event(self._free[slot])  # the step that used this staging set is done
            for d, b in zip(dst, batch):
                d.copy_(b, non_blocking=True)
            self._ready[slot].record(self.stream)
        self._state[slot] = self._FILLED
        self.bytes_per_batch = sum(b.numel() * b.element_size() for b in batch)
        self._n_submit += 1

    def next(self, following=None):
        """Device tensors of the oldest submitted batch (the current stream waits for its copy); ``following``
        is submitted first so that its transfer runs beside the compute of the returned batch."""
        if self._n_take > 0:
            prev = (self._n_take - 1) % self.depth
            if self._state[prev] == self._HANDED:  # the step that read it has been enqueued: free from here on
                self._mark_free(prev)
        if following is not None:
            self.submit(following)
        if self._n_take >= self._n_submit:
            raise RuntimeError("HostPrefetcher.next() without a submitted batch")
        slot = self._n_take % self.depth
        torch.cuda.current_stream(self.device).wait_event(self._ready[slot])
        self._state[slot] = self._HANDED
        self._n_take += 1
        return self._sets[slot]

    def release(self):
        """Optional: call after the compute of the batch last returned by next() has been enqueued; its staging
        set may be overwritten once that work completes (next() does the same on its following call)."""
        if self._n_take == 0:
            return
        slot = (self._n_take - 1) % self.depth
        if self._state[slot] == self._HANDED:
            self._mark_free(slot)
