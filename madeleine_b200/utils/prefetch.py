"""Host→device staging one batch ahead on a copy stream, so the H2D transfer of step i+1 overlaps the kernels of
step i (the reference copies synchronously inside forward, Model.py:113).

The device side is a two-slot ring of reusable buffers: no allocator traffic per step (a fresh 131 MB allocation per
batch, kept alive by ``record_stream``, makes the caching allocator grow and synchronise when steps are short) and no
``record_stream`` bookkeeping — an event recorded on the consumer's stream tells the copy stream when a slot may be
overwritten."""
from __future__ import annotations

import torch


class DevicePrefetcher:
    """Wraps an iterable of batches whose tensors live in (ideally pinned) host memory.

    Yields the same structure with tensors on ``device``.  The copy of the next batch is issued on a side stream
    before the current one is handed out; consumers just use the tensors on the current stream.  A yielded batch stays
    valid until the NEXT batch is requested: requesting batch i+1 starts the copy of batch i+2 into batch i's slot, ordered
    after the work enqueued on the current stream up to that moment — kernels on batch i must be enqueued before asking
    for batch i+1; clone what must live longer.
    """

    SLOTS = 2

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._buffers = [dict() for _ in range(self.SLOTS)]       # slot -> {leaf index: flat device buffer}
        self._free = [None] * self.SLOTS                          # slot -> event: consumer is done with the slot
        self._slot = 0
        self._leaf = 0
        self._next = None
        self._preload()

    # -- structure walk ---------------------------------------------------------------------------------------------
    def _stage(self, obj, slot):
        if isinstance(obj, torch.Tensor):
            if obj.is_cuda:
                return obj
            i = self._leaf
            self._leaf += 1
            buf = self._buffers[slot].get(i)
            if buf is None or buf.dtype != obj.dtype or buf.numel() < obj.numel():
                buf = torch.empty(max(obj.numel(), 1), dtype=obj.dtype, device=self.device)
                self._buffers[slot][i] = buf
            dst = buf[: obj.numel()].view(obj.shape)
            dst.copy_(obj, non_blocking=True)
            return dst
        if isinstance(obj, dict):
            out = {k: self._stage(v, slot) for k, v in obj.items() if k != "_h2d_done"}
            if isinstance(obj.get("_h2d_done"), torch.cuda.Event):
                obj["_h2d_done"].record(self.copy_stream)     # lets the producer know when its host buffer may be refilled
            return out
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._stage(v, slot) for v in obj)
        return obj

    def _preload(self):
        try:
            batch = next(self.it)
        except StopIteration:
            self._next = None
            return
        slot = self._slot
        self._slot = (slot + 1) % self.SLOTS
        with torch.cuda.stream(self.copy_stream):
            if self._free[slot] is not None:
                self.copy_stream.wait_event(self._free[slot])      # the consumer of this slot's previous batch has finished
            self._leaf = 0
            self._next = (slot, self._stage(batch, slot))

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.copy_stream)
        slot, batch = self._next
        # every kernel enqueued on `cur` so far belongs to earlier batches: the OTHER slot may be overwritten once they ran
        other = (slot + 1) % self.SLOTS
        ev = self._free[other] or torch.cuda.Event()
        ev.record(cur)
        self._free[other] = ev
        self._preload()
        return batch
