"""Host→device staging one batch ahead on a copy stream, so the H2D transfer of step i+1 overlaps the kernels of
step i (the reference copies synchronously inside forward, Model.py:113)."""
from __future__ import annotations

import torch


class DevicePrefetcher:
    """Wraps an iterable of batches whose tensors live in (ideally pinned) host memory.

    Yields the same structure with tensors on ``device``.  The copy of the next batch is issued on a side stream
    before the current one is handed out; consumers just use the tensors on the current stream.
    """

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._next = None
        self._preload()

    def _to_device(self, obj):
        if isinstance(obj, torch.Tensor):
            return obj.to(self.device, non_blocking=True)
        if isinstance(obj, dict):
            return {k: self._to_device(v) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._to_device(v) for v in obj)
        return obj

    def _record(self, obj, stream):
        if isinstance(obj, torch.Tensor) and obj.is_cuda:
            obj.record_stream(stream)
        elif isinstance(obj, dict):
            for v in obj.values():
                self._record(v, stream)
        elif isinstance(obj, (list, tuple)):
            for v in obj:
                self._record(v, stream)

    def _preload(self):
        try:
            batch = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.copy_stream):
            self._next = self._to_device(batch)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.copy_stream)
        batch = self._next
        self._record(batch, cur)
        self._preload()
        return batch
