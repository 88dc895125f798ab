"""Slide-embedding extraction driver (reference: madeleine/utils/utils.py:27-66 +
bin/extract_slide_embeddings.py, which run bs = 1, one blocking H2D and one blocking D2H per slide).

Here bags of any length are packed into token-budgeted batches, staged to the device one batch ahead on a copy stream
(straight from the caller's memory when it is pinned, through pinned staging buffers otherwise), encoded with
``MADELEINE.encode_packed`` and collected on the device; the embeddings come back in one transfer.  Under torch.distributed each rank takes every world-th slide (replicas only, no collective)."""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np
import torch



def plan_batches(lengths: Sequence[int], token_budget: int) -> List[List[int]]:
    """Greedy in-order packing of slide indices into batches of at most ``token_budget`` tokens (>= 1 slide each)."""
    batches, cur, tok = [], [], 0
    for i, n in enumerate(lengths):
        if cur and tok + n > token_budget:
            batches.append(cur)
            cur, tok = [], 0
        cur.append(i)
        tok += n
    if cur:
        batches.append(cur)
    return batches


class _PackedStager:
    """Packs the bags of one batch into a device buffer ``[sum N, D]`` on a copy stream, one batch ahead of the encoder.

    Pinned bags (``DataLoader(pin_memory=True)``, ``tensor.pin_memory()``) are copied straight from where they are — one
    async H2D per bag, no host-side staging copy, the link is the only cost.  Pageable bags are first gathered into one of
    three pinned staging buffers (an event per buffer keeps the host from refilling one the copy engine still reads) and go
    over in a single transfer.  The device side is a two-slot ring; an event recorded on the encoder's stream releases a slot.
    """

    def __init__(self, bags, batches, device):
        self.bags, self.batches, self.device = bags, batches, torch.device(device)
        self.cap = max(sum(bags[i].shape[0] for i in b) for b in batches)
        self.D = bags[0].shape[1]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.dev = [torch.empty(self.cap, self.D, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.slot_free = [None, None]
        n_cu = max(len(b) for b in batches) + 1
        self.cu_host = [torch.empty(n_cu, dtype=torch.int32, pin_memory=True) for _ in range(2)]
        self.cu_dev = [torch.empty(n_cu, dtype=torch.int32, device=self.device) for _ in range(2)]
        self.cu_done = [None, None]
        self.host = None                                     # pinned staging, allocated on first pageable bag
        self.host_done = [None, None, None]
        self.k_host = 0

    def _stage_pageable(self, rows, dst):
        if self.host is None:
            # page-locked directly (torch's caching host allocator keeps the blocks across calls); `.pin_memory()` on a fresh
            # pageable tensor would allocate it twice and copy 3 x cap x D floats of garbage inside the caller's timed region
            # ... first-touched from the GPU's NUMA node (hostmem.bind_to_gpu), the caller's affinity restored afterwards
            import os
            from . import hostmem
            saved = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
            hostmem.bind_to_gpu(self.device.index if self.device.index is not None else torch.cuda.current_device())
            try:
                self.host = [hostmem.pinned_empty((self.cap, self.D)) for _ in range(3)]
            finally:
                if saved is not None:
                    os.sched_setaffinity(0, saved)
        h = self.k_host % 3
        self.k_host += 1
        if self.host_done[h] is not None:
            self.host_done[h].synchronize()
        buf, o = self.host[h], 0
        for x in rows:
            buf[o:o + x.shape[0]].copy_(x)
            o += x.shape[0]
        dst[:o].copy_(buf[:o], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.copy_stream)
        self.host_done[h] = ev

    def issue(self, k):
        """Start the transfer of batch k; returns (features view, cu_seqlens device tensor, slot)."""
        slot = k % 2
        b = self.batches[k]
        cu = [0]
        for i in b:
            cu.append(cu[-1] + self.bags[i].shape[0])
        with torch.cuda.stream(self.copy_stream):
            if self.slot_free[slot] is not None:
                self.copy_stream.wait_event(self.slot_free[slot])
            dst = self.dev[slot]
            run, run_start = [], 0                           # consecutive pageable bags share one staging copy
            for j, i in enumerate(b):
                x = self.bags[i]
                if x.is_pinned() and x.is_contiguous() and x.dtype == torch.float32:
                    if run:
                        self._stage_pageable(run, dst[run_start:cu[j]])
                        run = []
                    dst[cu[j]:cu[j + 1]].copy_(x, non_blocking=True)
                else:
                    if not run:
                        run_start = cu[j]
                    run.append(x)
            if run:
                self._stage_pageable(run, dst[run_start:cu[-1]])
            # pinned, so that the copy is truly asynchronous: a pageable source makes cudaMemcpyAsync wait for everything
            # queued on the stream before it (the whole batch) and stalls the host for the length of the transfer
            if self.cu_done[slot] is not None:
                self.cu_done[slot].synchronize()             # issued two batches ago: long done, but the host must not race it
            self.cu_host[slot][: len(cu)] = torch.tensor(cu, dtype=torch.int32)
            cu_dev = self.cu_dev[slot][: len(cu)]
            cu_dev.copy_(self.cu_host[slot][: len(cu)], non_blocking=True)
            self.cu_done[slot] = self.cu_done[slot] or torch.cuda.Event()
            self.cu_done[slot].record(self.copy_stream)
        return dst[:cu[-1]], cu_dev, slot

    def release(self, slot):
        ev = self.slot_free[slot] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.slot_free[slot] = ev


@torch.no_grad()
def extract_slide_embeddings(model, bags: Sequence[torch.Tensor], device, token_budget: int = 131072,
                             rank: int = 0, world: int = 1) -> Tuple[np.ndarray, List[int]]:
    """Encode ``bags`` (list of [N_i, D] fp32 CPU tensors, pinned or pageable) → ([n_local, 512] fp32 numpy, indices of the
    slides this rank encoded), in input order.  Equivalent to calling ``model.encode_he`` per slide."""
    model.eval()
    mine = list(range(rank, len(bags), world))
    if not mine:
        return np.zeros((0, 512), dtype=np.float32), []
    local = [bags[i] for i in mine]
    batches = plan_batches([b.shape[0] for b in local], token_budget)
    stager = _PackedStager(local, batches, device)
    cur = torch.cuda.current_stream(stager.device)
    outs = []
    nxt = stager.issue(0)
    for k in range(len(batches)):
        feats, cu, slot = nxt
        cur.wait_stream(stager.copy_stream)
        if k + 1 < len(batches):
            nxt = stager.issue(k + 1)                         # overlaps the encoder kernels of batch k
        outs.append(model.encode_packed(feats, cu))
        stager.release(slot)
    emb = torch.cat(outs, dim=0).float().cpu().numpy()
    return emb, mine
