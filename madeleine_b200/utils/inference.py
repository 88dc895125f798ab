"""Slide-embedding extraction driver (reference: madeleine/utils/utils.py:27-66 +
bin/extract_slide_embeddings.py, which run bs = 1, one blocking H2D and one blocking D2H per slide).

Here bags of any length are packed into token-budgeted batches, staged to the device one batch ahead on a copy stream
from pinned memory, encoded with ``MADELEINE.encode_packed`` and collected on the device; the embeddings come back in
one transfer.  Under torch.distributed each rank takes every world-th slide (replicas only, no collective)."""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np
import torch

from .prefetch import DevicePrefetcher


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


def _host_batches(bags: Sequence[torch.Tensor], batches: List[List[int]], pin: bool):
    """Yield (packed features [sum N, D] in pinned memory, cu_seqlens int32) per batch, reusing two staging buffers."""
    cap = max(sum(bags[i].shape[0] for i in b) for b in batches)
    D = bags[0].shape[1]
    stage = [torch.empty(cap, D, dtype=torch.float32).pin_memory() if pin else torch.empty(cap, D) for _ in range(3)]
    for k, b in enumerate(batches):
        buf = stage[k % 3]
        o, cu = 0, [0]
        for i in b:
            n = bags[i].shape[0]
            buf[o:o + n].copy_(bags[i])
            o += n
            cu.append(o)
        yield {"feats": buf[:o], "cu": torch.tensor(cu, dtype=torch.int32)}


@torch.no_grad()
def extract_slide_embeddings(model, bags: Sequence[torch.Tensor], device, token_budget: int = 131072,
                             rank: int = 0, world: int = 1) -> Tuple[np.ndarray, List[int]]:
    """Encode ``bags`` (list of [N_i, D] fp32 CPU tensors) → ([n_local, 512] fp32 numpy, indices of the slides this rank
    encoded), in input order.  Equivalent to calling ``model.encode_he`` per slide."""
    model.eval()
    mine = list(range(rank, len(bags), world))
    if not mine:
        return np.zeros((0, 512), dtype=np.float32), []
    local = [bags[i] for i in mine]
    batches = plan_batches([b.shape[0] for b in local], token_budget)
    outs = []
    for batch in DevicePrefetcher(_host_batches(local, batches, pin=True), device):
        outs.append(model.encode_packed(batch["feats"], batch["cu"]))
    emb = torch.cat(outs, dim=0).float().cpu().numpy()
    return emb, mine
