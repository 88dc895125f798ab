"""One-process-per-GPU data parallelism for the hot path (replaces the reference's nn.DataParallel,
madeleine/utils/setup_components.py:185-187).

Cases are sharded across ranks; the encoder is embarrassingly parallel per bag.  The only data-path exchange is ONE
NCCL all-gather of the slide embeddings (and stain-availability mask) before the contrastive loss; every rank then
evaluates the full in-batch loss and back-propagates into its own slice, so parameter gradients must be SUMMED
(not averaged) across ranks — ``allreduce_gradients`` does that with one flat all-reduce.  The messages are tiny
(82 KB of embeddings per rank, 20 MB of gradients), i.e. latency-bound: NCCL over NVLink/NVSwitch is the right tool
and there is no compute to overlap a hand-written peer-memory transfer with.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist

from . import ops


class PeerExchange:
    """Symmetric peer-mapped buffers + the one-kernel exchanges of csrc/peer.cu (``mdl_peer_allgather_f32`` /
    ``mdl_peer_allreduce_f32``) for the two latency-bound messages of a case-sharded step.  torch's symmetric-memory
    allocator is used for what it is — allocation and the exchange of peer pointers at start-up; the exchanges themselves
    are this package's kernels on the compute stream.

    OPT-IN (``MADELEINE_B200_PEER=1``).  Measured on 8 x B200 (tools/peer_probe.py, profiles/r02_peer_probe_8gpu.json): 46 us
    against NCCL's 43 us for the 64 KB all-gather and 62 us against 62 us for the 2 MB all-reduce, and the same step time —
    at these sizes both are dominated by the ranks not arriving together (the forward pass of the slowest of 8 GPUs ends
    ~80 us after the fastest), not by the exchange protocol, so NCCL stays the default."""

    CAP = 1 << 20                 # floats per buffer half (4 MB): the late gradient range is 0.5 M floats
    GATHER, REDUCE = 0, 16        # signal-pad slot bases of the two channels
    _inst = None
    _failed = None

    def __init__(self, device):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:  # noqa: BLE001  (deprecated no-op on newer versions)
            pass
        self.buf = symm_mem.empty(2 * self.CAP, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        arr = ctypes.c_void_p * self.world
        self.bufs = arr(*[int(p) for p in self.hdl.buffer_ptrs])
        self.sigs = arr(*[int(p) for p in self.hdl.signal_pad_ptrs])
        if self.hdl.signal_pad_size < 4 * 32:
            raise RuntimeError("signal pad too small")
        self.epoch = {self.GATHER: 0, self.REDUCE: 0}
        self.device = device
        dist.barrier()            # every rank's pad is mapped (and still zero) before the first flag is written

    @classmethod
    def get(cls, device):
        import os
        if cls._inst is not None and cls._inst.device == device:
            return cls._inst
        if cls._failed is not None or os.environ.get("MADELEINE_B200_PEER", "0") != "1":
            return None
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and device.type == "cuda"):
            return None
        if dist.get_world_size() > 16:
            return None
        ok = torch.ones(1, device=device)
        try:
            inst = cls(device)
        except Exception as e:  # noqa: BLE001
            cls._failed = f"{type(e).__name__}: {e}"
            inst = None
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)       # all ranks or none
        if float(ok) < 1:
            cls._failed = cls._failed or "a peer rank could not map symmetric memory"
            return None
        cls._inst = inst
        return inst

    @classmethod
    def status(cls):
        if cls._inst is not None:
            return "peer memory (mdl_peer_* kernels over NVLink)"
        return f"nccl ({cls._failed})" if cls._failed else "nccl"

    def _stage(self, channel, x):
        from ._lib import stream_ptr
        self.epoch[channel] += 1
        ep = self.epoch[channel]
        off = (ep & 1) * self.CAP
        n = x.numel()
        self.buf[off:off + n].copy_(x.reshape(-1))      # this rank's contribution, stream-ordered before the exchange kernel
        return ep, off * 4, n, stream_ptr(self.device)

    def fits(self, x):
        return x.dtype == torch.float32 and x.is_cuda and 0 < x.numel() <= self.CAP and x.numel() % 4 == 0 and x.is_contiguous()

    def all_gather(self, x):
        from ._lib import call
        ep, off, n, st = self._stage(self.GATHER, x)
        out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        call("mdl_peer_allgather_f32", self.bufs, self.sigs, self.rank, self.world, off, n, out, ep & 0xFFFFFFFF, self.GATHER, st)
        return out

    def all_reduce_(self, x):
        from ._lib import call
        ep, off, n, st = self._stage(self.REDUCE, x)
        call("mdl_peer_allreduce_f32", self.bufs, self.sigs, self.rank, self.world, off, n, x, ep & 0xFFFFFFFF, self.REDUCE, st)
        return x


def small_all_reduce_(x: torch.Tensor):
    """Sum a small contiguous fp32 tensor over the ranks, in place: the peer-memory kernel when available, NCCL otherwise."""
    px = PeerExchange.get(x.device) if x.is_cuda else None
    if px is not None and px.fits(x):
        return px.all_reduce_(x)
    dist.all_reduce(x, op=dist.ReduceOp.SUM)
    return x


class _AllGatherRows(torch.autograd.Function):
    """[B_local, ...] → [world * B_local, ...] (rank-major). Backward: this rank's slice of the incoming gradient."""

    @staticmethod
    def forward(ctx, x):
        world = dist.get_world_size()
        x = x.contiguous()
        ctx.rows = x.shape[0]
        px = PeerExchange.get(x.device) if x.is_cuda else None
        if px is not None and px.fits(x):
            return px.all_gather(x)
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x)
        return out

    @staticmethod
    def backward(ctx, g):
        r = dist.get_rank()
        return g[r * ctx.rows:(r + 1) * ctx.rows]


def all_gather_rows(x: torch.Tensor) -> torch.Tensor:
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    return _AllGatherRows.apply(x)


_GRAD_SYNC = [False]


def enable_gradient_sync(flag: bool = True):
    """When on, the encoder's backward all-reduces (SUM) its flat parameter-gradient buffer once, in place, before
    handing the per-parameter views to autograd — no concatenation or copy-back (``allreduce_gradients`` then is a no-op
    for the encoder parameters and must not be called as well)."""
    _GRAD_SYNC[0] = bool(flag)


def gradient_sync_enabled() -> bool:
    return _GRAD_SYNC[0] and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class _GatheredEmbeddings(ops.EmbeddingDict):
    """Per-modality views of the all-gathered slide-embedding matrix, built only if somebody asks for them (the fused loss
    glue addresses the matrix by row and never does)."""

    def __init__(self, modalities, n_mod):
        super().__init__()
        self._modalities, self._n_mod = list(modalities), n_mod

    def __missing__(self, key):
        i = self._modalities.index(key)                       # KeyError semantics for unknown names
        world, v, bs, n_mod = self.b200_world, self.b200_n_views, self.b200_bs, self.b200_n_mod
        t = self.b200_base.view(world, v, bs, n_mod, -1)[:, :, :, i].permute(0, 2, 1, 3).reshape(world * bs, v, -1)
        if key == "HE":
            t = t.unsqueeze(-1).expand(*t.shape, n_mod - 1)
        self[key] = t
        return t


def gather_slide_embeddings(wsi_embs: Dict[str, torch.Tensor], modality_labels: torch.Tensor, global_labels_host=None):
    """All-gather the per-modality slide embeddings ([B_local, n_views, 512(, n_mod-1)]) and the availability mask in a
    single collective: everything is packed into one [B_local, F] fp32 buffer, gathered once, and unpacked.

    ``global_labels_host`` (optional, CPU tensor [B_global, n_mod]): the availability mask of the whole batch when the
    loader already knows it (it comes from the case list); it is returned instead of the gathered device copy so the loss
    glue needs no device->host sync.  In that case, and when ``wsi_embs`` comes from this package's MADELEINE.forward, the
    encoder's slide-embedding matrix itself is gathered (no packing copies) and the result addresses it by row."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return wsi_embs, (modality_labels if global_labels_host is None else global_labels_host)
    base = getattr(wsi_embs, "b200_base", None)
    if base is not None and global_labels_host is not None and getattr(wsi_embs, "b200_world", 1) == 1:
        out = _GatheredEmbeddings(list(wsi_embs.keys()), wsi_embs.b200_n_mod)
        out.b200_set(all_gather_rows(base), wsi_embs.b200_bs, wsi_embs.b200_n_mod, wsi_embs.b200_n_views, world=dist.get_world_size())
        return out, global_labels_host
    keys = list(wsi_embs.keys())
    B = modality_labels.shape[0]
    dev = wsi_embs[keys[0]].device
    # HE carries a stride-0 repeated trailing dim: ship one copy
    parts, shapes = [], []
    for k in keys:
        t = wsi_embs[k]
        if k == "HE" and t.dim() == 4:
            t = t[..., 0]
        shapes.append(tuple(t.shape[1:]))
        parts.append(t.reshape(B, -1).float())
    parts.append(modality_labels.to(dev).float().reshape(B, -1))
    packed = torch.cat(parts, dim=1)
    gathered = all_gather_rows(packed)
    out, o = {}, 0
    n_mod = modality_labels.shape[1]
    for k, shp in zip(keys, shapes):
        n = 1
        for s in shp:
            n *= s
        t = gathered[:, o:o + n].reshape((gathered.shape[0],) + shp)
        if k == "HE":
            t = t.unsqueeze(-1).expand(*t.shape, n_mod - 1)
        out[k] = t
        o += n
    labels = gathered[:, o:].detach() if global_labels_host is None else global_labels_host
    return out, labels


def allreduce_gradients(module: torch.nn.Module):
    """Sum parameter gradients across ranks with one flat NCCL all-reduce."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    o = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[o:o + n].view_as(g))
        o += n


def got_sharded(v_local: torch.Tensor, q_local: torch.Tensor, m_global: int, subsample: int = 256) -> torch.Tensor:
    """Local Graph-OT loss with cases sharded across ranks.  v_local / q_local: [m_local, T, 128] token embeddings of this
    rank's cases that have the stain (m_local may be 0); m_global: number of such cases over all ranks.

    Reproduces the reference's subsampling quirk (Q3, loss.py:281-284) on the GLOBAL batch: one permutation of
    range(m_global) — rank 0's draw from torch's CPU generator, broadcast — indexes the token axis on every rank.
    Returns the sum of the local problems' losses; summing over ranks gives the reference's value for the whole batch."""
    from . import ops
    dev = v_local.device
    perm = torch.randperm(m_global)[:subsample].to(dev)
    dist.broadcast(perm, src=0)
    return ops.GOTShardedFn.apply(v_local[:, perm, :], q_local[:, perm, :])


def calculate_losses_sharded(STAINS, loss_fn_interMod, use_local, wsi_embs_global, token_embs_local, labels_global_withoutHE,
                             args, rank: int, b_local: int):
    """Sharded counterpart of utils.trainer.calculate_losses (reference trainer.py:20-77): slide embeddings are the
    all-gathered ones (global batch), token embeddings and the local-loss problems stay on their rank.

    Returns (loss, flag, parts).  ``loss`` = InfoNCE(global batch) + local_loss_weight * GOT(this rank's problems) is the
    tensor to call ``backward()`` on: the all-gather's backward hands every rank only its own rows of dInfoNCE/dE, so
    the parameter gradients SUMMED over ranks (encoder gradient sync) are exactly those of
    InfoNCE + GOT(all problems).  The value of that global objective is  parts['global'] + all_reduce(parts['local'])."""
    dev = wsi_embs_global["HE"].device
    labels = labels_global_withoutHE.detach().to("cpu").bool()
    counts = labels.sum(dim=0).tolist()
    lo, hi = rank * b_local, (rank + 1) * b_local
    g_terms, l_terms, flag = [], [], False
    for s_idx, stain in enumerate(STAINS):
        if counts[s_idx] <= 1:
            continue
        rows = labels[:, s_idx].nonzero(as_tuple=True)[0]
        if loss_fn_interMod:
            idx = rows.to(dev, non_blocking=True)
            he = wsi_embs_global["HE"][:, 0, :, s_idx][idx]
            ihc = wsi_embs_global[stain][:, 0, :][idx]
            g_terms.append(loss_fn_interMod(query=he, positive_key=ihc, symmetric=args.symmetric_cl))
        if use_local:
            mine = (rows[(rows >= lo) & (rows < hi)] - lo).to(dev, non_blocking=True)
            he_t = token_embs_local["HE"][:, :, :, s_idx][mine]
            ihc_t = token_embs_local[stain][mine]
            l_terms.append(got_sharded(he_t, ihc_t, counts[s_idx]) * args.local_loss_weight)
        flag = True
    zero = torch.zeros((), device=dev)
    parts = {"global": sum(g_terms) if g_terms else zero, "local": sum(l_terms) if l_terms else zero}
    if not (g_terms or l_terms):
        return -1, flag, parts
    return parts["global"] + parts["local"], flag, parts
