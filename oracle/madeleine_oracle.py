"""Functional CPU restatement of the MADELEINE hot path (oracle; test infrastructure only).

Every function takes a flat ``state_dict``-style mapping ``sd`` (reference checkpoint
keys) and plain tensors; nothing here is an ``nn.Module`` and nothing
is shared with the product package.  Arithmetic runs in the dtype of the inputs (fp32 for
parity with the reference, fp64 for a tighter ground truth).

Citations are to /root/reference (mahmoodlab/MADELEINE @ 419287dc).
Parity is pinned by tests/golden fixtures produced by the real reference
(tests/golden/make_golden.py); the reference itself has no tests.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

__all__ = [
    "layer_norm", "gelu_erf", "pre_attn", "gated_attention_logits", "abmil_embedder",
    "encode_he", "encode_packed", "madeleine_forward_train", "madeleine_forward_eval",
    "madeleine_forward_attention", "info_nce", "cosine_cost", "ipot_plan", "ipot_distance",
    "thresholded_cosine_cost", "gromov_wasserstein", "got", "calculate_losses",
    "topk_attention_indices",
]

HE_POSITION = 0
LN_EPS = 1e-5  # nn.LayerNorm default, madeleine/models/Model.py:352,356,360


# --------------------------------------------------------------------------------------
# Encoder (madeleine/models/Model.py:346-451, madeleine/models/abmil.py:41-68)
# --------------------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = LN_EPS) -> torch.Tensor:
    """Row-wise LayerNorm with biased variance (torch.nn.LayerNorm semantics)."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc / torch.sqrt(var + eps) * gamma + beta


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """Exact (erf) GELU, nn.GELU() default (Model.py:353)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _lin(x, w, b):
    return x @ w.transpose(-1, -2) + b


def pre_attn(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "wsi_embedders.") -> torch.Tensor:
    """Three Linear→LayerNorm→GELU stages (Model.py:350-363); dropout is identity (eval)."""
    h = x
    for lin, ln in ((0, 1), (4, 5), (8, 9)):
        h = _lin(h, sd[f"{prefix}pre_attn.{lin}.weight"], sd[f"{prefix}pre_attn.{lin}.bias"])
        h = layer_norm(h, sd[f"{prefix}pre_attn.{ln}.weight"], sd[f"{prefix}pre_attn.{ln}.bias"])
        h = gelu_erf(h)
    return h


def gated_attention_logits(sd, x_head: torch.Tensor, head: int, prefix: str = "wsi_embedders.") -> torch.Tensor:
    """Raw gated-attention logit of one head, abmil.py:49-52: (tanh(xWa+ba) * sigmoid(xWb+bb)) Wc + bc."""
    p = f"{prefix}attn.{head}."
    a = torch.tanh(_lin(x_head, sd[p + "attention_a.0.weight"], sd[p + "attention_a.0.bias"]))
    b = torch.sigmoid(_lin(x_head, sd[p + "attention_b.0.weight"], sd[p + "attention_b.0.bias"]))
    return _lin(a * b, sd[p + "attention_c.weight"], sd[p + "attention_c.bias"])  # [..., T, 1]


def _n_heads(sd, prefix):
    h = 0
    while f"{prefix}attn.{h}.attention_c.weight" in sd:
        h += 1
    return h


def abmil_embedder(sd, bags: torch.Tensor, n_views: int = 1, prefix: str = "wsi_embedders.",
                   activation: str = "softmax"):
    """ABMILEmbedder.forward (Model.py:375-451).

    bags [B,T,Din] → (slide [B,E,H] or [B,3,E,H], tokens [B,T,E,H], raw_attention [B,T,1,H]).
    Head ``c`` reads channels ``c::H`` of the pre-attention output (einops 'b t (e c) -> b t e c').
    With n_views=3 the two half views use numpy's global RNG exactly like Model.py:427-430.
    """
    H = _n_heads(sd, prefix)
    h3 = pre_attn(sd, bags, prefix)
    B, T, C = h3.shape
    emb = h3.reshape(B, T, C // H, H)
    raw = torch.stack([gated_attention_logits(sd, emb[:, :, :, c], c, prefix) for c in range(H)], dim=-1)
    if activation == "softmax":
        attn = torch.softmax(raw, dim=1)
    elif activation == "leaky_relu":
        attn = torch.nn.functional.leaky_relu(raw)
    elif activation == "relu":
        attn = torch.relu(raw)
    elif activation == "sigmoid":
        attn = torch.sigmoid(raw)
    else:
        raise NotImplementedError("Activation not implemented.")
    whole = (emb * attn).sum(dim=1)  # [B,E,H]
    if n_views == 1:
        return whole, emb, raw
    order = np.arange(T)
    np.random.shuffle(order)
    mid = len(order) // 2
    views = [whole.unsqueeze(1)]
    for idx in (order[:mid], order[mid:]):
        sub_attn = torch.softmax(raw[:, idx], dim=1)
        views.append((emb[:, idx] * sub_attn).sum(dim=1).unsqueeze(1))
    return torch.cat(views, dim=1), emb, raw


def encode_he(sd, feats: torch.Tensor) -> torch.Tensor:
    """MADELEINE.encode_he (Model.py:97-107): [bs,T,D] → [bs,512]; never adds stain encodings (quirk Q2)."""
    slide, _, _ = abmil_embedder(sd, feats)
    bs = feats.shape[0]
    return _lin(slide.reshape(bs, -1), sd["projector.weight"], sd["projector.bias"])


def encode_packed(sd, feats: torch.Tensor, cu_seqlens: Sequence[int]) -> torch.Tensor:
    """Variable-length adapter: the reference cannot batch unequal bags, so loop bs=1 (SURVEY §8d cfg 2)."""
    out = []
    for r in range(len(cu_seqlens) - 1):
        out.append(encode_he(sd, feats[cu_seqlens[r]:cu_seqlens[r + 1]].unsqueeze(0)))
    return torch.cat(out, dim=0)


def madeleine_forward_train(sd, feats: torch.Tensor, modalities: List[str], stain_encoding: bool = False,
                            n_views: int = 1, activation: str = "softmax"):
    """MADELEINE.forward(train=True) (Model.py:110-159), including quirk Q1 (row r gets stain code r // bs)."""
    bs, n_mod, T, D = feats.shape
    x = feats.reshape(bs * n_mod, T, D)
    if stain_encoding:
        ind = torch.tensor([i for i in range(n_mod) for _ in range(bs)], dtype=torch.long, device=feats.device)
        enc = sd["embedding.weight"][ind]                       # [bs*n_mod, 32]
        x = torch.cat([x, enc.unsqueeze(1).expand(-1, T, -1)], dim=-1)
    slide, tok, _ = abmil_embedder(sd, x, n_views=n_views, activation=activation)      # config.activation (Model.py:63-70)
    tok = tok.reshape(bs, n_mod, T, -1)
    tok = _lin(tok, sd["token_projector.weight"], sd["token_projector.bias"])
    E, H = slide.shape[-2], slide.shape[-1]
    slide = slide.reshape(bs * n_mod, -1, E * H)
    slide = _lin(slide, sd["projector.weight"], sd["projector.bias"]).reshape(bs, n_mod, -1, E)
    embs, toks = {}, {}
    for i, m in enumerate(modalities):
        s, t = slide[:, i], tok[:, i]
        if m == "HE":
            s = s.unsqueeze(3).repeat(1, 1, 1, n_mod - 1)
            t = t.unsqueeze(3).repeat(1, 1, 1, n_mod - 1)
        embs[m], toks[m] = s, t
    return embs, toks


def madeleine_forward_eval(sd, feats: torch.Tensor, modalities: List[str], stain_encoding: bool = False,
                           custom_stain_idx: Optional[int] = None):
    """MADELEINE.forward(train=False) (Model.py:162-203); only meaningful for n_mod == 1."""
    bs, n_mod, T, D = feats.shape
    out = {}
    for s in range(n_mod):
        name = modalities[custom_stain_idx] if custom_stain_idx else modalities[s]
        x = feats[:, s]
        if stain_encoding:
            key = custom_stain_idx if custom_stain_idx else s
            enc = sd["embedding.weight"][key]
            x = torch.cat([x, enc.expand(bs, T, -1)], dim=-1)
        slide, _, _ = abmil_embedder(sd, x)
        E = slide.shape[-2]
        e = _lin(slide.reshape(bs * n_mod, -1), sd["projector.weight"], sd["projector.bias"])
        out[name] = e.reshape(bs, n_mod, E)
    return out


def madeleine_forward_attention(sd, feats: torch.Tensor):
    """MADELEINE.forward(return_attention=True) (Model.py:206-216): HE only → ([bs,1,512], raw [bs,T,1,H])."""
    bs, n_mod, T, D = feats.shape
    slide, _, raw = abmil_embedder(sd, feats[:, HE_POSITION])
    E = slide.shape[-2]
    e = _lin(slide.reshape(bs * n_mod, -1), sd["projector.weight"], sd["projector.bias"])
    return e.reshape(bs, n_mod, E), raw


def topk_attention_indices(raw_attention: torch.Tensor, k: int) -> torch.Tensor:
    """'Attention indices': top-k token ids per (bag, head) from raw logits [B,T,1,H] → [B,H,k]."""
    return raw_attention.squeeze(2).transpose(1, 2).topk(k, dim=-1).indices


# --------------------------------------------------------------------------------------
# Global loss (madeleine/utils/loss.py:58-133)
# --------------------------------------------------------------------------------------
def _l2_normalize(x, eps=1e-12):
    """F.normalize semantics: x / max(||x||, eps) (quirk Q6, loss.py:132-133)."""
    return x / x.norm(dim=-1, keepdim=True).clamp_min(eps)


def _ce_diag(logits, reduction):
    lse = torch.logsumexp(logits, dim=1)
    nll = lse - logits.diagonal()
    if reduction == "mean":
        return nll.mean()
    if reduction == "sum":
        return nll.sum()
    return nll


def info_nce(query, positive_key, temperature=0.1, reduction="mean", symmetric=False):
    """InfoNCE with in-batch negatives (loss.py:111-127). Explicit-negative modes return None upstream."""
    if query.dim() != 2:
        raise ValueError("<query> must have 2 dimensions.")
    if positive_key.dim() != 2:
        raise ValueError("<positive_key> must have 2 dimensions.")
    if len(query) != len(positive_key):
        raise ValueError("<query> and <positive_key> must must have the same number of samples.")
    if query.shape[-1] != positive_key.shape[-1]:
        raise ValueError("Vectors of <query> and <positive_key> should have the same number of components.")
    q, k = _l2_normalize(query), _l2_normalize(positive_key)
    logits = q @ k.t()
    loss = _ce_diag(logits / temperature, reduction)
    if symmetric:
        loss = 0.5 * loss + 0.5 * _ce_diag(logits.t() / temperature, reduction)
    return loss


# --------------------------------------------------------------------------------------
# Local loss: Graph Optimal Transport (madeleine/utils/loss.py:162-301)
# --------------------------------------------------------------------------------------
def cosine_cost(x, y):
    """loss.py:162-176. x [b,D,n], y [b,D,m] → cost[b,m,n] = 1 - cos(y_j, x_i); norm is x/(||x||+1e-12) (Q6)."""
    xn = x / (x.norm(dim=1, keepdim=True) + 1e-12)
    yn = y / (y.norm(dim=1, keepdim=True) + 1e-12)
    return (1.0 - xn.transpose(1, 2) @ yn).transpose(1, 2)


def _threshold_relu(c, beta=0.1):
    """Global (whole-tensor) min/max threshold + ReLU (loss.py:226-233, 288-292; quirk Q5)."""
    lo, hi = c.min(), c.max()
    return torch.relu(c - (lo + beta * (hi - lo)))


def thresholded_cosine_cost(x, y):
    """cos_batch_torch (loss.py:210-233)."""
    xn = x / (x.norm(dim=1, keepdim=True) + 1e-12)
    yn = y / (y.norm(dim=1, keepdim=True) + 1e-12)
    c = 1.0 - xn.transpose(1, 2) @ yn
    return _threshold_relu(c).transpose(1, 2)


def ipot_plan(C, beta=0.5, iteration=50):
    """IPOT_torch_batch_uniform (loss.py:179-193). C [b,n,m] → transport plan T [b,n,m]."""
    b, n, m = C.shape
    sigma = torch.ones(b, m, 1, dtype=C.dtype, device=C.device) / float(m)
    T = torch.ones(b, n, m, dtype=C.dtype, device=C.device)
    A = torch.exp(-C / beta)
    delta = None
    for _ in range(iteration):
        Q = A * T
        delta = 1.0 / (n * (Q @ sigma))
        sigma = 1.0 / (float(m) * (Q.transpose(1, 2) @ delta))
        T = delta * Q * sigma.transpose(1, 2)
    return T


def _trace(x):
    return x.diagonal(dim1=-2, dim2=-1).sum(-1, keepdim=True)


def ipot_distance(C, iteration=50):
    """IPOT_distance_torch_batch_uniform (loss.py:202-207): returns -tr(C^T T) per batch item, [b,1]."""
    T = ipot_plan(C, iteration=iteration)
    return -_trace(C.transpose(1, 2) @ T)


def gromov_wasserstein(X, Y, lamda=1e-1, iteration=5, ot_iteration=20):
    """GW_distance_uniform → GW_distance → GW_torch_batch (loss.py:236-275). X,Y [b,D,n] → [b,1]."""
    b, n_x, n_y = X.shape[0], X.shape[2], Y.shape[2]
    p = torch.ones(b, n_x, 1, dtype=X.dtype, device=X.device) / n_x
    q = torch.ones(b, n_y, 1, dtype=X.dtype, device=X.device) / n_y
    Cs = thresholded_cosine_cost(X, X)
    Ct = thresholded_cosine_cost(Y, Y)
    n, m = Cs.shape[2], Ct.shape[2]
    one_m = torch.ones(b, m, 1, dtype=X.dtype, device=X.device)
    one_n = torch.ones(b, n, 1, dtype=X.dtype, device=X.device)
    Cst = ((Cs ** 2) @ p) @ one_m.transpose(1, 2) + one_n @ (q.transpose(1, 2) @ (Ct ** 2).transpose(1, 2))
    gamma = p @ q.transpose(1, 2)
    for _ in range(iteration):
        Cg = Cst - 2.0 * (Cs @ gamma) @ Ct.transpose(1, 2)
        gamma = ipot_plan(Cg, beta=lamda, iteration=ot_iteration)
    Cg = Cst - 2.0 * (Cs @ gamma) @ Ct.transpose(1, 2)
    return _trace(Cg.transpose(1, 2) @ gamma.detach())


def got(v_, q_, subsample=None, perm: Optional[torch.Tensor] = None):
    """GOT (loss.py:278-301). v_, q_ [m,N,128] → scalar.

    Quirk Q3: the permutation is drawn over the *batch* size and indexes the token axis. ``perm`` lets a
    caller pass the permutation explicitly; otherwise torch's global CPU RNG is consumed like the reference.
    Quirk Q4: sums (not means) over the batch.
    """
    if subsample is not None:
        idx = (torch.randperm(v_.shape[0]) if perm is None else perm)[:subsample]
        v_, q_ = v_[:, idx, :], q_[:, idx, :]
    c = cosine_cost(v_.transpose(2, 1), q_.transpose(2, 1)).transpose(1, 2)
    c = _threshold_relu(c)
    wd = (-ipot_distance(c, iteration=30)).sum()
    gwd = gromov_wasserstein(v_.transpose(2, 1), q_.transpose(2, 1)).sum()
    return gwd + wd


# --------------------------------------------------------------------------------------
# Loss glue (madeleine/utils/trainer.py:20-77)
# --------------------------------------------------------------------------------------
def calculate_losses(stains, wsi_embs, token_embs, modality_labels_withoutHE, temperature=0.001, symmetric=True,
                     use_global=True, use_local=False, local_loss_weight=1.0, use_intra=False):
    """calculate_losses: per-stain mask → global InfoNCE (+ GOT, + intra-modality InfoNCE) → sum."""
    losses, flag = [], False
    for s_idx, stain in enumerate(stains):
        mask = modality_labels_withoutHE[:, s_idx].bool()
        if int(mask.sum()) > 1:
            if use_global:
                he = wsi_embs["HE"][:, 0, :, s_idx][mask]
                ihc = wsi_embs[stain][:, 0, :][mask]
                losses.append(info_nce(he, ihc, temperature=temperature, symmetric=symmetric))
            if use_local:
                he_t = token_embs["HE"][:, :, :, s_idx][mask]
                ihc_t = token_embs[stain].squeeze()[mask]
                losses.append(got(he_t, ihc_t, subsample=256) * local_loss_weight)
            if use_intra:
                he1, he2 = wsi_embs["HE"][:, 1, :, s_idx][mask], wsi_embs["HE"][:, 2, :, s_idx][mask]
                st1, st2 = wsi_embs[stain][:, 1, :][mask], wsi_embs[stain][:, 2, :][mask]
                losses.append(info_nce(he1, he2, temperature=temperature, symmetric=symmetric))
                losses.append(info_nce(st1, st2, temperature=temperature, symmetric=symmetric))
            flag = True
    if losses:
        return sum(losses), flag
    return -1, flag
