"""Deterministic synthetic checkpoints and inputs shared by the golden generator and the tests.

The published MADELEINE weights need a network download, so parity runs on seeded
random-init weights.  Tensors are drawn one by one from a seeded ``torch.Generator`` on CPU so that
a 20 MB checkpoint never has to be committed; ``checksum`` lets a fixture detect RNG drift.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float64) * 2.0 - 1.0).mul_(bound).float()


def make_state_dict(seed: int = 0, n_mod: int = 1, stain_encoding: bool = False, d_in: int = 512,
                    hidden: int = 512, n_heads: int = 4) -> "OrderedDict[str, torch.Tensor]":
    """Checkpoint with the reference layout: nn.Linear-like uniform init, perturbed LayerNorm."""
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    se = 32 if stain_encoding else 0

    def linear(name, out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = _uniform(g, (out_f, in_f), b)
        sd[name + ".bias"] = _uniform(g, (out_f,), b)

    def lnorm(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g, dtype=torch.float64)).float()
        sd[name + ".bias"] = (0.1 * torch.randn(c, generator=g, dtype=torch.float64)).float()

    if stain_encoding:
        sd["embedding.weight"] = torch.randn(n_mod, 32, generator=g, dtype=torch.float64).float()
    linear("token_projector", 128, hidden * n_heads)
    linear("wsi_embedders.pre_attn.0", hidden, d_in + se)
    lnorm("wsi_embedders.pre_attn.1", hidden)
    linear("wsi_embedders.pre_attn.4", hidden, hidden)
    lnorm("wsi_embedders.pre_attn.5", hidden)
    linear("wsi_embedders.pre_attn.8", hidden * n_heads, hidden)
    lnorm("wsi_embedders.pre_attn.9", hidden * n_heads)
    for h in range(n_heads):
        linear(f"wsi_embedders.attn.{h}.attention_a.0", 512, hidden)
        linear(f"wsi_embedders.attn.{h}.attention_b.0", 512, hidden)
        linear(f"wsi_embedders.attn.{h}.attention_c", 1, 512)
    linear("projector", 512, hidden * n_heads)
    return sd


def make_feats(seed: int, *shape: int) -> torch.Tensor:
    """N(0,1) stand-in for CONCH patch embeddings."""
    g = torch.Generator().manual_seed(7919 * seed + 5)
    return torch.randn(*shape, generator=g, dtype=torch.float32)


def make_ragged_lengths(seed: int, n_bags: int, lo: int, hi: int):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(lo, hi + 1, (n_bags,), generator=g).tolist()


def checksum(sd) -> float:
    """Order-dependent fp64 checksum of a mapping of tensors (or a single tensor)."""
    if isinstance(sd, torch.Tensor):
        sd = {"x": sd}
    acc = 0.0
    for i, (k, v) in enumerate(sd.items()):
        acc += (i + 1) * float(v.double().abs().sum()) + float(v.double().sum())
    return acc
