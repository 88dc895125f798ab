"""Yardstick helpers for the BASELINE-size parity tests (test infrastructure: drives the oracle, never the product).

``oracle_train_step`` evaluates the reference's training step — MADELEINE.forward(train=True) (Model.py:110-159) +
calculate_losses (trainer.py:20-77) + backward — with the oracle's functions ON THE GPU IN FP64, in three stages so that
the fp64 autograd tape of a 65 k - 330 k token batch never has to live at once:

  1. no-grad forward of the encoder over chunks of flattened (slide, stain) rows -> slide / token embeddings;
  2. the losses on leaf copies of those embeddings (this is where every case meets every other: InfoNCE over the batch,
     GOT's batch-wide thresholds) -> d loss / d embeddings;
  3. a second, differentiable forward per chunk, back-propagating the chunk's slice of those gradients into the parameters.

The values are those of one monolithic evaluation (the chain rule applied at the embedding boundary).  Quirk Q1 (the
flattened row r = slide * n_mod + stain receives stain code r // bs, Model.py:126-129) is kept by computing the codes
from the GLOBAL row index before chunking.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

import oracle
from oracle.madeleine_oracle import _lin


def to_oracle_sd(sd_cpu, device, dtype=torch.float64):
    return {k: v.detach().to(device=device, dtype=dtype).requires_grad_(True) for k, v in sd_cpu.items()}


def _encode_rows(sd, x_rows, codes_rows, want_tokens, n_views=1, activation="softmax"):
    """x_rows [r, T, D] (+ per-row stain codes) -> (slide [r, v, 512], tokens [r, T, 128] or None, raw logits [r, T, 1, H])."""
    r, T, _ = x_rows.shape
    x = x_rows
    if codes_rows is not None:
        enc = sd["embedding.weight"][codes_rows]
        x = torch.cat([x, enc.unsqueeze(1).expand(-1, T, -1)], dim=-1)
    slide, tok, raw = oracle.abmil_embedder(sd, x, n_views=n_views, activation=activation)
    E, H = slide.shape[-2], slide.shape[-1]
    slide = _lin(slide.reshape(r, -1, E * H), sd["projector.weight"], sd["projector.bias"])
    tokens = None
    if want_tokens:
        tokens = _lin(tok.reshape(r, T, -1), sd["token_projector.weight"], sd["token_projector.bias"])
    return slide, tokens, raw


def _pack_dicts(slide, tokens, bs, n_mod, modalities):
    """[R, v, 512] / [R, T, 128] -> the reference's per-modality dicts (HE carries the repeated trailing dim, Model.py:153-155)."""
    slide = slide.reshape(bs, n_mod, slide.shape[1], slide.shape[2])
    embs, toks = {}, {}
    if tokens is not None:
        tokens = tokens.reshape(bs, n_mod, tokens.shape[1], tokens.shape[2])
    for i, m in enumerate(modalities):
        s = slide[:, i]
        t = tokens[:, i] if tokens is not None else None
        if m == "HE":
            s = s.unsqueeze(3).repeat(1, 1, 1, n_mod - 1)
            if t is not None:
                t = t.unsqueeze(3).repeat(1, 1, 1, n_mod - 1)
        embs[m] = s
        toks[m] = t
    return embs, toks


def oracle_train_step(sd, feats, modalities: List[str], labels_wo_he, *, stain_encoding: bool, temperature: float,
                      use_local: bool, loss_seed: Optional[int] = None, chunk_rows: int = 16, symmetric: bool = True):
    """feats [bs, n_mod, T, D] in sd's dtype/device.  Returns (loss, embs dict, toks dict); parameter gradients are left in
    ``sd[k].grad``."""
    bs, n_mod, T, D = feats.shape
    R = bs * n_mod
    x = feats.reshape(R, T, D)
    codes = (torch.arange(R, device=feats.device) // bs) if stain_encoding else None
    chunks = [slice(a, min(a + chunk_rows, R)) for a in range(0, R, chunk_rows)]
    with torch.no_grad():
        parts = [_encode_rows(sd, x[c], None if codes is None else codes[c], use_local) for c in chunks]
    slide = torch.cat([p[0] for p in parts]).requires_grad_(True)
    tokens = torch.cat([p[1] for p in parts]).requires_grad_(True) if use_local else None
    embs, toks = _pack_dicts(slide, tokens, bs, n_mod, modalities)
    if loss_seed is not None:
        torch.manual_seed(loss_seed)
    loss, flag = oracle.calculate_losses(modalities[1:], embs, toks, labels_wo_he, temperature=temperature, symmetric=symmetric,
                                         use_local=use_local)
    assert flag
    loss.backward()
    g_slide, g_tok = slide.grad, (tokens.grad if tokens is not None else None)
    for c in chunks:
        s, t, _ = _encode_rows(sd, x[c], None if codes is None else codes[c], use_local)
        outs, grads = [s], [g_slide[c]]
        if use_local:
            outs.append(t)
            grads.append(g_tok[c])
        torch.autograd.backward(outs, grads)
    return loss.detach(), {k: v.detach() for k, v in embs.items()}, toks


def oracle_packed_infonce_step(sd, x, cu: Sequence[int], temperature: float, symmetric: bool = True):
    """BASELINE configs[1]: ragged bags, the reference can only loop bs = 1.  Bags [0, R/2) are the H&E slides
    of the R/2 cases, bags [R/2, R) their IHC slides.  Returns (loss, slide embeddings [R, 512]); grads in sd."""
    R = len(cu) - 1
    with torch.no_grad():
        emb = torch.cat([_encode_rows(sd, x[cu[r]:cu[r + 1]][None], None, False)[0][:, 0] for r in range(R)])
    leaf = emb.clone().requires_grad_(True)
    loss = oracle.info_nce(leaf[:R // 2], leaf[R // 2:], temperature=temperature, symmetric=symmetric)
    loss.backward()
    for r in range(R):
        s = _encode_rows(sd, x[cu[r]:cu[r + 1]][None], None, False)[0][:, 0]
        s.backward(leaf.grad[r:r + 1])
    return loss.detach(), emb


def grad_report(named_grads: Dict[str, Optional[torch.Tensor]], sd) -> Dict[str, float]:
    """Per-parameter relative gradient error  |g - g_ref| / max(|g_ref|, 1e-6 |g_ref of all parameters|)  (2-norms, fp64).
    The floor matters for parameters whose exact gradient is zero — attention_c.bias under a softmax over the tokens
    (shift invariance): the fp64 oracle leaves ~1e-17 there, any fp32 evaluation ~1e-8."""
    total = sum(float(v.grad.double().norm()) ** 2 for v in sd.values() if v.grad is not None) ** 0.5
    out = {}
    for name, g in named_grads.items():
        ref = sd[name].grad
        if ref is None:
            assert g is None or float(g.abs().max()) == 0.0, f"{name}: the oracle has no gradient but the CUDA path does"
            continue
        assert g is not None, f"{name}: gradient missing"
        out[name] = float((g.double() - ref.double()).norm()) / max(float(ref.double().norm()), 1e-6 * total)
    return out


def rank_statistics(raw_ours: torch.Tensor, raw_ref: torch.Tensor) -> Dict[str, float]:
    """'Attention indices' over the WHOLE ranking.  raw_* [B, T, 1, H] raw attention logits.  Returns the fraction of rank
    positions holding the same token, and — over the positions that differ — the largest gap, measured in the REFERENCE's
    logits, between the token we put there and the token the reference put there (how close a tie had to be to flip)."""
    ours = raw_ours.squeeze(2).transpose(1, 2).double()
    ref = raw_ref.squeeze(2).transpose(1, 2).double()
    o_ours = ours.argsort(dim=-1, descending=True, stable=True)
    o_ref = ref.argsort(dim=-1, descending=True, stable=True)
    same = o_ours == o_ref
    gap = (torch.gather(ref, -1, o_ours) - torch.gather(ref, -1, o_ref)).abs()
    ref_sorted = torch.gather(ref, -1, o_ref)
    spacing = (ref_sorted[..., :-1] - ref_sorted[..., 1:])
    return {"identical_rank_fraction": float(same.double().mean()),
            "mismatched_positions": int((~same).sum()),
            "max_reference_logit_gap_at_mismatch": float(gap.max()),
            "max_abs_logit_error": float((ours - ref).abs().max()),
            "median_reference_neighbour_spacing": float(spacing.median()),
            "top8_identical": bool(torch.equal(o_ours[..., :8], o_ref[..., :8]))}
