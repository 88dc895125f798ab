"""B200-native drop-in for madeleine/models/Model.py (reference @ 419287dc).

Same classes, constructor arguments, ``forward`` signatures, return structures and ``state_dict`` keys as the
reference, so ``bin/pretrain.py`` / ``bin/extract_slide_embeddings.py`` run unchanged; the
arithmetic is done by the sm_100a kernels in ``libmadeleine_b200.so`` (see ``madeleine_b200/ops.py``).  The
``nn.Linear`` / ``nn.LayerNorm`` submodules below are parameter containers that fix the checkpoint layout — their
own ``forward`` is never called.  There is no CPU path: tensors must be on a CUDA device.

Extensions over the reference API (all optional):
  * ``MADELEINE.encode_packed(feats[M, D], cu_seqlens)`` / ``forward_packed`` — variable-length, bag-packed input
    (the reference can only batch equal-length bags);
  * ``config.b200_precision`` ∈ {'auto', 'fp32', 'bf16'} (default 'auto': follow torch autocast like the reference);
  * ``config.b200_token_window`` ∈ {'off' (default), 'batch', int} or env ``MADELEINE_B200_TOKEN_WINDOW``: the training
    forward returns token embeddings for the first W tokens of every bag only ([bs, W, 128] instead of [bs, T, 128]).
    ``GOT(..., subsample=256)`` draws its permutation over the number of CASES and uses it to index the TOKEN axis
    (loss.py:281-284, quirk Q3), so tokens past the batch size can never reach the loss: with W = batch size the losses
    and gradients of ``calculate_losses`` are unchanged while token_projector and its backward touch ~3 % of the rows.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, Optional, Union

import numpy as np
import torch
from torch import nn

from .. import ops
from .._lib import require_cuda
from .abmil import BatchedABMIL

# global magic numbers
HE_POSITION = 0


def create_model(model_cfg: Union[str, Dict], device: Union[str, torch.device] = "cpu", checkpoint_path: Optional[str] = None):
    """Model.py:15-43 — build MADELEINE (no stain encodings, quirk Q2), optionally restore a checkpoint whose keys may
    carry a DataParallel ``module.`` prefix."""
    model = MADELEINE(config=model_cfg, stain_encoding=False).to(device)
    if checkpoint_path:
        state_dict = torch.load(checkpoint_path, weights_only=False)
        if any("module" in key for key in state_dict.keys()):
            state_dict = OrderedDict((key[7:], value) for key, value in state_dict.items())
        model.load_state_dict(state_dict, strict=True)
        print("* Loaded weights successfully!")
    return model


class ABMILEmbedder(nn.Module):
    """Multi-head gated-attention MIL encoder (Model.py:314-451), fused on the GPU."""

    def __init__(self, pre_attention_params: dict = None, attention_params: dict = None, aggregation: str = "regular") -> None:
        super().__init__()
        self.pre_attention_params = pre_attention_params
        self.attention_params = attention_params
        self.n_heads = attention_params["params"]["n_heads"]
        self._build_pre_attention_params(params=pre_attention_params)
        if attention_params is not None:
            self._build_attention_params(attn_model=attention_params["model"], params=attention_params["params"])
        self.agg_type = aggregation
        self._pack_cache = {}
        self._spec_cache = {}
        self._placeholders = None
        self._flat = None

    # -- parameter containers (checkpoint layout) ---------------------------------------------------------------
    def _build_pre_attention_params(self, params):
        d_in, hid = params["input_dim"], params["hidden_dim"]
        widths = [(d_in, hid), (hid, hid), (hid, hid * self.n_heads)]
        layers = []
        for fan_in, fan_out in widths:
            layers += [nn.Linear(fan_in, fan_out), nn.LayerNorm(fan_out), nn.GELU(), nn.Dropout(0.1)]
        self.pre_attn = nn.Sequential(*layers)

    def _build_attention_params(self, attn_model="ABMIL", params=None):
        if attn_model != "ABMIL":
            raise NotImplementedError("Attention model not implemented -- Options are ABMIL")
        self.attn = nn.ModuleList([BatchedABMIL(**params) for _ in range(self.n_heads)])

    # -- kernel plumbing ----------------------------------------------------------------------------------------------
    def _check_supported(self):
        hid = self.pre_attention_params["hidden_dim"]
        ap = self.attention_params["params"]
        if hid != ops.HID or ap["hidden_dim"] != ops.GATE or self.n_heads != 4 or ap.get("n_classes", 1) != 1:
            raise NotImplementedError(
                "madeleine_b200 kernels are built for wsi_encoder_hidden_dim=512, attention hidden_dim=512, n_heads=4 "
                f"(got hidden={hid}, attn_hidden={ap['hidden_dim']}, n_heads={self.n_heads})")

    def _own_params(self):
        p = self.pre_attn
        out = [p[0].weight, p[0].bias, p[1].weight, p[1].bias, p[4].weight, p[4].bias, p[5].weight, p[5].bias,
               p[8].weight, p[8].bias, p[9].weight, p[9].bias]
        for head in self.attn:
            out += head.gate_parameters()
        return out

    def _placeholder_heads(self, device):
        """token_projector / projector stand-ins when the embedder is used without a MADELEINE parent."""
        if self._placeholders is None or self._placeholders[0].device != device:
            C = ops.HID * self.n_heads
            z = lambda *s: torch.zeros(*s, device=device)  # noqa: E731
            self._placeholders = [z(ops.TOK, C), z(ops.TOK), z(ops.HID, C), z(ops.HID)]
        return self._placeholders

    def _flat_master(self, params, spec, dev):
        """Flat fp32 concatenation of ``params`` in ops.PARAM_ORDER.  The first call re-points every parameter's storage at a
        slice of ONE flat buffer (values unchanged; state_dict / load_state_dict / optimisers keep working on the same
        Parameter objects), so afterwards the 'concatenation' is that buffer itself — no per-step torch.cat — and the flat
        gradient the backward pass produces lines up with it for the fused optimiser.  Falls back to a plain copy whenever
        the parameters cannot be re-pointed (replicas under nn.DataParallel, non-fp32 storage)."""
        fm = self._flat
        if fm is not None and fm.device == dev and fm.numel() == spec.master_numel:
            base = fm.data_ptr()
            if all(p.data_ptr() == base + 4 * o for p, o in zip(params, spec.param_offsets)):
                return fm
        flat = torch.cat([p.detach().reshape(-1).float() for p in params])
        if all(p.is_leaf and p.dtype == torch.float32 and p.device == dev for p in params):
            with torch.no_grad():
                for p, o, n in zip(params, spec.param_offsets, spec.param_numels):
                    p.data = flat[o:o + n].view(p.shape)
            self._flat = flat
        return flat

    def run_kernels(self, x, cu, codes, head_params, embedding_weight, *, se_dim, want_tokens, want_projector, want_ref_feats,
                    views=None, precision=None, token_rows=None, token_sel_of_row=None):
        """x [M, d_in] fp32 bag-packed → dict(slide, logits, tokens?, ref_feats?). Differentiable w.r.t. parameters."""
        self._check_supported()
        require_cuda(x, "patch features")
        dev = x.device
        x = x.contiguous().float()
        d_in = x.shape[1]
        d_in_total = self.pre_attention_params["input_dim"]
        if d_in % 64 != 0 or d_in + se_dim != d_in_total:
            raise NotImplementedError(
                f"feature width {d_in} (+{se_dim} stain-encoding channels) does not match input_dim={d_in_total} or is not a "
                "multiple of 64; concatenated inputs must be expressed as stain codes")
        if head_params is None:
            head_params = self._placeholder_heads(dev)
        params = self._own_params() + list(head_params)
        if se_dim > 0:
            params.append(embedding_weight)
        precision = ops.resolve_precision(precision)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        # fp32-grade inference runs on fp16 hi/lo operand planes (2^-22 per element instead of the training format's 2^-17:
        # embeddings and attention logits at the accuracy of the reference's own fp32 evaluation); MADELEINE_B200_INFER_F16=0
        # keeps the bf16 hi/lo planes of the training forward
        f16 = (not need_grad) and precision in ("fp32", "fp32_fwd") and os.environ.get("MADELEINE_B200_INFER_F16", "1") != "0"
        key = (precision, str(dev), se_dim, f16)
        spec = self._spec_cache.get((str(dev), se_dim))
        if spec is None:
            spec = ops.PackSpec([tuple(p.shape) for p in params], self.n_heads, d_in_total, dev, d_in=d_in)
            self._spec_cache[(str(dev), se_dim)] = spec
        master = self._flat_master(params, spec, dev)
        # one int per parameter: version counters only grow, so their sum changes whenever any parameter was written
        versions = (master.data_ptr(), sum(p._version for p in params))
        cached = self._pack_cache.get(key)
        if cached is None or cached[0] != versions:
            pw = ops.PackedWeights(spec, master, 1 if precision == "bf16" else 2, f16=f16)
            # one live pack per operand format: weights change every optimiser step, evaluation alternates with training
            self._pack_cache = {k: v for k, v in self._pack_cache.items() if v[0] == versions}
            self._pack_cache[key] = (versions, pw)
        else:
            pw = cached[1]
        if need_grad and d_in % 128 != 0:
            raise NotImplementedError(f"training needs a feature width that is a multiple of 128 (got {d_in}); "
                                      "inference works for any multiple of 64")
        ops._step_counter[0] += 1
        seed = (torch.initial_seed() * 1000003 + ops._step_counter[0]) & 0x7FFFFFFFFFFFFFFF
        opt = ops.EncodeOptions(n_heads=self.n_heads, activation=self.attention_params["params"]["activation"],
                                precision=precision, training=self.training, want_tokens=want_tokens,
                                want_projector=want_projector, want_ref_feats=want_ref_feats, d_in=d_in, se_dim=se_dim,
                                views=views, seed=seed, token_rows=token_rows, token_sel_of_row=token_sel_of_row)
        if opt.activation not in ops.ACT_CODES:
            raise NotImplementedError("Activation not implemented.")
        holder = {"opt": opt, "cu": cu, "codes": codes, "pw": pw, "need_grad": need_grad}
        with torch.cuda.device(dev):      # kernels launch on the current device: make it the tensors' (model.to('cuda:1') etc.)
            slide, logits, tokens, ref = ops.EncodeFn.apply(holder, x, *params)
        return {"slide": slide, "logits": logits, "tokens": tokens, "ref_feats": ref}

    _cu_cache = {}

    @staticmethod
    def uniform_cu(n_bags, n_tokens, device):
        """cu_seqlens of a dense [n_bags, n_tokens] batch; read-only, so the last few shapes are kept (no arange launch per step)."""
        key = (int(n_bags), int(n_tokens), str(device))
        cu = ABMILEmbedder._cu_cache.get(key)
        if cu is None:
            if len(ABMILEmbedder._cu_cache) >= 16:
                ABMILEmbedder._cu_cache.clear()
            cu = torch.arange(0, (n_bags + 1) * n_tokens, n_tokens, dtype=torch.int32, device=device)
            ABMILEmbedder._cu_cache[key] = cu
        return cu

    @staticmethod
    def half_views(n_bags, n_tokens, device):
        """Token index lists of the two random half views (Model.py:427-430: numpy global RNG, same halves for every bag)."""
        order = np.arange(n_tokens)
        np.random.shuffle(order)
        mid = len(order) // 2
        halves = [order[:mid], order[mid:]]
        # segment order: [view1 of bag 0..R-1 | view2 of bag 0..R-1]
        idx, lens = [], []
        row2seg = np.empty(n_bags * n_tokens, dtype=np.int32)
        for v, h in enumerate(halves):
            for r in range(n_bags):
                idx.append(h + r * n_tokens)
                lens.append(len(h))
                row2seg[h + r * n_tokens] = v * n_bags + r
        tok_idx = torch.from_numpy(np.concatenate(idx).astype(np.int32)).to(device)
        cu2 = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)).to(device)
        return tok_idx, cu2, torch.from_numpy(row2seg).to(device)

    def forward(self, bags: torch.Tensor, return_attention: bool = False, return_preattn_feats: bool = False, n_views=1):
        """bags [B, T, input_dim] → slide embeddings [B, E, H] (n_views=1) or [B, 3, E, H];
        optionally raw attention [B, T, 1, H] or the pre-attention features [B, T, E, H] (Model.py:375-451)."""
        if self.agg_type != "regular":
            raise NotImplementedError('Agg type not supported. Options are "regular".')
        B, T, D = bags.shape
        H = self.n_heads
        cu = self.uniform_cu(B, T, bags.device)
        views = self.half_views(B, T, bags.device) if n_views != 1 else None
        out = self.run_kernels(bags.reshape(B * T, D), cu, None, None, None, se_dim=0, want_tokens=False, want_projector=False,
                               want_ref_feats=return_preattn_feats and not return_attention, views=views)
        slide = out["slide"]                                  # [B (+2B), E, H]
        if n_views != 1:
            whole, halves = slide[:B], slide[B:].view(2, B, ops.HID, H).transpose(0, 1)
            slide = torch.cat([whole.unsqueeze(1), halves], dim=1)
        if return_attention:
            return slide, out["logits"].view(B, T, 1, H)
        if return_preattn_feats:
            return slide, out["ref_feats"].view(B, T, ops.HID, H)
        return slide


class MADELEINE(nn.Module):
    """Model.py:45-216."""

    def __init__(self, config, stain_encoding=False):
        super().__init__()
        self.config = config
        self.modalities = config.MODALITIES
        self.stain_encoding = stain_encoding
        self.b200_precision = getattr(config, "b200_precision", None)
        # a missing stain arrives as an all-zero bag that the reference encodes in full and then
        # masks out of every loss.  When the batch carries `modality_labels`, such bags are encoded from ONE token (all of
        # their tokens are identical, so slide and token embeddings are unchanged) — ~27 % less work on ACROBAT.
        self.b200_skip_missing_bags = bool(getattr(config, "b200_skip_missing_bags", True))
        self.b200_token_window = getattr(config, "b200_token_window", None)
        if self.stain_encoding:
            self.stain_encoding_dim = 32
            self.embedding = nn.Embedding(len(self.modalities), self.stain_encoding_dim)
        else:
            self.stain_encoding_dim = 0
        if self.config.wsi_encoder != "abmil":
            raise ValueError('Unsupported wsi_encoder. Must be "abmil". Now is {}.'.format(self.config.wsi_encoder))
        pre_params = {"input_dim": self.config.patch_embedding_dim + self.stain_encoding_dim,
                      "hidden_dim": self.config.wsi_encoder_hidden_dim}
        attention_params = {"model": "ABMIL",
                            "params": {"input_dim": self.config.wsi_encoder_hidden_dim, "hidden_dim": 512, "dropout": True,
                                       "activation": self.config.activation, "n_heads": self.config.n_heads, "n_classes": 1}}
        width = attention_params["params"]["hidden_dim"] * attention_params["params"]["n_heads"]
        # construction order matches the reference so that the same torch seed gives the same initial weights
        self.token_projector = nn.Linear(width, 128)
        self.wsi_embedders = ABMILEmbedder(pre_params, attention_params)
        self.projector = nn.Linear(width, attention_params["params"]["hidden_dim"])

    # -- helpers ----------------------------------------------------------------------------------------------------------
    def _heads(self):
        return [self.token_projector.weight, self.token_projector.bias, self.projector.weight, self.projector.bias]

    def _encode(self, x, cu, codes, *, want_tokens, views=None, token_rows=None, token_sel_of_row=None):
        se = self.stain_encoding_dim if codes is not None else 0
        return self.wsi_embedders.run_kernels(x, cu, codes, self._heads(), self.embedding.weight if se else None, se_dim=se,
                                              want_tokens=want_tokens, want_projector=True, want_ref_feats=False, views=views,
                                              precision=self.b200_precision, token_rows=token_rows,
                                              token_sel_of_row=token_sel_of_row)

    def _token_window(self, bs, n_tokens, data):
        """Resolved window W (0 = off → full [bs, T, 128] token embeddings, the reference's return shape)."""
        req = data.get("b200_token_window") if isinstance(data, dict) else None
        if req is None:
            req = self.b200_token_window
        if req is None:
            req = os.environ.get("MADELEINE_B200_TOKEN_WINDOW")
        if req is None or req is False or str(req).lower() in ("off", "0", "none", "false", ""):
            return 0
        if str(req).lower() in ("batch", "on", "true", "auto"):
            world = 1
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                world = torch.distributed.get_world_size()      # cases are sharded: the permutation spans the GLOBAL batch
            w = bs * world
        else:
            w = int(req)
            if w <= 0:
                raise ValueError(f"b200_token_window must be 'off', 'batch' or a positive integer (got {req!r})")
        return min(w, n_tokens)

    @staticmethod
    def _window_plan(lens, cu_host, window, device):
        """Token rows that token_projector must see.  lens [R] CPU int64: packed length of every bag (T, or 1 for a missing
        bag encoded from one token); cu_host [R+1].  Returns (token_rows [n_sel] int32, sel_of_row [M] int32 with -1 for
        rows outside the window, dense_idx [R*window] int64 or None when the compact order already is [R, window])."""
        R = lens.numel()
        k = lens.clamp(max=window)
        base = torch.zeros(R + 1, dtype=torch.int64)
        base[1:] = k.cumsum(0)
        n_sel, M = int(base[-1]), int(cu_host[-1])
        rows_host = torch.repeat_interleave(cu_host[:-1] - base[:-1], k) + torch.arange(n_sel)
        token_rows = rows_host.to(torch.int32).to(device, non_blocking=True)
        sel_of_row = torch.full((M,), -1, dtype=torch.int32, device=device)
        sel_of_row[token_rows.long()] = torch.arange(n_sel, dtype=torch.int32, device=device)
        dense_idx = None
        if n_sel != R * window:
            # a bag shorter than the window is a missing bag (one token, all of its T tokens are identical): repeat its row
            t = torch.arange(R * window) % window
            full = torch.repeat_interleave(k == window, window)
            dense_idx = (torch.repeat_interleave(base[:-1], window) + t * full).to(device, non_blocking=True)
        return token_rows, sel_of_row, dense_idx

    @staticmethod
    def _compact_plan(present, n_tokens, device, want_inv=True):
        """Row maps for encoding missing (all-zero) bags from a single token.  present: CPU bool [R].
        Returns (rows [M_c] gather index into the dense [R*T] rows, cu_seqlens [R+1], inv [R*T] map back to packed rows);
        built from two [R]-sized host tensors, no device sync."""
        R = present.numel()
        lens = torch.where(present, torch.tensor(n_tokens), torch.tensor(1))
        cu_host = torch.zeros(R + 1, dtype=torch.int64)
        cu_host[1:] = lens.cumsum(0)
        m_c = int(cu_host[-1])
        lens_d = lens.to(device, non_blocking=True)
        starts = (torch.arange(R) * n_tokens - cu_host[:-1]).to(device, non_blocking=True)
        rows = torch.arange(m_c, device=device) + torch.repeat_interleave(starts, lens_d, output_size=m_c)
        cu_dev = cu_host.to(torch.int32).to(device, non_blocking=True)
        if not want_inv:
            return rows, cu_dev, None
        t_in_bag = torch.arange(R * n_tokens, device=device) % n_tokens
        pres_rep = torch.repeat_interleave(present.to(device, non_blocking=True).long(), n_tokens, output_size=R * n_tokens)
        inv = torch.repeat_interleave(cu_host[:-1].to(device, non_blocking=True), n_tokens, output_size=R * n_tokens) + t_in_bag * pres_rep
        return rows, cu_dev, inv

    # -- reference API --------------------------------------------------------------------------------------------------
    def encode_he(self, feats, device):
        """[bs, T, D] → [bs, 512]; never adds stain encodings (Model.py:97-107, quirk Q2)."""
        feats = feats.to(device)
        bs, T, D = feats.shape
        if self.stain_encoding:
            # the reference feeds D channels into a Linear expecting D+32 and fails with a shape error
            raise RuntimeError("encode_he does not add stain encodings (Model.py:97-107); a stain-encoding model cannot use it")
        cu = ABMILEmbedder.uniform_cu(bs, T, feats.device)
        return self._encode(feats.reshape(bs * T, D), cu, None, want_tokens=False)["slide"].view(bs, -1)

    def encode_packed(self, feats, cu_seqlens, stain_codes=None):
        """Extension: bag-packed variable-length inference. feats [sum N_i, D], cu_seqlens int32 [R+1] → [R, 512]."""
        cu = torch.as_tensor(cu_seqlens, dtype=torch.int32, device=feats.device)
        codes = None
        if self.stain_encoding:
            if stain_codes is None:
                raise ValueError("stain_codes ([R] int) are required for a stain-encoding model")
            codes = torch.as_tensor(stain_codes, dtype=torch.int32, device=feats.device)
        return self._encode(feats, cu, codes, want_tokens=False)["slide"]

    def forward_packed(self, feats, cu_seqlens, stain_codes=None, want_tokens=True):
        """Extension: training forward on bag-packed ragged bags → (slide [R, 512], tokens [sum N_i, 128] or None)."""
        cu = torch.as_tensor(cu_seqlens, dtype=torch.int32, device=feats.device)
        codes = None
        if self.stain_encoding:
            if stain_codes is None:
                raise ValueError("stain_codes ([R] int) are required for a stain-encoding model")
            codes = torch.as_tensor(stain_codes, dtype=torch.int32, device=feats.device)
        out = self._encode(feats, cu, codes, want_tokens=want_tokens)
        return out["slide"], out["tokens"]

    def forward(self, data, device, train=True, n_views=1, custom_stain_idx=None, return_attention=False):
        all_wsi_feats = data["feats"].to(device)
        all_embeddings, all_token_embeddings = ops.EmbeddingDict(), {}

        if train:
            bs, n_mod, n_tokens, d_in = all_wsi_feats.shape
            R = bs * n_mod
            cu = ABMILEmbedder.uniform_cu(R, n_tokens, all_wsi_feats.device)
            codes = None
            if self.stain_encoding:
                # quirk Q1 (Model.py:126-129): flattened row r (slide r // n_mod, modality r % n_mod) receives code r // bs
                codes = (torch.arange(R, device=all_wsi_feats.device) // bs).to(torch.int32)
            views = ABMILEmbedder.half_views(R, n_tokens, all_wsi_feats.device) if n_views != 1 else None
            flat = all_wsi_feats.reshape(R * n_tokens, d_in)
            labels = data.get("modality_labels") if isinstance(data, dict) else None
            compact = None
            window = self._token_window(bs, n_tokens, data) if n_views == 1 else 0
            if (self.b200_skip_missing_bags and labels is not None and n_views == 1 and n_tokens > 1
                    and tuple(labels.shape) == (bs, n_mod)):
                present = labels.detach().to("cpu").reshape(R) != 0
                if not bool(present.all()):
                    compact = self._compact_plan(present, n_tokens, all_wsi_feats.device, want_inv=not window)
            if window:
                # token window (quirk Q3): token_projector sees the first `window` tokens of every bag only
                if compact is None:
                    lens = torch.full((R,), n_tokens, dtype=torch.int64)
                else:
                    lens = torch.where(present, torch.tensor(n_tokens), torch.tensor(1))
                cu_host = torch.zeros(R + 1, dtype=torch.int64)
                cu_host[1:] = lens.cumsum(0)
                token_rows, sel_of_row, dense_idx = self._window_plan(lens, cu_host, window, all_wsi_feats.device)
                if compact is None:
                    out = self._encode(flat, cu, codes, want_tokens=True, token_rows=token_rows, token_sel_of_row=sel_of_row)
                else:
                    rows, cu_c, _ = compact
                    out = self._encode(flat.index_select(0, rows), cu_c, codes, want_tokens=True, token_rows=token_rows,
                                       token_sel_of_row=sel_of_row)
                out = dict(out)
                if dense_idx is not None:
                    out["tokens"] = out["tokens"].index_select(0, dense_idx)
                n_tokens = window
            elif compact is None:
                out = self._encode(flat, cu, codes, want_tokens=True, views=views)
            else:
                rows, cu_c, inv = compact
                out = self._encode(flat.index_select(0, rows), cu_c, codes, want_tokens=True)
                out = dict(out)
                out["tokens"] = out["tokens"].index_select(0, inv)      # missing bags: their single token row, T times
            d_out = out["slide"].shape[-1]
            slide = out["slide"]
            if n_views == 1:
                slide_embeddings = slide.view(bs, n_mod, 1, d_out)
            else:
                whole, halves = slide[:R], slide[R:].view(2, R, d_out).transpose(0, 1)
                slide_embeddings = torch.cat([whole.unsqueeze(1), halves], dim=1).view(bs, n_mod, 3, d_out)
            token_embeddings = out["tokens"].view(bs, n_mod, n_tokens, -1)
            # the loss glue may address the slide embeddings by row of this matrix instead of through the per-modality views
            all_embeddings.b200_set(slide, bs, n_mod, 1 if n_views == 1 else 3)
            for idx, modality in enumerate(self.modalities):
                slide_emb = slide_embeddings[:, idx, :, :]
                token_emb = token_embeddings[:, idx, :]
                if modality == "HE":
                    # the reference materialises n_mod-1 copies (Model.py:153-155); a stride-0 view holds the same values
                    slide_emb = slide_emb.unsqueeze(dim=3).expand(-1, -1, -1, n_mod - 1)
                    token_emb = token_emb.unsqueeze(dim=3).expand(-1, -1, -1, n_mod - 1)
                all_embeddings[modality] = slide_emb
                all_token_embeddings[modality] = token_emb
            return all_embeddings, all_token_embeddings

        elif not train and not return_attention:
            bs, n_mod, n_tokens, d_in = all_wsi_feats.shape
            if n_mod != 1:
                # Model.py:174,196: the per-stain view(bs*n_mod, ...) only holds for n_mod == 1
                raise RuntimeError(f"eval forward expects one stain per call (n_mod == 1), got n_mod={n_mod}")
            cu = ABMILEmbedder.uniform_cu(bs, n_tokens, all_wsi_feats.device)
            for stain_idx in range(n_mod):
                stain_name = self.modalities[custom_stain_idx] if custom_stain_idx else self.modalities[stain_idx]
                codes = None
                if self.stain_encoding:
                    key = custom_stain_idx if custom_stain_idx else stain_idx
                    codes = torch.full((bs,), int(key), dtype=torch.int32, device=all_wsi_feats.device)
                feats = all_wsi_feats[:, stain_idx].reshape(bs * n_tokens, d_in)
                emb = self._encode(feats, cu, codes, want_tokens=False)["slide"]
                all_embeddings[stain_name] = emb.view(bs, n_mod, -1)
            return all_embeddings

        else:
            bs, n_mod, n_tokens, d_in = all_wsi_feats.shape
            if n_mod != 1:
                raise RuntimeError(f"return_attention expects n_mod == 1 (Model.py:211), got n_mod={n_mod}")
            if self.stain_encoding:
                raise RuntimeError("return_attention does not add stain encodings (Model.py:209)")
            cu = ABMILEmbedder.uniform_cu(bs, n_tokens, all_wsi_feats.device)
            out = self._encode(all_wsi_feats[:, HE_POSITION].reshape(bs * n_tokens, d_in), cu, None, want_tokens=False)
            H = self.wsi_embedders.n_heads
            return out["slide"].view(bs, n_mod, -1), out["logits"].view(bs, n_tokens, 1, H)
