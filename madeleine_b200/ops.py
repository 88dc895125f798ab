"""Host orchestration of the sm_100a kernels: weight packing, the encoder forward/backward sequence and the
autograd Functions that put them behind the reference's module API.

PyTorch is used here for device memory (torch.empty/zeros), streams and autograd bookkeeping only; every
arithmetic step of the hot path is one of the C-ABI kernels in libmadeleine_b200.so.  There is no CPU path.

Layouts
  tokens are bag-packed rows [M = sum N_i]; bag r owns rows [cu[r], cu[r+1]);
  activations that feed a GEMM are bf16 "planes" [nplanes, M, C] (hi[, lo]);
  the 2048-wide pre-attention features are kept head-major (c' = h*512 + e) — the reference's
  'b t (e c) -> b t e c' rearrange (Model.py:396) makes head h read channels h::4, which is a row permutation of
  pre_attn.8 / LayerNorm(2048) and a column permutation of token_projector / projector, applied at pack time.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from . import executor as _ex
from ._lib import call, stream_ptr

HID = 512          # wsi_encoder_hidden_dim the kernels are built for
GATE = 512         # BatchedABMIL hidden_dim (Model.py:71)
TOK = 128          # token_projector out (Model.py:80-83)
LN_EPS = 1e-5
ACT_CODES = {"softmax": 0, "leaky_relu": 1, "relu": 2, "sigmoid": 3}

_step_counter = [0]
PLANES_F16 = 0x100      # MDL_PLANES_F16: flag on nplanes / nsplit arguments selecting fp16 hi/lo planes (fp32-grade inference)

# Optional phase timeline (bench.py --timeline): when this is a list, the encoder's backward appends (label, cuda event)
# pairs at its phase boundaries so that the multi-GPU cost of the two gradient all-reduces can be read off the device clock.
timeline = None


def _mark(label):
    if timeline is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        timeline.append((label, ev))


def resolve_precision(requested: Optional[str]) -> str:
    """'fp32' (3-pass split-bf16 tcgen05, fp32-grade, forward and backward), 'bf16' (1 pass) or 'fp32_fwd' (fp32-grade
    forward — outputs and losses at the reference's fp32 values — with 1-pass bf16 GEMMs in the backward pass: gradients
    at the accuracy the reference's own bf16-autocast training gives, for a third of the backward tensor work).
    'auto' follows torch autocast, which is how the reference picks bf16 (--precision bfloat16 → torch.amp.autocast,
    trainer.py:108)."""
    req = (requested or os.environ.get("MADELEINE_B200_PRECISION", "auto")).lower()
    if req in ("fp32", "float32", "fp32x3", "bf16x3"):
        return "fp32"
    if req in ("fp32_fwd", "fp32-fwd", "mixed"):
        return "fp32_fwd"
    if req in ("bf16", "bfloat16"):
        return "bf16"
    if req != "auto":
        raise ValueError(f"unknown precision {requested!r}; use 'auto', 'fp32', 'fp32_fwd' or 'bf16'")
    if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") in (torch.bfloat16, torch.float16):
        return "bf16"
    return "fp32"


# --------------------------------------------------------------------------------------------------------------
# weight packing
# --------------------------------------------------------------------------------------------------------------
@dataclass
class _Seg:
    off: int
    shape: Tuple[int, ...]

    @property
    def numel(self):
        n = 1
        for s in self.shape:
            n *= s
        return n


class PackSpec:
    """Index maps from kernel-layout buffers into a flat 'master' concatenation of the module parameters."""

    def __init__(self, param_shapes: Sequence[Tuple[int, ...]], n_heads: int, d_in_total: int, device, d_in: Optional[int] = None):
        """d_in_total = columns of pre_attn.0.weight (patch features + stain-encoding channels); d_in = patch features alone
        (the GEMM operand; the stain columns act through a per-bag bias).  Defaults to d_in_total."""
        H = n_heads
        d_in = d_in_total if d_in is None else d_in
        C = HID * H
        self.n_heads, self.C, self.d_in_total = H, C, d_in_total
        offs, o = [], 0
        for shp in param_shapes:
            offs.append(o)
            n = 1
            for s in shp:
                n *= s
            o += n
        self.master_numel = o
        self.param_offsets = offs
        self.param_shapes = list(param_shapes)
        P = {name: i for i, name in enumerate(PARAM_ORDER(H))}
        off = lambda name: offs[P[name]]  # noqa: E731
        ar = torch.arange
        # head-major permutation: packed channel c' = h*HID + e  <-  reference channel e*H + h
        cp = ar(C)
        perm = (cp % HID) * H + (cp // HID)

        bf, f32, gr = [], [], []          # lists of (name, idx tensor, shape)

        def mat(base, rows_src, cols_src, ld):
            return base + rows_src[:, None] * ld + cols_src[None, :]

        k512 = ar(HID)
        w1 = mat(off("pre0.w"), ar(HID), ar(d_in), d_in_total)             # [512, d_in] (stain columns excluded)
        w2 = mat(off("pre4.w"), ar(HID), k512, HID)
        w3 = mat(off("pre8.w"), perm, k512, HID)                           # [2048, 512] head-major rows
        # gated weights: packed row p = h*1024 + g*256 + is_b*128 + i, gate column j = g*128 + i
        pr = ar(H * 1024)
        ph, pin = pr // 1024, pr % 1024
        pj = (pin // 256) * 128 + (pin % 128)
        pis_b = (pin % 256) // 128
        base_a = torch.tensor([off(f"a{h}.w") for h in range(H)])
        base_b = torch.tensor([off(f"b{h}.w") for h in range(H)])
        row_base = torch.where(pis_b.bool(), base_b[ph], base_a[ph]) + pj * HID
        wab = row_base[:, None] + k512[None, :]                              # [H*1024, 512]
        wabT = wab.view(H, 1024, HID).transpose(1, 2).reshape(H * HID, 1024)  # [H*512 (h,k), 1024 (packed n)]
        tp = mat(off("tp.w"), ar(TOK), perm, C)                             # [128, 2048] head-major columns
        for name, idx in (("w1", w1), ("w2", w2), ("w2T", w2.t()), ("w3", w3), ("w3T", w3.t()), ("wab", wab),
                          ("wabT", wabT), ("tp", tp), ("tpT", tp.t())):
            bf.append((name, idx.contiguous()))
        hj = ar(H * GATE)
        hh, jj = hj // GATE, hj % GATE
        vec = {
            "b1": off("pre0.b") + k512, "g1": off("ln1.w") + k512, "be1": off("ln1.b") + k512,
            "b2": off("pre4.b") + k512, "g2": off("ln5.w") + k512, "be2": off("ln5.b") + k512,
            "b3": off("pre8.b") + perm, "g3": off("ln9.w") + perm, "be3": off("ln9.b") + perm,
            "ba": torch.tensor([off(f"a{h}.b") for h in range(H)])[hh] + jj,
            "bb": torch.tensor([off(f"b{h}.b") for h in range(H)])[hh] + jj,
            "wc": torch.tensor([off(f"c{h}.w") for h in range(H)])[hh] + jj,
            "bc": torch.tensor([off(f"c{h}.b") for h in range(H)]),
            "btp": off("tp.b") + ar(TOK),
            "wp": mat(off("proj.w"), ar(HID), perm, C),                      # [512, 2048] head-major columns
            "bp": off("proj.b") + ar(HID),
        }
        for name, idx in vec.items():
            f32.append((name, idx.contiguous()))
        # gradient buffer: same layouts as the forward operands it mirrors
        for name, idx in (("w1", w1), ("w2", w2), ("w3", w3), ("wab", wab), ("tp", tp)):
            gr.append((name, idx.contiguous()))
        for name, idx in vec.items():
            gr.append((name, idx.contiguous()))

        def layout(items):
            segs, o2, flat = {}, 0, []
            for name, idx in items:
                n = idx.numel()
                n_pad = (n + 63) // 64 * 64           # keep every segment 256-byte aligned (TMA / float4)
                segs[name] = _Seg(o2, tuple(idx.shape))
                flat.append(idx.reshape(-1))
                if n_pad != n:
                    flat.append(idx.reshape(-1)[:1].expand(n_pad - n))
                o2 += n_pad
            return segs, o2, torch.cat(flat).to(torch.int32)

        self.bf_segs, self.bf_numel, bf_idx = layout(bf)
        self.f32_segs, self.f32_numel, f32_idx = layout(f32)
        self.gr_segs, self.gr_numel, gr_idx = layout(gr)
        assert self.master_numel < 2 ** 31
        self.bf_idx = bf_idx.to(device)
        self.f32_idx = f32_idx.to(device)
        # padded grad slots must not scatter: compact (position, master index) lists without padding
        pos, dst = [], []
        for name, idx in gr:
            s = self.gr_segs[name]
            pos.append(torch.arange(s.off, s.off + idx.numel()))
            dst.append(idx.reshape(-1))
        gr_pos_all, gr_dst_all = torch.cat(pos), torch.cat(dst)
        self.gr_pos = gr_pos_all.to(torch.int32).to(device)
        self.gr_dst = gr_dst_all.to(torch.int32).to(device)
        # Parameters whose gradients are complete once the third pre-attention layer's wgrad has been issued (pre_attn.8/9,
        # the attention heads, token_projector, projector) are one contiguous range of the master layout; under gradient
        # sync that range is all-reduced while the backward pass of the first two layers is still running.
        self.early_lo = offs[P["pre8.w"]]
        self.early_hi = offs[P["emb.w"]] if P["emb.w"] < len(offs) else self.master_numel
        early = (gr_dst_all >= self.early_lo) & (gr_dst_all < self.early_hi)
        self.gr_pos_early = gr_pos_all[early].to(torch.int32).to(device)
        self.gr_dst_early = gr_dst_all[early].to(torch.int32).to(device)
        self.gr_pos_late = gr_pos_all[~early].to(torch.int32).to(device)
        self.gr_dst_late = gr_dst_all[~early].to(torch.int32).to(device)
        self.off = off
        self.P = P
        self.param_numels = [(offs[i + 1] if i + 1 < len(offs) else self.master_numel) - offs[i] for i in range(len(offs))]
        self.ip_static = _ex.static_iparams(self)


def PARAM_ORDER(n_heads: int) -> List[str]:
    names = ["pre0.w", "pre0.b", "ln1.w", "ln1.b", "pre4.w", "pre4.b", "ln5.w", "ln5.b", "pre8.w", "pre8.b", "ln9.w", "ln9.b"]
    for h in range(n_heads):
        names += [f"a{h}.w", f"a{h}.b", f"b{h}.w", f"b{h}.b", f"c{h}.w", f"c{h}.b"]
    names += ["tp.w", "tp.b", "proj.w", "proj.b", "emb.w"]
    return names


class PackedWeights:
    """Kernel-layout copies of the parameters for one precision; rebuilt whenever a parameter changes.

    ``master`` is the flat fp32 concatenation of the parameters in PARAM_ORDER — normally the very buffer the parameters
    are views of (ABMILEmbedder flattens them once), so re-packing after an optimiser step is two launches and no copy."""

    def __init__(self, spec: PackSpec, master: torch.Tensor, nplanes: int, f16: bool = False):
        """f16: fp16 hi/lo operand planes (weights scaled by 64) — the inference format of the fp32-grade mode, see
        include/madeleine_b200.h MDL_PLANES_F16; training packs bf16 planes."""
        dev = master.device
        st = stream_ptr(dev)
        self.spec, self.nplanes, self.f16 = spec, nplanes, f16
        self.master = master
        self.bf = torch.empty(nplanes, spec.bf_numel, dtype=torch.float16 if f16 else torch.bfloat16, device=dev)
        call("mdl_gather_split", master, spec.bf_idx, spec.bf_numel, self.bf, spec.bf_numel, nplanes | (PLANES_F16 if f16 else 0), st)
        self.f32 = torch.empty(spec.f32_numel, dtype=torch.float32, device=dev)
        call("mdl_gather_f32", master, spec.f32_idx, spec.f32_numel, self.f32, st)

    def planes(self, name):
        """(tensor view of plane 0 start, rows, cols, plane_stride)."""
        s = self.spec.bf_segs[name]
        return self.bf[0, s.off:], s.shape[0], s.shape[1], self.spec.bf_numel

    def vec(self, name):
        s = self.spec.f32_segs[name]
        return self.f32[s.off:s.off + s.numel]


# --------------------------------------------------------------------------------------------------------------
# thin kernel wrappers (tests, tools and the few callers outside the executor)
# --------------------------------------------------------------------------------------------------------------
def _planes_empty(nplanes, M, C, dev):
    return torch.empty(nplanes, M, C, dtype=torch.bfloat16, device=dev)


def split_planes(x: torch.Tensor, nplanes: int) -> torch.Tensor:
    M, C = x.shape
    out = _planes_empty(nplanes, M, C, x.device)
    call("mdl_split_planes", x, M, C, x.stride(0), out, M * C, nplanes, stream_ptr(x.device))
    return out


def gemm_nt(a: torch.Tensor, K: int, bw: Tuple, out_cols: int, nsplit: int, bias=None, rowbias=None, row2bag=None,
            grp_n_cols: int = 0, a_koff: int = 0, out: Optional[torch.Tensor] = None, out_dtype=torch.float32) -> torch.Tensor:
    """out[M, N] = a[:, :K(+offsets)] @ B^T (+bias). a: planes [np, M, Ca]; bw = (ptr tensor, rows, cols, plane_stride).
    out_dtype torch.bfloat16: the fp32 accumulator is rounded to bf16 on the way out (bf16 mode: what autocast's Linear returns)."""
    npl, M, Ca = a.shape
    bt, b_rows, b_cols, b_ps = bw
    N = out_cols
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    call("mdl_gemm_nt", a, M, Ca, Ca, M * Ca, bt, b_rows, b_cols, b_cols, b_ps, out, out.stride(0), M, N, K, nsplit,
         grp_n_cols, a_koff, bias, rowbias, row2bag, int(out.dtype == torch.bfloat16), stream_ptr(a.device))
    return out


def gemm_tn_accum(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, nsplit: int, grp_m_rows: int = 0, b_coff: int = 0):
    """out[Mo, No] += a^T b over tokens. a: planes [np, T, Ca] (Mo <= Ca), b: planes [np, T, Cb]."""
    npl, T, Ca = a.shape
    _, Tb, Cb = b.shape
    assert T == Tb
    Mo, No = out.shape
    call("mdl_gemm_tn_accum", a, Ca, Ca, T * Ca, b, Cb, Cb, T * Cb, T, out, out.stride(0), Mo, No, nsplit, grp_m_rows, b_coff, 0,
         stream_ptr(a.device))
    return out


def gate_buffer(M: int, HC: int, device) -> torch.Tensor:
    """fp16 scratch for the saved gates of mdl_gemm_gated / mdl_gate_bwd (tiled layout, whole 32-row blocks)."""
    return torch.empty((M + 31) // 32 * 32 * HC, dtype=torch.float16, device=device)


def gate_untile(buf: torch.Tensor, M: int, HC: int) -> torch.Tensor:
    """Tiled gate scratch -> row-major [M, HC] (tests / debugging): blocks of 32 rows x 8 columns are contiguous."""
    n_rb = (M + 31) // 32
    return buf.view(n_rb, HC // 16, 2, 32, 8).permute(0, 3, 1, 2, 4).reshape(n_rb * 32, HC)[:M]


def gate_tile(x: torch.Tensor) -> torch.Tensor:
    """Row-major [M, HC] fp16 -> tiled gate scratch (inverse of gate_untile; padding rows are zero)."""
    M, HC = x.shape
    n_rb = (M + 31) // 32
    pad = torch.zeros(n_rb * 32, HC, dtype=x.dtype, device=x.device)
    pad[:M] = x
    return pad.view(n_rb, 32, HC // 16, 2, 8).permute(0, 2, 3, 1, 4).contiguous().view(-1)


def pool_fwd(h3, npl, logits, cu, tok_idx, R, total_tokens, H, E, out, attn_p, act, tsplit=0):
    """Attention pooling; picks the token split and provides the (deterministic) partial-sum workspace."""
    M, C = h3.shape[1], h3.shape[2]
    if attn_p is None:
        attn_p = torch.empty(M, H, dtype=torch.float32, device=h3.device)
    if tsplit <= 0:
        tsplit = call("mdl_pool_tsplit", R, total_tokens, H, E)
    ws = None
    if tsplit > 1:
        ws = torch.empty(call("mdl_pool_workspace_bytes", R, H, E, tsplit), dtype=torch.uint8, device=h3.device)
    st = stream_ptr(h3.device)
    call("mdl_pool_weights", logits, cu, tok_idx, R, H, E, attn_p, act, tsplit, ws, st)
    call("mdl_pool_fwd", h3, M * C, npl, attn_p, cu, tok_idx, R, total_tokens, H, E, out, tsplit, ws, st)
    return out


def ln_gelu_fwd(z, gamma, beta, nplanes, drop_p, seed, stream_id):
    M, C = z.shape
    planes = _planes_empty(nplanes, M, C, z.device)
    mean = torch.empty(M, dtype=torch.float32, device=z.device)
    rstd = torch.empty(M, dtype=torch.float32, device=z.device)
    call("mdl_ln_gelu_fwd", z, M, C, gamma, beta, LN_EPS, drop_p, seed, stream_id, planes, M * C, nplanes, mean, rstd,
         int(z.dtype == torch.bfloat16), stream_ptr(z.device))
    return planes, mean, rstd


def ln_gelu_bwd(z, gamma, beta, mean, rstd, dh_a, dh_b, pool_terms, n_heads, nplanes, drop_p, seed, stream_id, dgamma, dbeta, dbias,
                dh_b_rows=None, row2bag=None, bag_dz=None):
    """dh_b_rows: dh_b is compact [n_sel, C] and dh_b_rows [M] int32 maps token → compact row (-1: none).
    bag_dz [n_bags, C] (zero-filled) + row2bag: also accumulate per-bag column sums of dz (first layer only)."""
    M, C = z.shape
    dz = _planes_empty(nplanes, M, C, z.device)
    t = list(pool_terms) + [(None, None, None)] * (2 - len(pool_terms))
    in_bf16 = z.dtype == torch.bfloat16              # z, dh_a and dh_b share one storage type
    for g_in in (dh_a, dh_b):
        if g_in is not None and g_in.dtype != z.dtype:
            raise RuntimeError(f"ln_gelu_bwd: upstream gradient dtype {g_in.dtype} != activation dtype {z.dtype}")
    call("mdl_ln_gelu_bwd", z, M, C, gamma, beta, mean, rstd, dh_a, dh_b, dh_b_rows, t[0][0], t[0][1], t[0][2], t[1][0], t[1][1], t[1][2],
         n_heads, drop_p, seed, stream_id, dz, M * C, nplanes, dgamma, dbeta, dbias, row2bag, bag_dz, int(in_bf16), stream_ptr(z.device))
    return dz


# --------------------------------------------------------------------------------------------------------------
# encoder forward / backward: one native call each (csrc/executor.cu)
# --------------------------------------------------------------------------------------------------------------
@dataclass
class EncodeOptions:
    n_heads: int = 4
    activation: str = "softmax"
    precision: str = "fp32"          # resolved: 'fp32' | 'bf16' | 'fp32_fwd'
    training: bool = False           # nn.Module.training → dropout active
    want_tokens: bool = False        # fused token_projector → tokens [M, 128]
    want_projector: bool = False     # fused projector → [R, 512] instead of pooled [R, 512, H]
    want_ref_feats: bool = False     # pre-attention features in reference order [M, 512, H] (ABMILEmbedder API)
    d_in: int = 512
    se_dim: int = 0
    views: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None  # n_views=3: (tok_idx, cu2, row2seg2)
    seed: int = 0
    # token window: token_projector only sees the packed rows `token_rows` [n_sel] (unique, int32); `token_sel_of_row` [M]
    # maps a packed row to its position in that list or -1.  tokens are then returned compact, [n_sel, 128].
    token_rows: Optional[torch.Tensor] = None
    token_sel_of_row: Optional[torch.Tensor] = None


def _nsplit(precision):
    """(GEMM passes, operand planes) of the forward pass."""
    return (3, 2) if precision in ("fp32", "fp32_fwd") else (1, 1)


def _act_dtype(precision):
    """Storage type of the pre-LayerNorm activations and of the dgrad outputs."""
    return torch.bfloat16 if precision == "bf16" else torch.float32


def _nsplit_bwd(precision):
    """(GEMM passes, planes of the gradient operands) of the backward pass."""
    return (3, 2) if precision == "fp32" else (1, 1)


class _Saved:
    pass


_arena_bytes_cache: Dict[Tuple, int] = {}


def _arena_bytes(which: str, ip: List[int], ipa) -> int:
    """Arena size for this shape/flag combination (asked from the library once per combination)."""
    I = _ex.I
    key = (which, ip[I['M']], ip[I['R']], ip[I['D_IN']], ip[I['SE_DIM']], ip[I['NPL_FWD']], ip[I['NPL_BWD']], ip[I['ACT_BF16']],
           ip[I['KEEP']], ip[I['WANT_TOKENS']], ip[I['N_VIEW_TOK']], ip[I['R2']], ip[I['N_SEL']], ip[I['GR_NUMEL']])
    n = _arena_bytes_cache.get(key)
    if n is None:
        n = call("mdl_encoder_fwd_arena_bytes" if which == "fwd" else "mdl_encoder_bwd_arena_bytes", ipa)
        if len(_arena_bytes_cache) > 4096:
            _arena_bytes_cache.clear()
        _arena_bytes_cache[key] = n
    return n


def encoder_forward(x: torch.Tensor, cu: torch.Tensor, codes: Optional[torch.Tensor], pw: PackedWeights, opt: EncodeOptions,
                    keep_for_backward: bool):
    """pre_attn → gated attention → pooling (→ projector, token_projector) in ONE native call.  Returns (outputs dict, saved)."""
    dev = x.device
    I = _ex.I
    spec = pw.spec
    nsplit, npl = _nsplit(opt.precision)
    nsplit_b, npl_b = _nsplit_bwd(opt.precision)
    H = opt.n_heads
    C = HID * H
    M = x.shape[0]
    R = cu.numel() - 1
    p_pre = 0.1 if opt.training else 0.0      # nn.Dropout(0.1) x3, Model.py:354,358,362
    p_gate = 0.25 if opt.training else 0.0    # nn.Dropout(0.25) on both gates, abmil.py:33-35
    ip = list(spec.ip_static)
    ip[I['M']], ip[I['R']], ip[I['D_IN']], ip[I['SE_DIM']] = M, R, opt.d_in, opt.se_dim
    ip[I['NSPLIT_FWD']], ip[I['NPL_FWD']], ip[I['NSPLIT_BWD']], ip[I['NPL_BWD']] = nsplit, npl, nsplit_b, npl_b
    ip[I['ACT_BF16']] = int(opt.precision == "bf16")
    ip[I['ACTIVATION']] = ACT_CODES[opt.activation]
    ip[I['KEEP']] = int(keep_for_backward)
    ip[I['WANT_TOKENS']], ip[I['WANT_PROJECTOR']], ip[I['WANT_REF']] = int(opt.want_tokens), int(opt.want_projector), int(opt.want_ref_feats)
    ip[I['SEED']] = opt.seed & 0x7FFFFFFFFFFFFFFF
    if pw.f16:
        if keep_for_backward or opt.precision == "bf16":
            raise RuntimeError("madeleine_b200: fp16 operand planes are the fp32-grade inference format (no backward state)")
        ip[I['PLANES_F16']] = 1
    R2 = 0
    tok_idx = cu2 = row2seg2 = None
    if opt.views is not None:
        tok_idx, cu2, row2seg2 = opt.views
        R2 = cu2.numel() - 1
        ip[I['N_VIEW_TOK']], ip[I['R2']] = tok_idx.numel(), R2
    n_sel = 0
    if opt.want_tokens and opt.token_rows is not None:
        n_sel = opt.token_rows.numel()
        ip[I['N_SEL']] = n_sel
    ipa = _ex.iarr(ip)
    fpa = _ex.farr(p_pre, p_gate)
    arena = torch.empty(_arena_bytes("fwd", ip, ipa), dtype=torch.uint8, device=dev)
    f32 = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)  # noqa: E731
    n_slide = R + R2
    logits = f32(M, H)
    slide_hm = f32(n_slide, C)
    slide = f32(n_slide, HID) if opt.want_projector else None
    tokens = f32(n_sel if n_sel else M, TOK) if opt.want_tokens else None
    ref = f32(M, HID, H) if opt.want_ref_feats else None
    ptrs = {"STREAM": stream_ptr(dev), "X": x, "CU": cu, "CODES": codes, "MASTER": pw.master, "WBF": pw.bf, "WF32": pw.f32, "ARENA": arena,
            "SLIDE_HM": slide_hm, "SLIDE": slide, "LOGITS": logits, "TOKENS": tokens, "REF": ref, "VIEW_TOK_IDX": tok_idx, "VIEW_CU": cu2,
            "VIEW_ROW2SEG": row2seg2, "TOKEN_ROWS": opt.token_rows if n_sel else None,
            "TOKEN_SEL_OF_ROW": opt.token_sel_of_row if n_sel else None}
    call("mdl_encoder_fwd", ipa, fpa, _ex.parr(ptrs))
    outs = {"logits": logits, "tokens": tokens, "ref_feats": ref}
    if opt.want_projector:
        outs["slide"] = slide
    else:
        # reference layout [*, E, H] (c = e*H + h) from head-major [*, H, E]
        outs["slide"] = slide_hm.view(-1, H, HID).transpose(1, 2).contiguous()
    sv = None
    if keep_for_backward:
        sv = _Saved()
        sv.opt, sv.pw, sv.M, sv.R, sv.n_slide = opt, pw, M, R, n_slide
        sv.ip, sv.fpa, sv.ptrs = ip, fpa, ptrs          # ptrs keeps every tensor the backward pass reads alive
    return outs, sv


def encoder_backward(sv, d_slide: Optional[torch.Tensor], d_logits: Optional[torch.Tensor], d_tokens: Optional[torch.Tensor],
                     d_ref_feats: Optional[torch.Tensor], early_sync=None) -> torch.Tensor:
    """Returns the flat master-layout gradient of all parameters (ONE native call; two when ``early_sync`` is given).

    ``early_sync(gmaster, lo, hi)`` (optional) is called as soon as ``gmaster[lo:hi]`` — everything except the first two
    pre-attention layers and the stain embedding — is final, while the rest of the backward pass is still to be issued."""
    opt, pw = sv.opt, sv.pw
    spec = pw.spec
    I = _ex.I
    dev = pw.master.device
    H = opt.n_heads
    C = HID * H
    M, n_slide = sv.M, sv.n_slide
    if d_slide is not None:
        if opt.want_projector:
            d_slide = d_slide.contiguous().float()
        else:
            d_slide = d_slide.float().reshape(n_slide, HID, H).transpose(1, 2).contiguous().view(n_slide, C)
    if d_logits is not None:
        d_logits = d_logits.reshape(M, H).contiguous().float()
    if d_tokens is not None:
        d_tokens = d_tokens.reshape(-1, TOK).contiguous().float()
    d_ref_hm = None
    if d_ref_feats is not None:
        # gradient w.r.t. the reference-order features → head-major, added as a second dh source
        d_ref_hm = d_ref_feats.float().reshape(M, HID, H).transpose(1, 2).contiguous().view(M, C).to(_act_dtype(opt.precision))
    ip = list(sv.ip)
    ipa = _ex.iarr(ip)
    barena = torch.empty(_arena_bytes("bwd", ip, ipa), dtype=torch.uint8, device=dev)
    gmaster = torch.empty(spec.master_numel, dtype=torch.float32, device=dev)
    ptrs = dict(sv.ptrs)
    ptrs.update({"STREAM": stream_ptr(dev), "BWD_ARENA": barena, "D_SLIDE": d_slide, "D_LOGITS": d_logits, "D_TOKENS": d_tokens,
                 "D_REF_HM": d_ref_hm, "GMASTER": gmaster, "GR_POS": spec.gr_pos, "GR_DST": spec.gr_dst,
                 "GR_POS_EARLY": spec.gr_pos_early, "GR_DST_EARLY": spec.gr_dst_early, "GR_POS_LATE": spec.gr_pos_late,
                 "GR_DST_LATE": spec.gr_dst_late})
    pp = _ex.parr(ptrs)
    if early_sync is None:
        call("mdl_encoder_bwd", ipa, sv.fpa, pp)
    else:
        ip[I['PHASE']] = 1
        call("mdl_encoder_bwd", _ex.iarr(ip), sv.fpa, pp)
        early_sync(gmaster, spec.early_lo, spec.early_hi)
        ip[I['PHASE']] = 2
        call("mdl_encoder_bwd", _ex.iarr(ip), sv.fpa, pp)
    return gmaster


class EncodeFn(torch.autograd.Function):
    """autograd wrapper: (x, *params) → (slide, logits, tokens?, ref_feats?)."""

    @staticmethod
    def forward(ctx, holder, x, *params):
        opt, cu, codes, pw = holder["opt"], holder["cu"], holder["codes"], holder["pw"]
        need_grad = holder["need_grad"]
        outs, sv = encoder_forward(x, cu, codes, pw, opt, keep_for_backward=need_grad)
        ctx.sv = sv if need_grad else None
        ctx.n_params = len(params)
        ctx.param_meta = [(p.shape, p.requires_grad) for p in params]
        ret = [outs["slide"], outs["logits"]]
        ret.append(outs.get("tokens"))
        ret.append(outs.get("ref_feats"))
        ctx.set_materialize_grads(False)   # unused outputs arrive as None, not as dense zero tensors
        return tuple(ret)

    @staticmethod
    def backward(ctx, d_slide, d_logits, d_tokens, d_ref):
        sv = ctx.sv
        if sv is None:
            raise RuntimeError("madeleine_b200: backward called on a forward that ran without grad state")
        from . import parallel
        spec = sv.pw.spec
        if parallel.gradient_sync_enabled():
            # 90 % of the 20 MB gradient (everything but the first two layers) is all-reduced asynchronously while those two
            # layers' backward kernels run; the remainder follows at the end
            import torch.distributed as dist
            pending = []

            def early(gm, lo, hi):
                pending.append(dist.all_reduce(gm[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

            _mark("bwd_begin")
            gmaster = encoder_backward(sv, d_slide, d_logits, d_tokens, d_ref, early_sync=early)
            _mark("bwd_kernels_done")
            # the gradients that only became final now (first two layers, stain embedding: 2 MB) — latency-bound, nothing
            # left to hide them under: one peer-memory kernel each instead of NCCL's protocol
            if spec.early_lo > 0:
                parallel.small_all_reduce_(gmaster[:spec.early_lo])
            if spec.early_hi < spec.master_numel:
                parallel.small_all_reduce_(gmaster[spec.early_hi:])
            _mark("late_allreduce_done")
            for work in pending:
                work.wait()
            _mark("early_allreduce_joined")
        else:
            gmaster = encoder_backward(sv, d_slide, d_logits, d_tokens, d_ref)
        # parameters that no incoming gradient can reach keep grad = None, as in the reference (AdamW then skips them: no
        # weight decay / moment updates on e.g. token_projector when no local loss is used)
        opt = sv.opt
        untouched = set()
        if d_tokens is None or not opt.want_tokens:
            untouched |= {"tp.w", "tp.b"}
        if d_slide is None or not opt.want_projector:
            untouched |= {"proj.w", "proj.b"}
        skip = {spec.P[n] for n in untouched}
        grads = []
        for i, (shape, req) in enumerate(ctx.param_meta):
            if not req or i in skip:
                grads.append(None)
                continue
            o = spec.param_offsets[i]
            grads.append(gmaster[o:o + spec.param_numels[i]].view(shape))
        ctx.sv = None
        return (None, None, *grads)


# --------------------------------------------------------------------------------------------------------------
# InfoNCE
# --------------------------------------------------------------------------------------------------------------
_RED = {"none": 0, "mean": 1, "sum": 2}


class InfoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, temperature, symmetric, reduction):
        m, D = q.shape
        dev = q.device
        q = q.contiguous().float()
        k = k.contiguous().float()
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)  # noqa: E731
        qn, kn, L, lse_r, lse_c, nll_r, nll_c = f(m), f(m), f(m, m), f(2 * m), f(2 * m), f(m), f(m)
        loss = f(())
        red = _RED[reduction]
        call("mdl_infonce_fwd", q, k, m, D, float(temperature), int(bool(symmetric)), red, qn, kn, L, lse_r, lse_c, nll_r, nll_c,
             loss if red else None, stream_ptr(dev))
        ctx.save_for_backward(q, k, qn, kn, L, lse_r, lse_c)
        ctx.cfg = (float(temperature), bool(symmetric), reduction, m, D)
        if red == 0:
            return 0.5 * nll_r + 0.5 * nll_c if symmetric else nll_r
        return loss

    @staticmethod
    def backward(ctx, go):
        q, k, qn, kn, L, lse_r, lse_c = ctx.saved_tensors
        temperature, symmetric, reduction, m, D = ctx.cfg
        dev = q.device
        go = go.float()
        if reduction == "none":
            w = go.reshape(m)
        else:
            w = go.reshape(1).expand(m) * (1.0 / m if reduction == "mean" else 1.0)
        w_r = (0.5 * w if symmetric else w).contiguous()
        w_c = w_r if symmetric else None
        G = torch.empty(m, m, dtype=torch.float32, device=dev)
        dq = torch.empty_like(q)
        dk = torch.empty_like(k)
        call("mdl_infonce_bwd", q, k, m, D, temperature, qn, kn, L, lse_r, lse_c, w_r, w_c, G, dq, dk, stream_ptr(dev))
        return dq, dk, None, None, None


def info_nce(query, positive_key, temperature=0.1, reduction="mean", symmetric=False):
    _lib.require_cuda(query, "InfoNCE query")
    _lib.require_cuda(positive_key, "InfoNCE positive_key")
    with torch.cuda.device(query.device):
        return InfoNCEFn.apply(query, positive_key, temperature, symmetric, reduction)


class EmbeddingDict(dict):
    """The ``{modality: tensor}`` dict MADELEINE.forward(train=True) returns (Model.py:147-159) plus a handle on the matrix
    all of its entries are views of, so that the loss glue can hand row INDEX LISTS to the kernels instead of materialising
    boolean-mask selections (trainer.py:28-33).

    ``b200_base`` [n_rows, 512]: the encoder's slide embeddings; ``b200_row(case, modality, view)`` -> row of that matrix
    (works on LongTensors / ints)."""

    b200_base: Optional[torch.Tensor] = None

    def b200_set(self, base, bs, n_mod, n_views, world=1):
        self.b200_base, self.b200_bs, self.b200_n_mod, self.b200_n_views, self.b200_world = base, bs, n_mod, n_views, world
        return self

    def b200_row(self, case, modality, view):
        # encoder output order: [whole views of all (case, modality) rows | half view 1 | half view 2]; under case sharding
        # every rank contributes one such block, rank-major (parallel.all_gather_rows)
        R = self.b200_bs * self.b200_n_mod
        rank, local = case // self.b200_bs, case % self.b200_bs
        return rank * (R * self.b200_n_views) + view * R + local * self.b200_n_mod + modality


class InfoNCERowsFn(torch.autograd.Function):
    """Sum of mean-reduced (optionally symmetric) InfoNCE terms whose operands are ROWS of one matrix; one autograd node for
    all stains / views of a step, no gathered copies, no index_put in backward.  ``plan``: dict(idx=int32 device tensor with
    all row lists back to back, terms=[(q_offset, k_offset, m, temperature, symmetric), ...])."""

    @staticmethod
    def forward(ctx, base, plan):
        dev = base.device
        base_c = base.contiguous().float()
        n_rows, D = base_c.shape
        st = stream_ptr(dev)
        terms, idx = plan["terms"], plan["idx"]
        sizes = [int(call("mdl_infonce_rows_workspace_floats", m)) for _, _, m, _, _ in terms]
        ws = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        out = torch.zeros(len(terms) + 1, dtype=torch.float32, device=dev)          # [total | per-term losses]
        o = 0
        offs = []
        for t, ((qo, ko, m, tau, sym), n) in enumerate(zip(terms, sizes)):
            call("mdl_infonce_rows_fwd", base_c, D, idx[qo:], idx[ko:], m, D, float(tau), int(bool(sym)), ws[o:], out[t + 1:], out, st)
            offs.append(o)
            o += n
        ctx.save_for_backward(base_c, idx, ws)
        ctx.meta = (terms, offs)
        ctx.per_term = out[1:]
        return out[0]

    @staticmethod
    def backward(ctx, go):
        base_c, idx, ws = ctx.saved_tensors
        terms, offs = ctx.meta
        dev = base_c.device
        D = base_c.shape[1]
        go = go.contiguous().float()
        dbase = torch.zeros_like(base_c)
        st = stream_ptr(dev)
        for (qo, ko, m, tau, sym), o in zip(terms, offs):
            call("mdl_infonce_rows_bwd", base_c, D, idx[qo:], idx[ko:], m, D, float(tau), int(bool(sym)), ws[o:], go, dbase, st)
        return dbase, None


def info_nce_rows(base: torch.Tensor, row_lists, temperatures, symmetric):
    """row_lists: [(q_rows, k_rows)] CPU int tensors/lists of equal length per pair; returns the SUM of the terms' losses."""
    _lib.require_cuda(base, "slide embeddings")
    flat, terms, o = [], [], 0
    for (q_rows, k_rows), tau, sym in zip(row_lists, temperatures, symmetric):
        q_rows = torch.as_tensor(q_rows, dtype=torch.int32)
        k_rows = torch.as_tensor(k_rows, dtype=torch.int32)
        m = q_rows.numel()
        terms.append((o, o + m, m, tau, sym))
        flat += [q_rows, k_rows]
        o += 2 * m
    idx = torch.cat(flat).pin_memory().to(base.device, non_blocking=True)          # ONE small upload for all terms
    with torch.cuda.device(base.device):
        return InfoNCERowsFn.apply(base, {"idx": idx, "terms": terms})


# --------------------------------------------------------------------------------------------------------------
# Graph-OT local loss
# --------------------------------------------------------------------------------------------------------------
_got_streams = {}


def _side_stream(device, slot):
    key = (str(device), slot)
    st = _got_streams.get(key)
    if st is None:
        st = torch.cuda.Stream(device)
        _got_streams[key] = st
    return st


class GOTFn(torch.autograd.Function):
    """Graph-OT loss; forward computes the loss AND d loss / d tokens in the same launches (backward only scales).

    ``slot >= 0`` runs the kernels on a side stream (one per slot) so that independent per-stain problems — each only
    <= 65 CTAs — overlap on the 148 SMs; the caller must then call ``got_join`` before consuming the result."""

    @staticmethod
    def forward(ctx, v, q, slot):
        m, n, D = v.shape
        dev = v.device
        v = v.contiguous().float()
        q = q.contiguous().float()
        max_n = call("mdl_got_max_tokens")
        if n > max_n:
            raise RuntimeError(f"madeleine_b200 GOT kernel supports at most {max_n} tokens per problem (got {n}); "
                               "the reference's subsampling quirk (loss.py:281-284) bounds n by the number of cases")
        cur = torch.cuda.current_stream(dev)
        side = _side_stream(dev, slot) if slot >= 0 else None
        if side is not None:
            side.wait_stream(cur)
        with torch.cuda.stream(side if side is not None else cur):
            ws_bytes = call("mdl_got_workspace_bytes", m, n, D)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            extrema = torch.empty(6, dtype=torch.float32, device=dev)
            st = stream_ptr(dev)
            call("mdl_got_extrema", v, q, m, n, D, ws, extrema, st)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            wd = torch.empty(m, dtype=torch.float32, device=dev)
            gwd = torch.empty(m, dtype=torch.float32, device=dev)
            dv = torch.empty_like(v)
            dq = torch.empty_like(q)
            call("mdl_got_fwd_bwd", v, q, m, n, D, ws, extrema, loss, wd, gwd, dv, dq, st)
        if side is not None:
            for t in (v, q):
                t.record_stream(side)           # inputs were produced on `cur`, consumed on `side`
            for t in (loss, dv, dq, wd, gwd):
                t.record_stream(cur)            # outputs were allocated on `side`, consumed on `cur` after got_join
        ctx.save_for_backward(dv, dq)
        ctx.parts = (wd, gwd)
        return loss

    @staticmethod
    def backward(ctx, go):
        dv, dq = ctx.saved_tensors
        return dv * go, dq * go, None


class GOTShardedFn(torch.autograd.Function):
    """GOT over this rank's share of the (case, stain) problems when cases are sharded across ranks.

    The batch-wide min/max of the three cost tensors (quirk Q5) are all-reduced between the cost kernel and the main
    kernel, and the three threshold-gradient sums between the main kernel and the gradient kernel; token embeddings
    never leave their rank.  Every rank must call this for every stain, even with zero local problems.  Returns the sum
    of the LOCAL problems' losses (the global loss is the sum over ranks)."""

    @staticmethod
    def forward(ctx, v, q):
        import torch.distributed as dist
        m, n, D = v.shape
        dev = v.device
        st = stream_ptr(dev)
        neg = torch.tensor([-1.0, 1.0, -1.0, 1.0, -1.0, 1.0], device=dev)
        if m > 0:
            v = v.contiguous().float()
            q = q.contiguous().float()
            max_n = call("mdl_got_max_tokens")
            if n > max_n:
                raise RuntimeError(f"madeleine_b200 GOT kernel supports at most {max_n} tokens per problem (got {n})")
            ws = torch.empty(call("mdl_got_workspace_bytes", m, n, D), dtype=torch.uint8, device=dev)
            extrema = torch.empty(6, dtype=torch.float32, device=dev)
            call("mdl_got_extrema", v, q, m, n, D, ws, extrema, st)
            packed = extrema * neg                       # (-min, max) pairs -> one MAX all-reduce
        else:
            packed = torch.full((6,), float("-inf"), device=dev)
        dist.all_reduce(packed, op=dist.ReduceOp.MAX)
        extrema_g = (packed * neg).contiguous()
        dthr = torch.zeros(3, dtype=torch.float32, device=dev)
        if m > 0:
            wd = torch.empty(m, dtype=torch.float32, device=dev)
            gwd = torch.empty(m, dtype=torch.float32, device=dev)
            call("mdl_got_main", m, n, D, ws, extrema_g, wd, gwd, dthr, st)
        dist.all_reduce(dthr, op=dist.ReduceOp.SUM)
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        if m > 0:
            dv = torch.empty_like(v)
            dq = torch.empty_like(q)
            call("mdl_got_finish", v, q, m, n, D, ws, extrema_g, dthr, wd, gwd, loss, dv, dq, st)
            ctx.save_for_backward(dv, dq)
        ctx.has_local = m > 0
        return loss

    @staticmethod
    def backward(ctx, go):
        if not ctx.has_local:
            return None, None
        dv, dq = ctx.saved_tensors
        return dv * go, dq * go


def got_loss(v, q, slot: int = -1):
    _lib.require_cuda(v, "GOT tokens")
    with torch.cuda.device(v.device):
        return GOTFn.apply(v, q, slot)


def got_join(device, slots):
    """Make the current stream wait for the GOT problems issued on the given side-stream slots."""
    cur = torch.cuda.current_stream(device)
    for slot in slots:
        cur.wait_stream(_side_stream(device, slot))
