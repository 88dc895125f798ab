"""CPU-side checks: the C-ABI library loads, exports every symbol include/madeleine_b200.h declares, the ctypes
table matches the header, and the host-side weight packing index maps are correct (no GPU compute)."""
import ctypes
import os
import re

import pytest
import torch

from madeleine_b200 import _lib, ops

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "madeleine_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdl_\w+)\s*\(", text)))


def test_header_and_ctypes_table_agree():
    assert sorted(_lib.exported_symbols()) == header_symbols()


def test_library_loads_and_exports_every_symbol():
    lib = _lib.load()
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} missing from {_lib.LIB_PATH}"
    assert lib.mdl_built_arch() == 100
    assert isinstance(lib.mdl_last_error(), bytes)


def test_argument_counts_match_header():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, argtypes in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", text, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert n == len(argtypes), f"{name}: header has {n} args, ctypes table {len(argtypes)}"


def test_cpu_tensor_rejected_without_fallback():
    from madeleine_b200.utils.loss import InfoNCE
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        InfoNCE(temperature=0.1)(torch.zeros(4, 8), torch.zeros(4, 8))
    with pytest.raises(ValueError):
        InfoNCE()(torch.zeros(4, 8, 2), torch.zeros(4, 8))
    assert InfoNCE()(torch.zeros(4, 8), torch.zeros(4, 8), negative_keys=torch.zeros(3, 8)) is None  # reference quirk


def _params(n_heads=4, se=True, n_mod=3):
    from weights import make_state_dict
    sd = make_state_dict(3, n_mod=n_mod, stain_encoding=se)
    names = ["wsi_embedders.pre_attn.0", "wsi_embedders.pre_attn.1", "wsi_embedders.pre_attn.4", "wsi_embedders.pre_attn.5",
             "wsi_embedders.pre_attn.8", "wsi_embedders.pre_attn.9"]
    ps = []
    for n in names:
        ps += [sd[n + ".weight"], sd[n + ".bias"]]
    for h in range(n_heads):
        for part in ("attention_a.0", "attention_b.0", "attention_c"):
            ps += [sd[f"wsi_embedders.attn.{h}.{part}.weight"], sd[f"wsi_embedders.attn.{h}.{part}.bias"]]
    ps += [sd["token_projector.weight"], sd["token_projector.bias"], sd["projector.weight"], sd["projector.bias"]]
    if se:
        ps.append(sd["embedding.weight"])
    return sd, ps


def test_pack_spec_index_maps():
    sd, ps = _params()
    spec = ops.PackSpec([tuple(p.shape) for p in ps], 4, 544, "cpu", d_in=512)
    master = torch.cat([p.reshape(-1) for p in ps])
    assert spec.master_numel == master.numel()

    def bf(name):
        s = spec.bf_segs[name]
        return master[spec.bf_idx[s.off:s.off + s.numel].long()].view(s.shape)

    def f32(name):
        s = spec.f32_segs[name]
        return master[spec.f32_idx[s.off:s.off + s.numel].long()].view(s.shape)

    W1, W2, W3 = (sd[f"wsi_embedders.pre_attn.{i}.weight"] for i in (0, 4, 8))
    assert torch.equal(bf("w1"), W1[:, :512])
    assert torch.equal(bf("w2"), W2) and torch.equal(bf("w2T"), W2.t())
    # head-major rows: packed row h*512+e <- reference row e*4+h  (einops 'b t (e c) -> b t e c', head = c)
    w3p = W3.view(512, 4, 512).permute(1, 0, 2).reshape(2048, 512)
    assert torch.equal(bf("w3"), w3p) and torch.equal(bf("w3T"), w3p.t())
    g3 = sd["wsi_embedders.pre_attn.9.weight"].view(512, 4).t().reshape(-1)
    assert torch.equal(f32("g3"), g3)
    wab = bf("wab").view(4, 4, 2, 128, 512)
    for h in range(4):
        Wa = sd[f"wsi_embedders.attn.{h}.attention_a.0.weight"]
        Wb = sd[f"wsi_embedders.attn.{h}.attention_b.0.weight"]
        assert torch.equal(wab[h, :, 0].reshape(512, 512), Wa)
        assert torch.equal(wab[h, :, 1].reshape(512, 512), Wb)
        assert torch.equal(bf("wabT")[h * 512:(h + 1) * 512], bf("wab")[h * 1024:(h + 1) * 1024].t())
        assert torch.equal(f32("wc")[h * 512:(h + 1) * 512], sd[f"wsi_embedders.attn.{h}.attention_c.weight"][0])
        assert f32("bc")[h] == sd[f"wsi_embedders.attn.{h}.attention_c.bias"][0]
    tp = sd["token_projector.weight"].view(128, 512, 4).permute(0, 2, 1).reshape(128, 2048)
    assert torch.equal(bf("tp"), tp) and torch.equal(bf("tpT"), tp.t())
    wp = sd["projector.weight"].view(512, 512, 4).permute(0, 2, 1).reshape(512, 2048)
    assert torch.equal(f32("wp"), wp)
    # the gradient scatter list is injective and covers every parameter except W1's stain columns and the embedding
    dst = spec.gr_dst.long()
    assert dst.unique().numel() == dst.numel()
    covered = torch.zeros(spec.master_numel, dtype=torch.bool)
    covered[dst] = True
    expect = torch.ones(spec.master_numel, dtype=torch.bool)
    o = spec.off("pre0.w")
    w1mask = torch.zeros(512, 544, dtype=torch.bool)
    w1mask[:, 512:] = True
    expect[o:o + 512 * 544] = ~w1mask.reshape(-1)
    e = spec.off("emb.w")
    expect[e:e + 3 * 32] = False
    assert torch.equal(covered, expect)


@pytest.mark.parametrize("se", [True, False])
def test_pack_spec_early_late_gradient_split(se):
    """Under gradient sync the master-layout gradient is all-reduced in two parts (ops.EncodeFn.backward): the range
    [early_lo, early_hi) — pre_attn.8/9, the heads, token_projector, projector — as soon as the third layer's wgrad is issued,
    the rest at the end.  The two scatter lists must partition the packed gradient exactly and respect that range."""
    sd, ps = _params(se=se)
    d_total = 544 if any(tuple(p.shape) == (512, 544) for p in ps) else 512
    spec = ops.PackSpec([tuple(p.shape) for p in ps], 4, d_total, "cpu", d_in=512)
    names = ops.PARAM_ORDER(4)
    offs = dict(zip(names, spec.param_offsets))
    assert spec.early_lo == offs["pre8.w"]
    assert spec.early_hi == (offs["emb.w"] if "emb.w" in offs and len(spec.param_offsets) == len(names) else spec.master_numel)
    e_dst, l_dst = spec.gr_dst_early.long(), spec.gr_dst_late.long()
    assert int(e_dst.min()) >= spec.early_lo and int(e_dst.max()) < spec.early_hi
    assert bool(((l_dst < spec.early_lo) | (l_dst >= spec.early_hi)).all())
    both = torch.cat([spec.gr_pos_early, spec.gr_pos_late]).sort().values
    assert torch.equal(both, spec.gr_pos.sort().values)
    assert torch.equal(torch.cat([e_dst, l_dst]).sort().values, spec.gr_dst.long().sort().values)
    # everything inside the early range except nothing is missing: every master element of those parameters is written once
    covered = torch.zeros(spec.master_numel, dtype=torch.int32)
    covered[spec.gr_dst.long()] += 1
    stain_cols = 0 if d_total == 512 else 512 * 32
    emb = spec.master_numel - spec.early_hi
    assert int((covered == 0).sum()) == stain_cols + emb          # only the stain columns of W1 and the embedding table come later
    assert int(covered.max()) == 1


# ---------------------------------------------------------------------------------------------- native step executor
def _header_enum(name):
    text = open(HEADER).read()
    body = re.search(r"enum %s \{(.*?)\};" % name, text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return [n.split("=")[0].strip() for n in body.split(",") if n.strip()]


def test_executor_enums_mirror_the_header():
    from madeleine_b200 import executor as ex
    assert _header_enum("mdl_enc_i") == ["MDL_ENC_I_" + n for n in ex.ENC_I] + ["MDL_ENC_I_COUNT"]
    assert _header_enum("mdl_enc_f") == ["MDL_ENC_F_" + n for n in ex.ENC_F] + ["MDL_ENC_F_COUNT"]
    assert _header_enum("mdl_enc_p") == ["MDL_ENC_P_" + n for n in ex.ENC_P] + ["MDL_ENC_P_COUNT"]
    assert len(_header_enum("mdl_prof_tag")) - 1 == len(ex.PROF_TAGS) == len(_lib._PROF_NAMES)
    assert _lib.load().mdl_encoder_abi() == ex.ABI


def test_executor_arena_plan_is_a_dry_run_of_the_launch_sequence():
    """mdl_encoder_{fwd,bwd}_arena_bytes walk the same code as the real pass with a null arena (no GPU needed): sizes grow
    with the token count, keep-for-backward costs the two fp16 gate buffers, bf16 mode halves the activation planes, and
    the backward scratch covers the 16 KB/token gate-gradient planes."""
    from madeleine_b200 import executor as ex
    _, ps = _params()
    spec = ops.PackSpec([tuple(p.shape) for p in ps], 4, 544, "cpu", d_in=512)
    I = ex.I

    def ip(M, R, keep=1, npl=2, want_tokens=1, n_sel=0, act_bf16=0):
        v = list(spec.ip_static)
        v[I["M"]], v[I["R"]], v[I["D_IN"]], v[I["SE_DIM"]] = M, R, 512, 32
        v[I["NSPLIT_FWD"]], v[I["NPL_FWD"]], v[I["NSPLIT_BWD"]], v[I["NPL_BWD"]] = 3 if npl == 2 else 1, npl, 3 if npl == 2 else 1, npl
        v[I["KEEP"]], v[I["WANT_TOKENS"]], v[I["WANT_PROJECTOR"]], v[I["N_SEL"]], v[I["ACT_BF16"]] = keep, want_tokens, 1, n_sel, act_bf16
        return ex.iarr(v)

    fwd = lambda *a, **k: _lib.call("mdl_encoder_fwd_arena_bytes", ip(*a, **k))  # noqa: E731
    bwd = lambda *a, **k: _lib.call("mdl_encoder_bwd_arena_bytes", ip(*a, **k))  # noqa: E731
    M = 64000
    full = fwd(M, 32)
    # per token, fp32-grade: x planes 2 KB, z1/z2 2 KB each, h1/h2 2 KB each, z3 8 KB, h3 8 KB, gates 2 x 4 KB, + small
    per_token = 2048 + 2 * 2048 + 2 * 2048 + 8192 + 8192 + 2 * 4096
    assert per_token * M <= full <= (per_token + 256) * M
    assert fwd(M, 32, keep=0) == pytest.approx(full - 2 * 4096 * M, rel=1e-3)
    assert fwd(2 * M, 32) > 1.9 * full
    assert fwd(M, 32, npl=1, act_bf16=1) < 0.65 * full          # the fp16 gate buffers do not shrink
    assert fwd(M, 32, n_sel=1024) - full == pytest.approx(2 * 1024 * 2048 * 2, rel=1e-2)     # gathered rows for the token window
    b = bwd(M, 32)
    assert 16384 * M + 8192 * M <= b <= 56 * 1024 * M
    assert bwd(M, 32, n_sel=1024) <= b                                                        # compact token_projector gradient
