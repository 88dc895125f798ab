"""Per-kernel parity on the GPU, each kernel called through the C ABI (madeleine_b200._lib.call).

The comparator is plain torch fp32/fp64 math on the same device inputs (torch is the checker here, never the path)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from madeleine_b200 import ops  # noqa: E402
from madeleine_b200._lib import call, stream_ptr  # noqa: E402

DEV = "cuda"


def planes_f32(p):
    return p.float().sum(0)


def rel_err(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _st():
    return stream_ptr(torch.device(DEV))


# ---------------------------------------------------------------------------------------------- elementwise
@pytest.mark.parametrize("npl", [1, 2])
def test_split_planes(npl):
    x = torch.randn(257, 512, device=DEV) * 3
    p = ops.split_planes(x, npl)
    hi = x.to(torch.bfloat16)
    assert torch.equal(p[0], hi)
    if npl == 2:
        assert torch.equal(p[1], (x - hi.float()).to(torch.bfloat16))
        assert rel_err(planes_f32(p), x) < 2e-5


def test_gather_scatter_row2bag():
    src = torch.randn(1000, device=DEV)
    idx = torch.randperm(1000, device=DEV)[:640].to(torch.int32)
    out = torch.empty(640, device=DEV)
    call("mdl_gather_f32", src, idx, 640, out, _st())
    assert torch.equal(out, src[idx.long()])
    dst = torch.zeros(1000, device=DEV)
    call("mdl_scatter_f32", out, idx, 640, dst, 0, _st())
    assert torch.equal(dst[idx.long()], out)
    pl = torch.empty(2, 640, dtype=torch.bfloat16, device=DEV)
    call("mdl_gather_split", src, idx, 640, pl, 640, 2, _st())
    assert rel_err(planes_f32(pl), src[idx.long()]) < 2e-5
    cu = torch.tensor([0, 5, 5, 40, 100], dtype=torch.int32, device=DEV)
    r2b = torch.empty(100, dtype=torch.int32, device=DEV)
    call("mdl_row2bag", cu, 4, r2b, 100, _st())
    ref = torch.repeat_interleave(torch.arange(4, device=DEV), torch.tensor([5, 0, 35, 60], device=DEV))
    assert torch.equal(r2b.long(), ref)


def _ln_gelu_ref(z, g, b):
    y = torch.nn.functional.layer_norm(z, (z.shape[1],), g, b, 1e-5)
    return torch.nn.functional.gelu(y)


@pytest.mark.parametrize("C", [512, 2048])
@pytest.mark.parametrize("M", [1, 37, 1000])
def test_ln_gelu_fwd(C, M):
    z = torch.randn(M, C, device=DEV) * 2 + 0.3
    g = 1 + 0.1 * torch.randn(C, device=DEV)
    b = 0.1 * torch.randn(C, device=DEV)
    planes, mean, rstd = ops.ln_gelu_fwd(z, g, b, 2, 0.0, 0, 1)
    ref = _ln_gelu_ref(z, g, b)
    torch.testing.assert_close(planes_f32(planes), ref, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(mean, z.mean(1), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rstd, 1 / torch.sqrt(z.var(1, unbiased=False) + 1e-5), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("C,H", [(512, 1), (2048, 4)])
def test_ln_gelu_bwd(C, H):
    M, R = 333, 3
    z = (torch.randn(M, C, device=DEV) * 2).requires_grad_(True)
    g = (1 + 0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    b = (0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    dh_a = torch.randn(M, C, device=DEV)
    dh_b = torch.randn(M, C, device=DEV)
    p = torch.rand(M, H, device=DEV)
    dS = torch.randn(R, C, device=DEV)
    seg = torch.randint(0, R, (M,), device=DEV, dtype=torch.int32)
    dh = dh_a + dh_b + (p.repeat_interleave(C // H, dim=1) * dS[seg.long()])
    out = _ln_gelu_ref(z, g, b)
    out.backward(dh)
    _, mean, rstd = ops.ln_gelu_fwd(z.detach(), g.detach(), b.detach(), 2, 0.0, 0, 1)
    dg, db, dbias = (torch.zeros(C, device=DEV) for _ in range(3))
    dz = ops.ln_gelu_bwd(z.detach(), g.detach(), b.detach(), mean, rstd, dh_a, dh_b, [(p, dS, seg)], H, 2, 0.0, 0, 1, dg, db, dbias)
    scale = float(z.grad.abs().max())
    torch.testing.assert_close(planes_f32(dz), z.grad, rtol=1e-3, atol=1e-4 * scale)
    torch.testing.assert_close(dg, g.grad, rtol=1e-3, atol=1e-3 * float(g.grad.abs().max()))
    torch.testing.assert_close(db, b.grad, rtol=1e-3, atol=1e-3 * float(b.grad.abs().max()))
    torch.testing.assert_close(dbias, z.grad.sum(0), rtol=1e-3, atol=1e-3 * float(z.grad.sum(0).abs().max()) + 1e-4)


def test_ln_gelu_bwd_row_indexed_second_gradient():
    """dh_b given compactly for a subset of rows (token window) == dense dh_b that is zero elsewhere."""
    M, C, H, R = 2500, 2048, 4, 5
    z = torch.randn(M, C, device=DEV) * 2
    g = 1 + 0.1 * torch.randn(C, device=DEV)
    b = 0.1 * torch.randn(C, device=DEV)
    dh_a = torch.randn(M, C, device=DEV)
    rows = torch.randperm(M, device=DEV)[:300].sort().values.to(torch.int32)
    dh_sel = torch.randn(rows.numel(), C, device=DEV)
    dense = torch.zeros(M, C, device=DEV)
    dense[rows.long()] = dh_sel
    sel_of_row = torch.full((M,), -1, dtype=torch.int32, device=DEV)
    sel_of_row[rows.long()] = torch.arange(rows.numel(), dtype=torch.int32, device=DEV)
    p = torch.rand(M, H, device=DEV)
    dS = torch.randn(R, C, device=DEV)
    seg = torch.randint(0, R, (M,), device=DEV, dtype=torch.int32)
    _, mean, rstd = ops.ln_gelu_fwd(z, g, b, 2, 0.0, 0, 1)
    res = []
    for kw, dh_b in (({"dh_b_rows": sel_of_row}, dh_sel), ({}, dense)):
        acc = [torch.zeros(C, device=DEV) for _ in range(3)]
        dz = ops.ln_gelu_bwd(z, g, b, mean, rstd, dh_a, dh_b, [(p, dS, seg)], H, 2, 0.1, 99, 3, *acc, **kw)
        res.append((planes_f32(dz), acc))
    assert torch.equal(res[0][0], res[1][0])
    for a0, a1 in zip(res[0][1], res[1][1]):
        torch.testing.assert_close(a0, a1, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("lens", [[3, 10, 1, 20, 5, 8, 2], [2048, 1, 1, 2048, 700, 1], [5000]])
def test_ln_gelu_bwd_per_bag_sums(lens):
    """Per-bag column sums of dz accumulated inside the first layer's backward (stain-encoding gradient)."""
    R = len(lens)
    cu, M = _ragged(lens)
    C = 512
    z = torch.randn(M, C, device=DEV) * 2
    g = 1 + 0.1 * torch.randn(C, device=DEV)
    b = 0.1 * torch.randn(C, device=DEV)
    dh = torch.randn(M, C, device=DEV)
    _, mean, rstd = ops.ln_gelu_fwd(z, g, b, 2, 0.0, 0, 1)
    row2bag = torch.empty(M, dtype=torch.int32, device=DEV)
    call("mdl_row2bag", cu, R, row2bag, M, _st())
    acc = [torch.zeros(C, device=DEV) for _ in range(3)]
    G = torch.zeros(R, C, device=DEV)
    dz = planes_f32(ops.ln_gelu_bwd(z, g, b, mean, rstd, dh, None, [], 1, 2, 0.0, 0, 1, *acc, row2bag=row2bag, bag_dz=G))
    acc2 = [torch.zeros(C, device=DEV) for _ in range(3)]
    dz2 = planes_f32(ops.ln_gelu_bwd(z, g, b, mean, rstd, dh, None, [], 1, 2, 0.0, 0, 1, *acc2))
    assert torch.equal(dz, dz2)
    ref = torch.stack([dz[int(cu[i]):int(cu[i + 1])].sum(0) for i in range(R)])
    torch.testing.assert_close(G, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()) + 1e-5)
    torch.testing.assert_close(acc[2], G.sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("C", [512, 2048])
def test_ln_gelu_bf16_activations_equal_fp32_kernels_on_rounded_inputs(C):
    """bf16 mode stores the pre-LayerNorm activations and the dgrad outputs as bf16: the kernels must give exactly what the
    fp32-input kernels give on the same (bf16-representable) values."""
    M, H, R = 777, 4 if C == 2048 else 1, 3
    z16 = (torch.randn(M, C, device=DEV) * 2).bfloat16()
    g = 1 + 0.1 * torch.randn(C, device=DEV)
    b = 0.1 * torch.randn(C, device=DEV)
    p16, mean16, rstd16 = ops.ln_gelu_fwd(z16, g, b, 1, 0.1, 77, 2)
    p32, mean32, rstd32 = ops.ln_gelu_fwd(z16.float(), g, b, 1, 0.1, 77, 2)
    assert torch.equal(p16, p32) and torch.equal(mean16, mean32) and torch.equal(rstd16, rstd32)
    dh_a = torch.randn(M, C, device=DEV).bfloat16()
    dh_b = torch.randn(M, C, device=DEV).bfloat16()
    p = torch.rand(M, H, device=DEV)
    dS = torch.randn(R, C, device=DEV)
    seg = torch.randint(0, R, (M,), device=DEV, dtype=torch.int32)
    res = []
    for cast in (lambda t: t, lambda t: t.float()):
        acc = [torch.zeros(C, device=DEV) for _ in range(3)]
        dz = ops.ln_gelu_bwd(cast(z16), g, b, mean16, rstd16, cast(dh_a), cast(dh_b), [(p, dS, seg)], H, 1, 0.1, 77, 2, *acc)
        res.append((dz, acc))
    assert torch.equal(res[0][0], res[1][0])
    for a0, a1 in zip(res[0][1], res[1][1]):
        torch.testing.assert_close(a0, a1, rtol=1e-4, atol=1e-3)
    with pytest.raises(RuntimeError, match="dtype"):
        ops.ln_gelu_bwd(z16, g, b, mean16, rstd16, dh_a.float(), None, [], H, 1, 0.0, 0, 1, *[torch.zeros(C, device=DEV) for _ in range(3)])


def test_gather_rows_planes():
    M, C = 777, 2048
    x = torch.randn(M, C, device=DEV)
    xp = ops.split_planes(x, 2)
    rows = torch.tensor([0, 5, 5, 776, 13, 400], dtype=torch.int32, device=DEV)
    out = torch.empty(2, rows.numel(), C, dtype=torch.bfloat16, device=DEV)
    call("mdl_gather_rows_planes", xp, M * C, 2, C, rows, rows.numel(), out, rows.numel() * C, _st())
    assert torch.equal(out, xp[:, rows.long()])


def test_ln_gelu_dropout_consistency():
    """Dropout masks are regenerated in backward from (seed, stream, index): kept fraction ~ 1-p and fwd/bwd agree."""
    M, C, p = 256, 512, 0.1
    z = torch.randn(M, C, device=DEV)
    g = torch.ones(C, device=DEV)
    b = torch.zeros(C, device=DEV)
    h, mean, rstd = ops.ln_gelu_fwd(z, g, b, 2, p, 1234, 7)
    h0, _, _ = ops.ln_gelu_fwd(z, g, b, 2, 0.0, 1234, 7)
    h, h0 = planes_f32(h), planes_f32(h0)
    kept = h != 0
    frac = float(kept.float().mean())
    assert abs(frac - (1 - p)) < 0.01
    torch.testing.assert_close(h[kept], h0[kept] / (1 - p), rtol=1e-4, atol=1e-5)
    dh = torch.ones(M, C, device=DEV)
    acc = [torch.zeros(C, device=DEV) for _ in range(3)]
    dz = planes_f32(ops.ln_gelu_bwd(z, g, b, mean, rstd, dh, None, [], 1, 2, p, 1234, 7, *acc))
    # reference with the recovered mask
    zz = z.clone().requires_grad_(True)
    (_ln_gelu_ref(zz, g, b) * kept.float() / (1 - p)).sum().backward()
    torch.testing.assert_close(dz, zz.grad, rtol=1e-3, atol=1e-4)


# ---------------------------------------------------------------------------------------------- pooling
def _ragged(lens):
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=DEV)
    return cu, int(sum(lens))


@pytest.mark.parametrize("npl", [1, 2])
@pytest.mark.parametrize("lens", [[7], [200, 1, 63, 500], [2000] * 3])
@pytest.mark.parametrize("tsplit", [0, 1, 3])
def test_pool_fwd_bwd(npl, lens, tsplit):
    H, E = 4, 512
    C = H * E
    cu, M = _ragged(lens)
    R = len(lens)
    x = torch.randn(M, C, device=DEV)
    xp = ops.split_planes(x, npl)
    xr = planes_f32(xp).requires_grad_(True)
    logits = (torch.randn(M, H, device=DEV) * 3).requires_grad_(True)
    out = torch.empty(R, C, device=DEV)
    attn = torch.empty(M, H, device=DEV)
    ops.pool_fwd(xp, npl, logits.detach(), cu, None, R, M, H, E, out, attn, 0, tsplit)
    out_again = torch.empty_like(out)
    ops.pool_fwd(xp, npl, logits.detach(), cu, None, R, M, H, E, out_again, None, 0, tsplit)
    assert torch.equal(out, out_again)            # split partial sums are combined in a fixed order
    ref_rows, ref_p = [], []
    o = 0
    for n in lens:
        p = torch.softmax(logits[o:o + n], dim=0)                      # [n, H]
        ref_p.append(p)
        ref_rows.append((xr[o:o + n].view(n, H, E) * p[:, :, None]).sum(0).reshape(C))
        o += n
    ref = torch.stack(ref_rows)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(attn, torch.cat(ref_p).detach(), rtol=1e-4, atol=1e-7)
    dS = torch.randn(R, C, device=DEV)
    ref.backward(dS)
    dlogit = torch.empty(M, H, device=DEV)
    call("mdl_pool_bwd_dlogit", xp, M * C, npl, dS, out, attn, cu, None, R, M, H, E, dlogit, 0, logits.detach(), 0, tsplit, _st())
    torch.testing.assert_close(dlogit, logits.grad, rtol=1e-3, atol=1e-4 * float(logits.grad.abs().max()) + 1e-6)
    # the dX term is fused into ln_gelu_bwd: check p * dS against autograd's dX
    r2b = torch.repeat_interleave(torch.arange(R, device=DEV), torch.tensor(lens, device=DEV))
    dx = attn.repeat_interleave(E, dim=1) * dS[r2b]
    torch.testing.assert_close(dx, xr.grad, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("act,fn", [(1, torch.nn.functional.leaky_relu), (2, torch.relu), (3, torch.sigmoid)])
def test_pool_other_activations(act, fn):
    H, E, lens = 4, 512, [50, 130]
    C = H * E
    cu, M = _ragged(lens)
    x = torch.randn(M, C, device=DEV)
    xp = ops.split_planes(x, 2)
    logits = torch.randn(M, H, device=DEV).requires_grad_(True)
    out = torch.empty(2, C, device=DEV)
    attn = torch.empty(M, H, device=DEV)
    ops.pool_fwd(xp, 2, logits.detach(), cu, None, 2, M, H, E, out, attn, act)
    w = fn(logits)
    xr = planes_f32(xp)
    ref = torch.stack([(xr[a:b].view(b - a, H, E) * w[a:b, :, None]).sum(0).reshape(C) for a, b in ((0, 50), (50, 180))])
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)
    dS = torch.randn(2, C, device=DEV)
    ref.backward(dS)
    dlogit = torch.empty(M, H, device=DEV)
    call("mdl_pool_bwd_dlogit", xp, M * C, 2, dS, out, attn, cu, None, 2, M, H, E, dlogit, 0, logits.detach(), act, 0, _st())
    torch.testing.assert_close(dlogit, logits.grad, rtol=1e-3, atol=1e-3)


def test_pool_gather_views():
    """tok_idx segments (n_views=3 half views) pool a gathered subset with its own softmax."""
    H, E, T, R = 4, 512, 64, 2
    C = H * E
    x = torch.randn(R * T, C, device=DEV)
    xp = ops.split_planes(x, 2)
    logits = torch.randn(R * T, H, device=DEV)
    perm = torch.randperm(T)
    halves = [perm[:T // 2], perm[T // 2:]]
    idx = torch.cat([h + r * T for h in halves for r in range(R)]).to(torch.int32).to(DEV)
    cu2 = torch.arange(0, (2 * R + 1) * (T // 2), T // 2, dtype=torch.int32, device=DEV)
    out = torch.empty(2 * R, C, device=DEV)
    ops.pool_fwd(xp, 2, logits, cu2, idx, 2 * R, idx.numel(), H, E, out, None, 0)
    xr = planes_f32(xp)
    s = 0
    for v, h in enumerate(halves):
        for r in range(R):
            rows = (h + r * T).to(DEV)
            p = torch.softmax(logits[rows], dim=0)
            ref = (xr[rows].view(-1, H, E) * p[:, :, None]).sum(0).reshape(C)
            torch.testing.assert_close(out[s], ref, rtol=1e-4, atol=1e-5)
            s += 1


def test_planes_to_ref_order():
    M, H, E = 50, 4, 512
    x = torch.randn(M, H * E, device=DEV)
    xp = ops.split_planes(x, 2)
    out = torch.empty(M, E, H, device=DEV)
    call("mdl_planes_to_ref_order", xp, M * H * E, 2, M, H, E, out, _st())
    torch.testing.assert_close(out, planes_f32(xp).view(M, H, E).transpose(1, 2), rtol=0, atol=0)


# ---------------------------------------------------------------------------------------------- skinny / stain
@pytest.mark.parametrize("R,C,O", [(1, 2048, 512), (32, 2048, 512), (77, 2048, 512), (325, 2048, 512), (1500, 2048, 512),
                                   (5, 512, 512), (40, 2048, 200), (9, 512, 24), (33, 2048, 8)])
def test_skinny_linear(R, C, O):
    X = torch.randn(R, C, device=DEV, requires_grad=True)
    W = (torch.randn(O, C, device=DEV) / math.sqrt(C)).requires_grad_(True)
    b = torch.randn(O, device=DEV, requires_grad=True)
    Y = torch.empty(R, O, device=DEV)
    call("mdl_skinny_linear_fwd", X.detach(), W.detach(), b.detach(), R, C, O, Y, _st())
    ref = torch.nn.functional.linear(X.double(), W.double(), b.double())
    torch.testing.assert_close(Y.double(), ref, rtol=1e-5, atol=1e-5)
    dY = torch.randn(R, O, device=DEV)
    ref.backward(dY.double())
    dX = torch.empty(R, C, device=DEV)
    dW = torch.zeros(O, C, device=DEV)
    db = torch.zeros(O, device=DEV)
    call("mdl_skinny_linear_bwd", dY, X.detach(), W.detach(), R, C, O, dX, dW, db, _st())
    torch.testing.assert_close(dX, X.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dW, W.grad, rtol=1e-4, atol=1e-5 * max(1.0, R ** 0.5))
    torch.testing.assert_close(db, b.grad, rtol=1e-4, atol=1e-5 * max(1.0, R ** 0.5))


def test_stain_rowbias_fwd_bwd():
    R, n_mod, se, d_in, n_out = 7, 3, 32, 512, 512
    emb = torch.randn(n_mod, se, device=DEV, requires_grad=True)
    w1 = (torch.randn(n_out, d_in + se, device=DEV) * 0.05).requires_grad_(True)
    code = torch.randint(0, n_mod, (R,), device=DEV, dtype=torch.int32)
    rb = torch.empty(R, n_out, device=DEV)
    call("mdl_stain_rowbias", emb.detach(), code, w1.detach(), d_in + se, d_in, se, n_out, R, rb, _st())
    ref = emb[code.long()] @ w1[:, d_in:].t()
    torch.testing.assert_close(rb, ref, rtol=1e-5, atol=1e-6)
    G = torch.randn(R, n_out, device=DEV)
    ref.backward(G)
    dw1 = torch.zeros_like(w1)
    demb = torch.zeros_like(emb)
    call("mdl_stain_rowbias_bwd", G, emb.detach(), code, w1.detach(), d_in + se, d_in, se, n_out, R, dw1, demb, _st())
    torch.testing.assert_close(dw1, w1.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(demb, emb.grad, rtol=1e-4, atol=1e-5)
    # per-bag column sums of planes
    lens = [3, 10, 1, 20, 5, 8, 2]
    cu, M = _ragged(lens)
    x = torch.randn(M, 512, device=DEV)
    xp = ops.split_planes(x, 2)
    out = torch.empty(R, 512, device=DEV)
    call("mdl_bag_colsum_planes", xp, M * 512, 2, 512, cu, R, out, _st())
    xr = planes_f32(xp)
    ref2 = torch.stack([xr[int(cu[i]):int(cu[i + 1])].sum(0) for i in range(R)])
    torch.testing.assert_close(out, ref2, rtol=1e-5, atol=1e-5)
    cs = torch.zeros(512, device=DEV)
    call("mdl_colsum_f32", x, M, 512, cs, _st())
    torch.testing.assert_close(cs, x.sum(0), rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------------------------------------- gate backward
def test_gate_bwd():
    M, H = 100, 4
    pre_a = torch.randn(M, H * 512, device=DEV, requires_grad=True)
    pre_b = torch.randn(M, H * 512, device=DEV, requires_grad=True)
    wc = torch.randn(H * 512, device=DEV, requires_grad=True)
    a, b = torch.tanh(pre_a), torch.sigmoid(pre_b)
    ga, gb = a.detach().half(), b.detach().half()
    # autograd reference built on the fp16-rounded gates so only the kernel math is compared
    a16 = (ga.float() + (a - a.detach()))
    b16 = (gb.float() + (b - b.detach()))
    logit = (a16 * b16 * wc).view(M, H, 512).sum(-1)
    dlogit = torch.randn(M, H, device=DEV)
    logit.backward(dlogit)
    dpre = torch.empty(2, M, H * 1024, dtype=torch.bfloat16, device=DEV)
    dba, dbb, dwc = (torch.zeros(H * 512, device=DEV) for _ in range(3))
    dbc = torch.zeros(H, device=DEV)
    call("mdl_gate_bwd", ops.gate_tile(ga), ops.gate_tile(gb), dlogit, wc.detach(), M, H, 0.0, 0, dpre, M * H * 1024, 2, dba, dbb, dwc, dbc,
         _st())
    d = planes_f32(dpre).view(M, H, 4, 2, 128)               # [m, h, group, a|b, i]
    d_a = d[:, :, :, 0].reshape(M, H * 512)
    d_b = d[:, :, :, 1].reshape(M, H * 512)
    # the kernel differentiates tanh/sigmoid at the fp16-rounded gate values
    ref_a = (dlogit.repeat_interleave(512, 1) * wc.detach() * gb.float() * (1 - ga.float() ** 2))
    ref_b = (dlogit.repeat_interleave(512, 1) * wc.detach() * ga.float() * gb.float() * (1 - gb.float()))
    torch.testing.assert_close(d_a, ref_a, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(d_b, ref_b, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dba, ref_a.sum(0), rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dbb, ref_b.sum(0), rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dwc, wc.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dbc, dlogit.sum(0), rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------------------------------------- InfoNCE vs golden
def test_infonce_golden(golden):
    for c in golden("infonce")["cases"]:
        q = c["q"].to(DEV).requires_grad_(True)
        k = c["k"].to(DEV).requires_grad_(True)
        loss = ops.info_nce(q, k, temperature=c["tau"], reduction="mean", symmetric=c["symmetric"])
        loss.backward()
        torch.testing.assert_close(loss.cpu(), c["loss"], rtol=1e-3, atol=1e-3)
        for got, ref in ((q.grad.cpu(), c["dq"]), (k.grad.cpu(), c["dk"])):
            torch.testing.assert_close(got, ref, rtol=1e-2, atol=1e-2 * float(ref.abs().max()) + 1e-6)


@pytest.mark.parametrize("m", [65, 512, 1031])
@pytest.mark.parametrize("symmetric", [True, False])
def test_infonce_large_batches_against_fp64(m, symmetric):
    """Global batches of 8 x 64 cases and more (sharded runs gather every rank's slide embeddings): loss and gradients
    against an fp64 evaluation of loss.py:111-127."""
    g = torch.Generator().manual_seed(m)
    q = torch.randn(m, 512, generator=g)
    k = q + 2.0 * torch.randn(m, 512, generator=g)      # cos ~ 0.45: a loss of order one, well-conditioned gradients
    tau = 0.1
    qd = q.double().to(DEV).requires_grad_(True)
    kd = k.double().to(DEV).requires_grad_(True)
    qn, kn = torch.nn.functional.normalize(qd, dim=-1), torch.nn.functional.normalize(kd, dim=-1)
    logits = qn @ kn.t() / tau
    labels = torch.arange(m, device=DEV)
    ref = torch.nn.functional.cross_entropy(logits, labels)
    if symmetric:
        ref = 0.5 * ref + 0.5 * torch.nn.functional.cross_entropy(logits.t(), labels)
    ref.backward()
    a = q.to(DEV).requires_grad_(True)
    b = k.to(DEV).requires_grad_(True)
    loss = ops.info_nce(a, b, temperature=tau, reduction="mean", symmetric=symmetric)
    loss.backward()
    torch.testing.assert_close(loss.double(), ref.detach(), rtol=1e-4, atol=1e-5)
    for got, want in ((a.grad, qd.grad), (b.grad, kd.grad)):
        assert float((got.double() - want).norm() / want.norm()) < 1e-3


def test_infonce_reductions():
    q = torch.randn(9, 512, device=DEV, requires_grad=True)
    k = torch.randn(9, 512, device=DEV, requires_grad=True)
    none = ops.info_nce(q, k, temperature=0.1, reduction="none", symmetric=True)
    mean = ops.info_nce(q, k, temperature=0.1, reduction="mean", symmetric=True)
    ssum = ops.info_nce(q, k, temperature=0.1, reduction="sum", symmetric=True)
    torch.testing.assert_close(none.mean(), mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(none.sum(), ssum, rtol=1e-5, atol=1e-5)
    w = torch.rand(9, device=DEV)
    (none * w).sum().backward()
    qn = torch.nn.functional.normalize(q.detach(), dim=-1).requires_grad_(True)
    # torch reference
    q2 = q.detach().clone().requires_grad_(True)
    k2 = k.detach().clone().requires_grad_(True)
    L = torch.nn.functional.normalize(q2, dim=-1) @ torch.nn.functional.normalize(k2, dim=-1).t() / 0.1
    lab = torch.arange(9, device=DEV)
    ref = 0.5 * torch.nn.functional.cross_entropy(L, lab, reduction="none") + 0.5 * torch.nn.functional.cross_entropy(L.t(), lab, reduction="none")
    (ref * w).sum().backward()
    torch.testing.assert_close(none, ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(q.grad, q2.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(k.grad, k2.grad, rtol=1e-3, atol=1e-5)


def test_pool_empty_and_single_token_bags():
    """Edge cases of the packed layout: an empty bag pools to zeros, a one-token bag returns that token."""
    H, E = 4, 512
    C = H * E
    lens = [0, 1, 5, 0, 3]
    cu, M = _ragged(lens)
    x = torch.randn(M, C, device=DEV)
    xp = ops.split_planes(x, 2)
    logits = torch.randn(M, H, device=DEV)
    out = torch.full((len(lens), C), 7.0, device=DEV)
    attn = torch.zeros(M, H, device=DEV)
    ops.pool_fwd(xp, 2, logits, cu, None, len(lens), M, H, E, out, attn, 0)
    xr = planes_f32(xp)
    assert float(out[0].abs().max()) == 0.0 and float(out[3].abs().max()) == 0.0
    torch.testing.assert_close(out[1], xr[0], rtol=1e-6, atol=1e-7)
    p = torch.softmax(logits[1:6], dim=0)
    torch.testing.assert_close(out[2], (xr[1:6].view(5, H, E) * p[:, :, None]).sum(0).reshape(C), rtol=1e-5, atol=1e-6)
    dS = torch.randn(len(lens), C, device=DEV)
    dlogit = torch.zeros(M, H, device=DEV)
    call("mdl_pool_bwd_dlogit", xp, M * C, 2, dS, out, attn, cu, None, len(lens), M, H, E, dlogit, 0, logits, 0, 1, _st())
    assert torch.isfinite(dlogit).all()
    assert float(dlogit[0].abs().max()) < 1e-3          # a one-token bag has a constant softmax: zero logit gradient (up to fp32 rounding of two 512-term dots)


def test_infonce_rows_equals_gathered_operands():
    """mdl_infonce_rows_fwd/bwd (operands = indexed rows of one matrix, several terms accumulated into one gradient) against
    the dense entry points on explicitly gathered copies, and against torch autograd through the index ops."""
    from madeleine_b200 import ops
    from madeleine_b200.utils.loss import InfoNCE
    g = torch.Generator().manual_seed(3)
    base = torch.randn(40, 512, generator=g).to(DEV).requires_grad_(True)
    pairs = [(torch.tensor([0, 2, 4, 6, 8, 10]), torch.tensor([1, 3, 5, 7, 9, 11])),
             (torch.tensor([0, 4, 8, 20, 30]), torch.tensor([13, 15, 17, 19, 39]))]        # H&E rows shared between the terms
    for tau, sym in ((0.1, True), (0.001, True), (0.07, False)):
        base.grad = None
        total = ops.info_nce_rows(base, pairs, [tau] * len(pairs), [sym] * len(pairs))
        total.backward()
        got = base.grad.clone()
        base.grad = None
        fn = InfoNCE(temperature=tau)
        ref = sum(fn(base[q.to(DEV)], base[k.to(DEV)], symmetric=sym) for q, k in pairs)
        ref.backward()
        torch.testing.assert_close(total.detach(), ref.detach(), rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(got, base.grad, rtol=1e-5, atol=1e-7)
    # upstream scaling
    base.grad = None
    (ops.info_nce_rows(base, pairs, [0.1, 0.1], [True, True]) * 3.0).backward()
    g3 = base.grad.clone()
    base.grad = None
    ops.info_nce_rows(base, pairs, [0.1, 0.1], [True, True]).backward()
    torch.testing.assert_close(g3, 3.0 * base.grad, rtol=1e-6, atol=1e-8)
