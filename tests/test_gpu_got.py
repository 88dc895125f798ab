"""Graph-OT local loss kernel (forward + hand-written reverse sweep) against fixtures from the real reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from madeleine_b200 import ops  # noqa: E402
from madeleine.utils.loss import GOT  # noqa: E402

DEV = "cuda"


def test_got_golden_loss_and_grads(golden):
    for c in golden("got")["cases"]:
        v = c["v"].to(DEV).requires_grad_(True)
        q = c["q"].to(DEV).requires_grad_(True)
        torch.manual_seed(c["torch_seed"])          # GOT draws randperm(m) from the global CPU generator (quirk Q3)
        loss = GOT(v, q, subsample=256)
        loss.backward()
        torch.testing.assert_close(loss.cpu(), c["loss"], rtol=1e-3, atol=1e-4)
        for got, ref in ((v.grad.cpu(), c["dv"]), (q.grad.cpu(), c["dq"])):
            torch.testing.assert_close(got, ref, rtol=2e-2, atol=2e-3 * float(ref.abs().max()))
        # tokens beyond the first m are never read (quirk Q3) -> exactly zero gradient
        m = v.shape[0]
        assert float(v.grad[:, m:].abs().max() if v.shape[1] > m else 0.0) == 0.0


def test_got_parts_against_oracle(golden):
    """wd / gwd per problem against the CPU oracle on the same (already subsampled) tokens."""
    import oracle
    g = golden("got")["internals"]
    v = g["x"].transpose(1, 2).contiguous()      # [b, n, D]
    q = g["y"].transpose(1, 2).contiguous()
    loss = ops.got_loss(v.to(DEV), q.to(DEV))
    ref = oracle.got(v, q)
    torch.testing.assert_close(loss.cpu(), ref, rtol=1e-3, atol=1e-4)


def test_got_deterministic_and_scaled_backward():
    torch.manual_seed(0)
    v = torch.randn(6, 6, 128, device=DEV, requires_grad=True)
    q = (v.detach() + 0.3 * torch.randn(6, 6, 128, device=DEV)).requires_grad_(True)
    l1 = ops.got_loss(v, q)
    (3.0 * l1).backward()
    g1 = v.grad.clone()
    v.grad = None
    l2 = ops.got_loss(v, q)
    l2.backward()
    assert torch.equal(l1, l2)
    torch.testing.assert_close(g1, 3.0 * v.grad, rtol=1e-6, atol=1e-9)


def test_got_too_many_tokens():
    v = torch.randn(2, 300, 128, device=DEV)          # GOT(..., subsample=256) never produces more than 256 tokens
    with pytest.raises(RuntimeError, match="at most"):
        ops.got_loss(v, v)


def _tokens(m, n, seed):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(m, n, 128, generator=g)
    q = v + 0.5 * torch.randn(m, n, 128, generator=g)          # correlated stains, as real HE/IHC tokens are
    return v, q


@pytest.fixture
def force_big():
    from madeleine_b200._lib import call
    call("mdl_got_force_big", 1)
    yield
    call("mdl_got_force_big", 0)


@pytest.mark.parametrize("m,n", [(3, 6), (4, 53), (2, 96)])
def test_got_big_kernels_match_shared_memory_kernels(m, n, force_big):
    """The global-memory kernels (got_big.cu, used for n > 96) against the shared-memory ones on problems both can take."""
    from madeleine_b200._lib import call
    v, q = _tokens(m, n, 10 + n)
    res = []
    for big in (0, 1):
        call("mdl_got_force_big", big)
        a = v.to(DEV).requires_grad_(True)
        b = q.to(DEV).requires_grad_(True)
        loss = ops.got_loss(a, b)
        loss.backward()
        res.append((loss.detach().cpu(), a.grad.cpu(), b.grad.cpu()))
    (l0, dv0, dq0), (l1, dv1, dq1) = res
    torch.testing.assert_close(l1, l0, rtol=1e-4, atol=1e-5)
    for g1, g0 in ((dv1, dv0), (dq1, dq0)):
        # different summation orders; single entries of tiny problems are ill-conditioned (relu thresholds), so norm-wise
        assert float((g1 - g0).norm() / g0.norm()) < 1e-2


def test_got_big_golden(golden, force_big):
    """The reference's own GOT fixtures through the global-memory kernels."""
    for c in golden("got")["cases"]:
        v = c["v"].to(DEV).requires_grad_(True)
        q = c["q"].to(DEV).requires_grad_(True)
        torch.manual_seed(c["torch_seed"])
        loss = GOT(v, q, subsample=256)
        loss.backward()
        torch.testing.assert_close(loss.cpu(), c["loss"], rtol=1e-3, atol=1e-4)
        for got, ref in ((v.grad.cpu(), c["dv"]), (q.grad.cpu(), c["dq"])):
            torch.testing.assert_close(got, ref, rtol=2e-2, atol=2e-3 * float(ref.abs().max()))


@pytest.mark.parametrize("m,n", [(2, 97), (3, 130), (2, 256)])
def test_got_large_problems_against_oracle(m, n):
    """96 < n <= 256 tokens per problem (batches of more than 96 cases): loss and token gradients against the CPU oracle."""
    import oracle
    v, q = _tokens(m, n, n)
    vo = v.clone().requires_grad_(True)
    qo = q.clone().requires_grad_(True)
    ref = oracle.got(vo, qo)
    ref.backward()
    a = v.to(DEV).requires_grad_(True)
    b = q.to(DEV).requires_grad_(True)
    loss = ops.got_loss(a, b)
    loss.backward()
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=1e-3, atol=1e-4)
    for got, want in ((a.grad.cpu(), vo.grad), (b.grad.cpu(), qo.grad)):
        assert float((got - want).norm() / want.norm()) < 2e-2


def test_got_subsample_256_of_a_large_batch():
    """m = 260 cases: the permutation over the cases keeps 256 token indices (loss.py:281-284) -> n = 256 problems."""
    import oracle
    m, T = 260, 264
    v, q = _tokens(m, T, 5)
    torch.manual_seed(3)
    ref = oracle.got(v[:4].clone(), q[:4].clone(), subsample=256, perm=torch.randperm(m))   # oracle on 4 of the cases
    torch.manual_seed(3)
    perm = torch.randperm(m)[:256]
    loss4 = ops.got_loss(v[:4][:, perm].to(DEV), q[:4][:, perm].to(DEV))
    torch.testing.assert_close(loss4.cpu(), ref, rtol=1e-3, atol=1e-3)
    torch.manual_seed(3)
    full = GOT(v.to(DEV), q.to(DEV), subsample=256)            # all 260 problems of 256 tokens in one call
    assert bool(torch.isfinite(full))


def test_got_at_the_shipped_batch_size_against_the_real_reference(golden):
    """65 cases with the stain (the reference's shipped batch of 65: problems of 65 tokens) and 96 (the largest problem of the
    shared-memory kernel) against the REAL reference (tests/golden/got_shipped_batch.pt): loss, gradient norms, the full gradient of
    two cases and 512 sampled entries."""
    from weights import make_feats
    for c in golden("got_shipped_batch")["cases"]:
        m = c["m"]
        v0 = make_feats(c["seed_v"], m, m + 16, 128)
        q0 = v0 + 0.5 * make_feats(c["seed_q"], m, m + 16, 128)
        v, q = v0.to(DEV).requires_grad_(True), q0.to(DEV).requires_grad_(True)
        torch.manual_seed(c["torch_seed"])
        loss = GOT(v, q, subsample=256)
        loss.backward()
        torch.testing.assert_close(loss.detach().cpu(), c["loss"], rtol=1e-3, atol=1e-4)
        for t, d in ((v.grad.cpu(), c["dv"]), (q.grad.cpu(), c["dq"])):
            assert float(t.double().norm()) == pytest.approx(float(d["norm"]), rel=2e-2)
            torch.testing.assert_close(t[:2], d["first2"], rtol=2e-2, atol=2e-3 * float(d["first2"].abs().max()))
            torch.testing.assert_close(t.flatten()[d["idx"]], d["samples"], rtol=2e-2, atol=2e-3 * float(d["samples"].abs().max()))
            assert float(t[:, m:].abs().max()) == 0.0
