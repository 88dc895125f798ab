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
    v = torch.randn(2, 200, 128, device=DEV)
    with pytest.raises(RuntimeError, match="at most"):
        ops.got_loss(v, v)
