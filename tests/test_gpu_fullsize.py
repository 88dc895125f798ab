"""BASELINE-size checks through size-independent properties (the oracle only finishes small cases in seconds).

configs[1]: 16 cases x 2 stains x 2000 tokens.  Properties: packing independence (a batch equals its bags encoded one
by one), token-permutation invariance of a bag, attention weights sum to one, pooling linearity, loss symmetry,
zero gradient for parameters that cannot influence the loss, finite gradients everywhere."""
from argparse import Namespace

import pytest
import torch

pytestmark = pytest.mark.gpu

from madeleine.models.Model import MADELEINE  # noqa: E402
from madeleine.utils.loss import InfoNCE  # noqa: E402
from madeleine.utils.trainer import calculate_losses  # noqa: E402
from madeleine_b200 import ops  # noqa: E402
from weights import make_state_dict  # noqa: E402

DEV = torch.device("cuda")
B, S, T, D = 16, 2, 2000, 512


def _model(mods=("HE", "IHC")):
    cfg = Namespace(MODALITIES=list(mods), wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512,
                    activation="softmax", n_heads=4, b200_precision="fp32")
    m = MADELEINE(cfg, stain_encoding=False)
    m.load_state_dict(make_state_dict(0, n_mod=len(mods)))
    return m.to(DEV).eval()


@pytest.fixture(scope="module")
def feats():
    g = torch.Generator(device=DEV).manual_seed(11)
    return torch.randn(B, S, T, D, generator=g, device=DEV)


def test_packing_independence_and_ragged(feats):
    """A packed batch of 32 bags equals the same bags encoded separately; ragged lengths N in [200, 4000] too."""
    model = _model()
    with torch.no_grad():
        batch = model.encode_he(feats.view(B * S, T, D), DEV)
        for r in (0, 7, 31):
            single = model.encode_he(feats.view(B * S, T, D)[r:r + 1], DEV)
            torch.testing.assert_close(batch[r:r + 1], single, rtol=1e-5, atol=1e-6)
        g = torch.Generator().manual_seed(1234)
        lens = torch.randint(200, 4001, (32,), generator=g).tolist()
        cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32)
        x = torch.randn(sum(lens), D, device=DEV)
        packed = model.encode_packed(x, cu)
        for r in (0, 13, 31):
            one = model.encode_packed(x[cu[r]:cu[r + 1]], torch.tensor([0, lens[r]], dtype=torch.int32))
            torch.testing.assert_close(packed[r:r + 1], one, rtol=1e-5, atol=1e-6)


def test_token_permutation_invariance(feats):
    model = _model()
    x = feats[0, 0]
    perm = torch.randperm(T, device=DEV)
    with torch.no_grad():
        a = model.encode_he(x[None], DEV)
        b = model.encode_he(x[perm][None], DEV)
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)


def test_attention_weights_and_pooling_linearity():
    H, E, R = 4, 512, 32
    C = H * E
    M = R * T
    cu = torch.arange(0, (R + 1) * T, T, dtype=torch.int32, device=DEV)
    x = torch.randn(M, C, device=DEV)
    logits = torch.randn(M, H, device=DEV) * 4
    xp = ops.split_planes(x, 2)
    out1 = torch.empty(R, C, device=DEV)
    attn = torch.empty(M, H, device=DEV)
    ops.pool_fwd(xp, 2, logits, cu, None, R, M, H, E, out1, attn, 0)
    sums = attn.view(R, T, H).sum(1)
    torch.testing.assert_close(sums, torch.ones_like(sums), rtol=1e-5, atol=1e-5)       # softmax normalisation per (bag, head)
    # shift invariance of softmax and linearity in X:  pool(2X, logits + c) = 2 pool(X, logits)
    out2 = torch.empty(R, C, device=DEV)
    ops.pool_fwd(ops.split_planes(2 * x, 2), 2, logits + 3.0, cu, None, R, M, H, E, out2, None, 0)
    torch.testing.assert_close(out2, 2 * out1, rtol=1e-4, atol=1e-5)
    # a one-hot attention (huge logit on one token) returns that token's features
    logits2 = torch.full((M, H), -50.0, device=DEV)
    pick = torch.randint(0, T, (R,), device=DEV) + torch.arange(R, device=DEV) * T
    logits2[pick] = 50.0
    out3 = torch.empty(R, C, device=DEV)
    ops.pool_fwd(xp, 2, logits2, cu, None, R, M, H, E, out3, None, 0)
    torch.testing.assert_close(out3, xp.float().sum(0)[pick], rtol=1e-5, atol=1e-6)


def test_training_step_properties(feats):
    """Full configs[1] step: loss symmetric under swapping the two stains' roles, finite gradients, token_projector gets
    no gradient when only the global loss is used, and two runs give the same loss (forward is deterministic)."""
    model = _model()
    args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    labels = torch.ones(B, 1)
    loss_fn = InfoNCE(temperature=0.001)

    def run(x):
        model.zero_grad(set_to_none=True)
        embs, toks = model({"feats": x}, device=DEV, n_views=1)
        loss, flag = calculate_losses(["IHC"], loss_fn, None, None, embs, toks, labels, args)
        assert flag
        return loss, embs

    loss, embs = run(feats)
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters()}
    assert grads["token_projector.weight"] is None or float(grads["token_projector.weight"].abs().max()) == 0.0
    for n, gr in grads.items():
        if gr is not None:
            assert torch.isfinite(gr).all(), n
    assert float(grads["projector.weight"].abs().max()) > 0
    loss2, _ = run(feats)
    assert torch.equal(loss.detach(), loss2.detach())
    # symmetric InfoNCE: swapping query and key roles (HE <-> IHC bags) leaves the loss unchanged
    loss_sw, _ = run(feats.flip(1))
    torch.testing.assert_close(loss_sw.detach(), loss.detach(), rtol=1e-5, atol=1e-5)
    he = embs["HE"][:, 0, :, 0]
    assert he.shape == (B, 512) and embs["IHC"].shape == (B, 1, 512)


def test_n_views3_backward_runs():
    """Intra-modality path (n_views=3): half-view pooling participates in backward and gradients are finite."""
    model = _model()
    import numpy as np
    np.random.seed(0)
    x = torch.randn(4, 2, 256, 512, device=DEV)
    embs, toks = model({"feats": x}, device=DEV, n_views=3)
    assert embs["HE"].shape == (4, 3, 512, 1) and embs["IHC"].shape == (4, 3, 512)
    args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    fn = InfoNCE(temperature=0.1)
    loss, _ = calculate_losses(["IHC"], fn, None, fn, embs, toks, torch.ones(4, 1), args)
    loss.backward()
    for n, p in model.named_parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), n
    assert float(model.wsi_embedders.pre_attn[0].weight.grad.abs().max()) > 0
