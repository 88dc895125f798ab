"""CPU tests of host-side logic that needs no GPU."""
import pytest
import torch

from madeleine_b200.utils.inference import plan_batches
from madeleine_b200.utils.utils import set_model_precision, smooth_rank_measure
from madeleine_b200 import ops


def test_plan_batches():
    assert plan_batches([10, 20, 30, 5], 35) == [[0, 1], [2, 3]]
    assert plan_batches([100], 10) == [[0]]                       # a slide larger than the budget still gets a batch
    assert plan_batches([], 10) == []
    lens = [7] * 10
    b = plan_batches(lens, 21)
    assert [len(x) for x in b] == [3, 3, 3, 1] and sum(b, []) == list(range(10))


def test_precision_resolution(monkeypatch):
    assert ops.resolve_precision("fp32") == "fp32" and ops.resolve_precision("bfloat16") == "bf16"
    monkeypatch.delenv("MADELEINE_B200_PRECISION", raising=False)
    assert ops.resolve_precision(None) == "fp32"                  # no autocast → fp32-grade, like the reference's default
    monkeypatch.setenv("MADELEINE_B200_PRECISION", "bf16")
    assert ops.resolve_precision(None) == "bf16"
    assert set_model_precision("bfloat16") is torch.bfloat16


# values printed by the reference's own smooth_rank_measure (madeleine/utils/utils.py:180-201, executed from /root/reference
# in the build container) on make_feats(1, n, d) * linspace(0.1, 2.0, d)
SMOOTH_RANK_REFERENCE = {(300, 512): 255.13, (40, 512): 39.35, (700, 64): 54.15}


def _rank_input(n, d):
    from weights import make_feats
    return make_feats(1, n, d) * torch.linspace(0.1, 2.0, d)


def test_smooth_rank_measure_matches_reference_values():
    for (n, d), want in SMOOTH_RANK_REFERENCE.items():
        assert smooth_rank_measure(_rank_input(n, d)) == pytest.approx(want, abs=0.011)


@pytest.mark.gpu
def test_smooth_rank_measure_on_device_matches_reference_values():
    """CUDA input: Gram matrix in fp64 on the device (mdl_gram_f64) + eigenvalues instead of a CPU SVD; same numbers."""
    for (n, d), want in SMOOTH_RANK_REFERENCE.items():
        x = _rank_input(n, d).cuda()
        assert smooth_rank_measure(x) == pytest.approx(want, abs=0.011)
        assert smooth_rank_measure(x) == pytest.approx(smooth_rank_measure(x.cpu()), abs=0.011)


def test_smooth_rank_measure():
    x = torch.eye(8)
    assert abs(smooth_rank_measure(x) - 8.0) < 1e-3                # 8 equal singular values → rank 8
    assert smooth_rank_measure(torch.ones(8, 8)) < 1.1            # rank-1 matrix


def test_module_construction_and_state_dict_keys():
    """Checkpoint layout without touching a GPU."""
    from argparse import Namespace
    from madeleine.models.Model import MADELEINE
    from weights import make_state_dict
    cfg = Namespace(MODALITIES=["HE", "ER", "PGR"], wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512,
                    activation="softmax", n_heads=4)
    for se in (False, True):
        m = MADELEINE(cfg, stain_encoding=se)
        sd = make_state_dict(0, n_mod=3, stain_encoding=se)
        assert list(m.state_dict().keys()) == list(sd.keys())
        assert all(m.state_dict()[k].shape == v.shape for k, v in sd.items())
        m.load_state_dict({"module." + k: v for k, v in sd.items()} if False else sd, strict=True)
        assert sum(p.numel() for p in m.parameters()) == (5013284 if se else 4996740) + (0 if not se else (3 - 5) * 32)
    import pytest
    with pytest.raises(ValueError):
        MADELEINE(Namespace(**{**vars(cfg), "wsi_encoder": "transformer"}))


def test_error_behaviour_matches_the_reference():
    """ValueError for an unsupported encoder and for InfoNCE shape mismatches (Model.py:66-67, loss.py:66-89),
    NotImplementedError for unknown aggregation / activation (Model.py:372,443; abmil.py:63) — all raised before any kernel."""
    from argparse import Namespace
    import pytest
    from madeleine.models.Model import MADELEINE, ABMILEmbedder
    from madeleine.models.abmil import BatchedABMIL
    from madeleine.utils.loss import InfoNCE
    cfg = dict(MODALITIES=["HE"], patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax", n_heads=4)
    with pytest.raises(ValueError, match="abmil"):
        MADELEINE(Namespace(wsi_encoder="transformer", **cfg))
    loss = InfoNCE(temperature=0.1)
    q = torch.zeros(4, 8)
    with pytest.raises(ValueError, match="2 dimensions"):
        loss(torch.zeros(4, 2, 8), q)
    with pytest.raises(ValueError, match="same number of samples"):
        loss(q, torch.zeros(5, 8))
    with pytest.raises(ValueError, match="same number of components"):
        loss(q, torch.zeros(4, 9))
    with pytest.raises(ValueError, match="negative_keys"):
        loss(q, q, negative_keys=torch.zeros(3, 2, 8))
    assert loss(q, q, negative_keys=torch.zeros(3, 8)) is None          # the reference's explicit-negatives branch returns None
    emb = ABMILEmbedder({"input_dim": 512, "hidden_dim": 512},
                        {"model": "ABMIL", "params": {"input_dim": 512, "hidden_dim": 512, "dropout": True, "activation": "softmax",
                                                      "n_heads": 4, "n_classes": 1}}, aggregation="max")
    with pytest.raises(NotImplementedError, match="Agg type"):
        emb(torch.zeros(1, 4, 512))
    with pytest.raises(NotImplementedError, match="Attention model"):
        ABMILEmbedder({"input_dim": 512, "hidden_dim": 512}, {"model": "TransMIL", "params": {"n_heads": 4}})
    with pytest.raises(NotImplementedError, match="Activation"):
        BatchedABMIL(512, 512, activation="gelu")(torch.zeros(1, 4, 512))
    model = MADELEINE(Namespace(wsi_encoder="abmil", **cfg))
    keys = set(model.state_dict())
    assert len(keys) == 40 and "wsi_embedders.attn.3.attention_c.bias" in keys and "token_projector.weight" in keys


def test_state_dict_layout_with_stain_encodings():
    from argparse import Namespace
    from madeleine.models.Model import MADELEINE
    cfg = Namespace(MODALITIES=["HE", "A", "B"], wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512,
                    activation="softmax", n_heads=4)
    sd = MADELEINE(cfg, stain_encoding=True).state_dict()
    assert len(sd) == 41 and tuple(sd["embedding.weight"].shape) == (3, 32)      # 40 tensors + the stain embedding table
    assert tuple(sd["wsi_embedders.pre_attn.0.weight"].shape) == (512, 544)
    want = (128 * 2048 + 128) + (512 * 544 + 512) + (512 * 512 + 512) + (2048 * 512 + 2048) + 2 * (512 + 512 + 2048) \
        + 4 * (2 * (512 * 512 + 512) + 512 + 1) + (512 * 2048 + 512) + 3 * 32
    assert sum(v.numel() for v in sd.values()) == want == 5013284 - 2 * 32        # SURVEY §8a1 quotes the 5-modality count


def test_token_window_plan_host_logic():
    """MADELEINE._window_plan: which packed rows token_projector sees, the row -> compact-row map used by the backward
    LayerNorm kernel, and the gather that rebuilds the dense [R, W] token grid (missing bags repeat their single row)."""
    from madeleine.models.Model import MADELEINE
    T, W = 10, 4
    lens = torch.tensor([T, 1, T, 1, 1])                     # bags 1, 3, 4 are missing stains encoded from one token
    cu = torch.zeros(6, dtype=torch.int64)
    cu[1:] = lens.cumsum(0)
    rows, sel_of_row, dense = MADELEINE._window_plan(lens, cu, W, "cpu")
    assert rows.tolist() == [0, 1, 2, 3, 10, 11, 12, 13, 14, 21, 22]
    assert sel_of_row.numel() == int(cu[-1])
    assert [int(sel_of_row[r]) for r in rows.tolist()] == list(range(11))
    assert int((sel_of_row >= 0).sum()) == 11 and int(sel_of_row[4]) == -1 and int(sel_of_row[20]) == -1
    assert dense.view(5, W).tolist() == [[0, 1, 2, 3], [4, 4, 4, 4], [5, 6, 7, 8], [9, 9, 9, 9], [10, 10, 10, 10]]
    # no missing bag: the compact order already is the dense grid
    lens = torch.full((3,), T)
    cu = torch.zeros(4, dtype=torch.int64)
    cu[1:] = lens.cumsum(0)
    rows, sel_of_row, dense = MADELEINE._window_plan(lens, cu, W, "cpu")
    assert dense is None and rows.tolist() == [0, 1, 2, 3, 10, 11, 12, 13, 20, 21, 22, 23]


def test_token_window_resolution(monkeypatch):
    from argparse import Namespace
    from madeleine.models.Model import MADELEINE
    cfg = Namespace(MODALITIES=["HE", "A"], wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512,
                    activation="softmax", n_heads=4)
    monkeypatch.delenv("MADELEINE_B200_TOKEN_WINDOW", raising=False)
    m = MADELEINE(cfg)
    assert m._token_window(8, 2048, {}) == 0                                  # default: the reference's [bs, T, 128] tokens
    monkeypatch.setenv("MADELEINE_B200_TOKEN_WINDOW", "batch")
    assert m._token_window(8, 2048, {}) == 8 and m._token_window(8, 5, {}) == 5   # never more than the bag length
    assert m._token_window(8, 2048, {"b200_token_window": 64}) == 64           # per-batch override (global batch under sharding)
    cfg.b200_token_window = "off"
    assert MADELEINE(cfg)._token_window(8, 2048, {}) == 0                      # config beats the environment
    cfg.b200_token_window = -3
    with __import__("pytest").raises(ValueError):
        MADELEINE(cfg)._token_window(8, 2048, {})
