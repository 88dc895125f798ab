"""CPU tests of host-side logic that needs no GPU."""
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
