"""The callers either side of the hot path, as the reference's scripts use them (SURVEY.md §8b / §8f-1,2): train_loop,
run_inference, create_model / load_checkpoint / create_model_from_pretrained."""
import json
import os
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from madeleine.models.Model import MADELEINE, create_model  # noqa: E402
from madeleine.models.factory import create_model_from_pretrained  # noqa: E402
from madeleine.utils.loss import InfoNCE, GOT  # noqa: E402
from madeleine.utils.trainer import train_loop  # noqa: E402
from madeleine.utils.utils import run_inference, load_checkpoint  # noqa: E402
from madeleine_b200.optim import FusedAdamW  # noqa: E402
from weights import make_state_dict, make_feats  # noqa: E402

DEV = torch.device("cuda")
MODS = ["HE", "ER", "PR"]


def _cfg(mods=MODS, precision="float32"):
    return dict(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512,
                activation="softmax", n_heads=4, precision=precision)


def _batches(n, bs=4, T=48, seed=0, he_only_at=None):
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(n):
        labels = (torch.rand(bs, len(MODS), generator=g) < 0.8).float()
        labels[:, 0] = 1
        labels[:2] = 1
        if he_only_at == b:
            labels[:, 1:] = 0
        feats = torch.randn(bs, len(MODS), T, 512, generator=g) * labels[:, :, None, None]
        out.append({"feats": feats, "modality_labels": labels, "slide_ids": [f"b{b}c{i}" for i in range(bs)]})
    return out


@pytest.mark.parametrize("precision,fused", [("float32", True), ("bfloat16", True), ("float32", False)])
def test_train_loop_runs_like_the_reference_script(precision, fused, capsys):
    torch.manual_seed(0)
    model = MADELEINE(Namespace(**_cfg()), stain_encoding=True).to(DEV)
    args = Namespace(precision=precision, STAINS=MODS[1:], warmup_epochs=0, global_loss="info-nce", symmetric_cl=True,
                     local_loss_weight=1.0)
    opt = FusedAdamW(model.parameters(), lr=1e-3) if fused else torch.optim.AdamW(model.parameters(), lr=1e-3)
    warm = torch.optim.lr_scheduler.LinearLR(opt, start_factor=0.1, total_iters=4)
    cos = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=8)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    data = _batches(4, he_only_at=2)
    losses = []
    for epoch in range(6):
        # global + local loss in the first epochs (dropout and GOT's random permutation make that objective noisy), then the
        # global loss alone, whose decrease on the same three usable batches is unambiguous
        local = GOT if epoch < 2 else None
        ep_loss, rank = train_loop(args, InfoNCE(temperature=0.1), local, None, model, epoch, data, opt, warm, cos)
        assert np.isfinite(ep_loss) and 1.0 <= rank <= 512.0
        losses.append(ep_loss)
    assert "Skipping batch with only HE" in capsys.readouterr().out
    assert losses[5] < losses[2]
    changed = [n for n, p in model.named_parameters() if not torch.equal(p.detach(), before[n])]
    assert len(changed) >= 38                                      # every tensor of the checkpoint layout was updated


def test_create_model_checkpoint_roundtrip_and_factory(tmp_path):
    sd = make_state_dict(12, n_mod=1)
    cfg = Namespace(**_cfg(mods=["HE"]))
    ref = MADELEINE(cfg, stain_encoding=False)
    ref.load_state_dict(sd, strict=True)
    ref.to(DEV).eval()
    x = make_feats(4, 2, 100, 512)
    with torch.no_grad():
        want = ref.encode_he(x, DEV)
    # DataParallel-style keys (Model.py:36-38)
    torch.save({("module." + k): v for k, v in sd.items()}, tmp_path / "dp.pt")
    m1 = create_model(cfg, device="cuda", checkpoint_path=str(tmp_path / "dp.pt")).eval()
    with torch.no_grad():
        assert torch.equal(m1.encode_he(x, DEV), want)
    # load_checkpoint (utils.py:92-122): plain keys, then the module.-prefixed retry
    m2 = MADELEINE(cfg, stain_encoding=False).to(DEV)
    torch.save(sd, tmp_path / "model.pt")
    load_checkpoint(Namespace(RESULS_SAVE_PATH=str(tmp_path)), m2)
    m3 = MADELEINE(cfg, stain_encoding=False).to(DEV)
    load_checkpoint(None, m3, path_to_checkpoint=str(tmp_path / "dp.pt"))
    with torch.no_grad():
        assert torch.equal(m2.eval().encode_he(x, DEV), want) and torch.equal(m3.eval().encode_he(x, DEV), want)
    # HF bundle layout: model.pt + model_config.json already on disk -> no download (factory.py:16-39)
    json.dump(_cfg(mods=["HE"], precision="bfloat16"), open(tmp_path / "model_config.json", "w"))
    m4, precision = create_model_from_pretrained(str(tmp_path))
    assert precision is torch.bfloat16
    with torch.no_grad():
        assert torch.equal(m4.eval().encode_he(x, DEV), want)
    # strict: a missing key must fail like the reference
    bad = dict(sd)
    bad.pop("projector.bias")
    torch.save(bad, tmp_path / "bad.pt")
    with pytest.raises(RuntimeError):
        create_model(cfg, device="cuda", checkpoint_path=str(tmp_path / "bad.pt"))


def test_run_inference_output_format():
    """utils.py:27-66: {"embeds": [n, 512] fp32 numpy, "slide_ids": [...]} + the rank metric, bs = 1 batches of any length."""
    sd = make_state_dict(12, n_mod=1)
    model = MADELEINE(Namespace(**_cfg(mods=["HE"])), stain_encoding=False)
    model.load_state_dict(sd, strict=True)
    model.to(DEV)
    lens = [50, 333, 7, 128, 64, 90]
    loader = [(make_feats(i, 1, n, 512), [f"slide{i}"]) for i, n in enumerate(lens)]
    res, rank = run_inference(model, loader, config=Namespace(precision="float32"))
    assert res["embeds"].shape == (len(lens), 512) and res["embeds"].dtype == np.float32
    assert res["slide_ids"] == [f"slide{i}" for i in range(len(lens))]
    assert 1.0 <= rank <= len(lens) + 1e-3
    with torch.no_grad():
        want = model.encode_he(loader[1][0], DEV).cpu().numpy()
    np.testing.assert_allclose(res["embeds"][1], want[0], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("activation", ["softmax", "sigmoid"])
def test_standalone_batched_abmil_head(activation):
    """BatchedABMIL used on its own (abmil.py:41-68): activated attention and raw logits against plain torch math."""
    from madeleine.models.abmil import BatchedABMIL
    torch.manual_seed(3)
    head = BatchedABMIL(input_dim=512, hidden_dim=512, dropout=True, n_classes=1, n_heads=1, activation=activation).to(DEV).eval()
    x = torch.randn(2, 77, 512, device=DEV)
    with torch.no_grad():
        act, raw = head(x, return_raw_attention=True)
        a = torch.tanh(torch.nn.functional.linear(x.double(), head.attention_a[0].weight.double(), head.attention_a[0].bias.double()))
        b = torch.sigmoid(torch.nn.functional.linear(x.double(), head.attention_b[0].weight.double(), head.attention_b[0].bias.double()))
        ref = torch.nn.functional.linear(a * b, head.attention_c.weight.double(), head.attention_c.bias.double())
    assert raw.shape == (2, 77, 1) and act.shape == (2, 77, 1)
    torch.testing.assert_close(raw.double(), ref, rtol=1e-3, atol=1e-4)
    want = torch.softmax(ref, dim=1) if activation == "softmax" else torch.sigmoid(ref)
    torch.testing.assert_close(act.double(), want, rtol=1e-3, atol=1e-5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        head(x.cpu())
