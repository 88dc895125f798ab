"""Per-entry-point device time of one bench step (events around every C-ABI call), for any precision mode."""
import argparse
import json
import os
import sys
from argparse import Namespace

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
from madeleine.models.Model import MADELEINE  # noqa: E402
from madeleine.utils.loss import InfoNCE  # noqa: E402
from madeleine.utils.trainer import calculate_losses  # noqa: E402
from madeleine_b200 import _lib  # noqa: E402
from weights import make_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda", 0)
MODS = ["HE", "IHC"]
cfg = Namespace(MODALITIES=MODS, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax",
                n_heads=4, b200_precision=a.precision)
model = MADELEINE(cfg, stain_encoding=False)
model.load_state_dict(make_state_dict(0, n_mod=2), strict=True)
model.to(dev).train()
loss_fn = InfoNCE(temperature=0.001)
largs = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
labels = torch.ones(16, 2)
feats = torch.randn(16, 2, 2000, 512, device=dev)


def step():
    model.zero_grad(set_to_none=True)
    embs, toks = model({"feats": feats}, device=dev, n_views=1)
    loss, _ = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, labels[:, 1:], largs)
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms_plain = e0.elapsed_time(e1) / a.steps
_lib.start_timing("all")
for _ in range(a.steps):
    step()
kt = {k: round(sum(v) / a.steps, 4) for k, v in _lib.stop_timing().items()}
print(json.dumps({"precision": a.precision, "ms_per_step": round(ms_plain, 3), "sum_of_entry_points_ms": round(sum(kt.values()), 3),
                  "entry_point_ms_per_step": dict(sorted(kt.items(), key=lambda kv: -kv[1]))}))
