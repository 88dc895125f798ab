"""Host-side enqueue time of one bench step (Python + ctypes + torch bookkeeping) next to its device time: if the two are
close, the step is launch-bound and kernel speed-ups stop showing."""
import json
import os
import sys
import time
from argparse import Namespace

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
from madeleine.models.Model import MADELEINE  # noqa: E402
from madeleine.utils.loss import InfoNCE  # noqa: E402
from madeleine.utils.trainer import calculate_losses  # noqa: E402
from weights import make_state_dict  # noqa: E402

dev = torch.device("cuda", 0)
MODS = ["HE", "IHC"]
cfg = Namespace(MODALITIES=MODS, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax",
                n_heads=4, b200_precision="fp32")
model = MADELEINE(cfg, stain_encoding=False)
model.load_state_dict(make_state_dict(0, n_mod=2), strict=True)
model.to(dev).train()
loss_fn = InfoNCE(temperature=0.001)
largs = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
labels = torch.ones(16, 2)


def step(feats):
    model.zero_grad(set_to_none=True)
    embs, toks = model({"feats": feats}, device=dev, n_views=1)
    loss, _ = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, labels[:, 1:], largs)
    loss.backward()
    return loss


for T in (2000, 250):
    feats = torch.randn(16, 2, T, 512, device=dev)
    for _ in range(5):
        step(feats)
    torch.cuda.synchronize()
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        step(feats)
    t_host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / n
    print(json.dumps({"tokens_per_bag": T, "host_enqueue_ms_per_step": round(t_host * 1e3, 3), "wall_ms_per_step": round(t_all * 1e3, 3)}))
