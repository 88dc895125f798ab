"""Two plain bench steps (train mode, fp32-grade, no optimiser) for ncu captures of individual kernels:
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c <n> -o gpurun_out/<name> python tools/ncu_targets.py
"""
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
from weights import make_state_dict  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
MODS = ["HE", "IHC"]
cfg = Namespace(MODALITIES=MODS, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax",
                n_heads=4, b200_precision=precision)
model = MADELEINE(cfg, stain_encoding=False)
model.load_state_dict(make_state_dict(0, n_mod=2), strict=True)
model.to(dev).train()
loss_fn = InfoNCE(temperature=0.001)
largs = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
labels = torch.ones(16, 2)
feats = torch.randn(16, 2, 2000, 512, device=dev)
for _ in range(steps):
    model.zero_grad(set_to_none=True)
    embs, toks = model({"feats": feats}, device=dev, n_views=1)
    loss, _ = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, labels[:, 1:], largs)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()))
