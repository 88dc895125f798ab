"""How much of a bench step is launch gaps / host time?  Captures forward + calculate_losses + backward of the bench step
(eval mode, so no per-step dropout seed is baked in) into a CUDA graph and compares replay time with the eager launch
sequence, fp32-grade and bf16.  One JSON line per precision."""
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
from weights import make_state_dict  # noqa: E402

dev = torch.device("cuda", 0)
MODS = ["HE", "IHC"]
labels = torch.ones(16, 2)
largs = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
loss_fn = InfoNCE(temperature=0.001)


def timed(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for precision in ("fp32", "bf16"):
    for T in (2000, 500):
        cfg = Namespace(MODALITIES=MODS, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax",
                        n_heads=4, b200_precision=precision)
        model = MADELEINE(cfg, stain_encoding=False)
        model.load_state_dict(make_state_dict(0, n_mod=2), strict=True)
        model.to(dev).eval()
        feats = torch.randn(16, 2, T, 512, device=dev)

        def step():
            model.zero_grad(set_to_none=True)
            embs, toks = model({"feats": feats}, device=dev, n_views=1)
            loss, _ = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, labels[:, 1:], largs)
            loss.backward()
            return loss

        eager = timed(step)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            static_loss = step()
        replay = timed(g.replay)
        eager_loss = float(step().detach())
        g.replay()
        print(json.dumps({"precision": precision, "tokens_per_bag": T, "eager_ms": round(eager, 4), "graph_replay_ms": round(replay, 4),
                          "gain_ms": round(eager - replay, 4), "loss_eager": eager_loss, "loss_graph": float(static_loss)}))
        del g, model
        torch.cuda.empty_cache()
