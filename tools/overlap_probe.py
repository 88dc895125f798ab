"""Can the HBM/issue-bound elementwise kernels of one half-batch hide under the tensor-bound GEMMs of the other?
Runs the bench step as ONE batch of 32 bags and as TWO half-batches of 16 bags on two CUDA streams (same total work) and
prints ms per 32 bags for each."""
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
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
cfg = Namespace(MODALITIES=MODS, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax",
                n_heads=4, b200_precision=prec)
model = MADELEINE(cfg, stain_encoding=False)
model.load_state_dict(make_state_dict(0, n_mod=2), strict=True)
model.to(dev).train()
loss_fn = InfoNCE(temperature=0.001)
largs = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
full = torch.randn(16, 2, 2000, 512, device=dev)
halves = [full[:8].contiguous(), full[8:].contiguous()]
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]


def step(feats):
    embs, toks = model({"feats": feats}, device=dev, n_views=1)
    lab = torch.ones(feats.shape[0], 2)
    loss, _ = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, lab[:, 1:], largs)
    loss.backward()


def one():
    model.zero_grad(set_to_none=True)
    step(full)


def two_streams():
    model.zero_grad(set_to_none=True)
    cur = torch.cuda.current_stream(dev)
    for s, h in zip(streams, halves):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            step(h)
    for s in streams:
        cur.wait_stream(s)


def two_sequential():
    model.zero_grad(set_to_none=True)
    for h in halves:
        step(h)


def timed(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(json.dumps({"precision": prec, "one_batch_of_32_bags_ms": round(timed(one), 3),
                  "two_half_batches_sequential_ms": round(timed(two_sequential), 3),
                  "two_half_batches_on_two_streams_ms": round(timed(two_streams), 3)}))
