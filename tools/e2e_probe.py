"""Where does the end-to-end leg of bench.py lose time against the device-resident one?  Runs the bench step fed from
pinned host memory under a few staging variants and prints one JSON line each (ms per step).

    python tools/e2e_probe.py [--steps 20]
"""
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
from madeleine_b200.utils.prefetch import DevicePrefetcher  # noqa: E402
from weights import make_state_dict  # noqa: E402

B, S, T, D = 16, 2, 2000, 512
MODS = ["HE", "IHC"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = Namespace(MODALITIES=MODS, wsi_encoder="abmil", patch_embedding_dim=D, wsi_encoder_hidden_dim=512, activation="softmax",
                    n_heads=4, b200_precision="fp32")
    model = MADELEINE(cfg, stain_encoding=False)
    model.load_state_dict(make_state_dict(0, n_mod=2), strict=True)
    model.to(dev).train()
    loss_fn = InfoNCE(temperature=0.001)
    largs = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    labels = torch.ones(B, S)
    host = [torch.randn(B, S, T, D).pin_memory() for _ in range(2)]
    feats_dev = torch.randn(B, S, T, D, device=dev)

    def step(feats):
        model.zero_grad(set_to_none=True)
        embs, toks = model({"feats": feats}, device=dev, n_views=1)
        loss, _ = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, labels[:, 1:], largs)
        loss.backward()
        return loss

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(n)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def resident(n):
        for _ in range(n):
            step(feats_dev)

    def prefetch(lag, readback):
        def run(n):
            batches = ({"feats": host[i % 2]} for i in range(n))
            hl = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(lag + 1)]
            ev = [torch.cuda.Event() for _ in range(lag + 1)]
            for i, b in enumerate(DevicePrefetcher(batches, dev)):
                loss = step(b["feats"])
                if readback:
                    hl[i % (lag + 1)].copy_(loss.detach(), non_blocking=True)
                    ev[i % (lag + 1)].record()
                    if i >= lag:
                        ev[(i - lag) % (lag + 1)].synchronize()
                        float(hl[(i - lag) % (lag + 1)])
        return run

    def fixed_buffers(n):
        """H2D into two preallocated device buffers (no allocator traffic), copy stream + events."""
        bufs = [torch.empty(B, S, T, D, device=dev) for _ in range(2)]
        cs = torch.cuda.Stream(dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(cs):
            bufs[0].copy_(host[0], non_blocking=True)
            ready[0].record(cs)
        for i in range(n):
            k = i % 2
            if i + 1 < n:
                with torch.cuda.stream(cs):
                    if i >= 1:
                        cs.wait_event(free[1 - k])
                    bufs[1 - k].copy_(host[(i + 1) % 2], non_blocking=True)
                    ready[1 - k].record(cs)
            cur.wait_event(ready[k])
            step(bufs[k])
            free[k].record(cur)

    def sync_copy(n):
        for i in range(n):
            step(host[i % 2].to(dev, non_blocking=True))

    for _ in range(5):
        step(feats_dev)
    variants = [("device-resident", resident), ("prefetch lag2 readback (bench e2e)", prefetch(2, True)),
                ("prefetch lag4 readback", prefetch(4, True)), ("prefetch no readback", prefetch(2, False)),
                ("fixed double buffer, no readback", fixed_buffers), ("same-stream H2D, no readback", sync_copy)]
    for name, fn in variants:
        fn(3)
        print(json.dumps({"variant": name, "ms_per_step": round(timed(fn, a.steps), 4)}))


if __name__ == "__main__":
    main()
