"""In-process A/B of a library switch on the bench step (boxes of the pool differ by ~2 %, so two bench runs do not compare):
alternating blocks of steps with the switch off / on, no profiler events between the kernels.
  --switch pdl                      programmatic dependent launch (mdl_set_pdl)
  --switch env:NAME                 an environment variable the library reads per call (off = "0", on = "1")"""
import argparse
import json
import os
import statistics
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
from madeleine_b200.optim import FusedAdamW  # noqa: E402
from weights import make_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--tokens", type=int, default=2000)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--rounds", type=int, default=4)
ap.add_argument("--switch", default="pdl")
ap.add_argument("--power", action="store_true", help="sample board power and SM clock (NVML, 5 ms) during the timed blocks")
a = ap.parse_args()
dev = torch.device("cuda", 0)
MODS = ["HE", "IHC"]
cfg = Namespace(MODALITIES=MODS, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax",
                n_heads=4, b200_precision=a.precision)
model = MADELEINE(cfg, stain_encoding=False)
model.load_state_dict(make_state_dict(0, n_mod=2), strict=True)
model.to(dev).train()
opt = FusedAdamW(model.parameters(), lr=1e-4)
loss_fn = InfoNCE(temperature=0.001)
largs = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
labels = torch.ones(16, 2)
feats = torch.randn(16, 2, a.tokens, 512, device=dev)


def step():
    model.zero_grad(set_to_none=True)
    embs, toks = model({"feats": feats}, device=dev, n_views=1)
    loss, _ = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, labels[:, 1:], largs)
    loss.backward()
    opt.step()


def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def set_switch(on):
    if a.switch == "pdl":
        _lib.call("mdl_set_pdl", on)
    else:
        os.environ[a.switch.split(":", 1)[1]] = str(on)


class Power:
    def __init__(self):
        import threading
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        self.w, self.mhz, self.stop = [], [], threading.Event()
        self.t = threading.Thread(target=self.poll, daemon=True)
        self.t.start()

    def poll(self):
        import time
        while not self.stop.is_set():
            self.w.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
            self.mhz.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            time.sleep(0.005)

    def done(self):
        self.stop.set()
        self.t.join()
        q = lambda v, f: sorted(v)[int(f * (len(v) - 1))]
        return {"samples": len(self.w), "power_w_p10_p50_p90": [q(self.w, .1), q(self.w, .5), q(self.w, .9)],
                "sm_mhz_p10_p50_p90": [q(self.mhz, .1), q(self.mhz, .5), q(self.mhz, .9)],
                "power_limit_w": self.nv.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1e3}


for _ in range(5):
    step()
res = {0: [], 1: []}
pw = Power() if a.power else None
for r in range(a.rounds):
    for on in (0, 1) if r % 2 == 0 else (1, 0):
        set_switch(on)
        timed(5)
        res[on].append(timed(a.steps))
out = {"switch": a.switch, "precision": a.precision, "tokens_per_bag": a.tokens, "steps": a.steps,
       "ms_per_step_off": [round(x, 4) for x in res[0]], "ms_per_step_on": [round(x, 4) for x in res[1]],
       "median_off": round(statistics.median(res[0]), 4), "median_on": round(statistics.median(res[1]), 4)}
if pw is not None:
    out["nvml"] = pw.done()
print(json.dumps(out))
