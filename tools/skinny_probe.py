"""Isolated timing of the projector ("skinny") kernels, L2 flushed between launches (what ncu's per-launch times show)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madeleine_b200._lib import call, stream_ptr  # noqa: E402

dev = torch.device("cuda", 0)
C, O = 2048, 512
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = stream_ptr(dev)


def timed(fn, n=20):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 1)


for R in (32, 96, 325):
    X = torch.randn(R, C, device=dev); W = torch.randn(O, C, device=dev) / 45; b = torch.randn(O, device=dev)
    Y = torch.empty(R, O, device=dev); dY = torch.randn(R, O, device=dev)
    dX = torch.empty(R, C, device=dev); dW = torch.zeros(O, C, device=dev); db = torch.zeros(O, device=dev)
    null = None
    out = {"R": R,
           "fwd_us": timed(lambda: call("mdl_skinny_linear_fwd", X, W, b, R, C, O, Y, st)),
           "dgrad_us": timed(lambda: call("mdl_skinny_linear_bwd", dY, X, W, R, C, O, dX, null, null, st)),
           "wgrad_us": timed(lambda: call("mdl_skinny_linear_bwd", dY, X, W, R, C, O, null, dW, db, st))}
    print(json.dumps(out))
