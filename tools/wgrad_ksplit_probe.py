"""Split-K factor of the wgrad GEMM (mdl_gemm_tn_accum): time per launch for the shapes of one bench step."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madeleine_b200 import ops  # noqa: E402
from madeleine_b200._lib import call, stream_ptr  # noqa: E402

dev = torch.device("cuda")
T = 64000
shapes = {"L1/L2 (512 x 512)": (512, 512, 0, 0), "L3 (2048 x 512)": (2048, 512, 0, 0), "attention (4 x 1024 x 512)": (4096, 512, 1024, 512),
          "L3^T-like (512 x 2048)": (512, 2048, 0, 0)}
for name, (Mo, No, grp, coff) in shapes.items():
    a = torch.randn(2, T, Mo, device=dev).bfloat16()
    b = torch.randn(2, T, 2048 if coff else No, device=dev).bfloat16()
    out = torch.zeros(Mo, No, device=dev)
    res = {}
    for ks in (0, 7, 9, 10, 13, 18, 19, 23, 28, 37, 55, 74, 148):
        def run():
            call("mdl_gemm_tn_accum", a, Mo, Mo, T * Mo, b, b.shape[2], b.shape[2], T * b.shape[2], T, out, No, Mo, No, 3, grp, coff, ks,
                 stream_ptr(dev))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        res[ks] = round(e0.elapsed_time(e1) / 10 * 1e3, 1)
    print(json.dumps({"wgrad": name, "us_per_launch_by_ksplit (0 = auto)": res}))
