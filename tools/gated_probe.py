"""Where does the gated-attention GEMM (mdl_gemm_gated, abmil.py:49-52 for all heads) spend its time?  Times the launch at
the bench size (64 000 tokens, fp32-grade 3-pass and bf16 1-pass) with dropout off / on and with / without saving the fp16
gates for backward.  One JSON line."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from madeleine_b200 import ops  # noqa: E402
from madeleine_b200._lib import call, stream_ptr  # noqa: E402

dev = torch.device("cuda", 0)
M, H, C = 64000, 4, 2048
res = {}
for nsplit, npl in ((3, 2), (1, 1)):
    h3 = ops.split_planes(torch.randn(M, C, device=dev), npl)
    wab = ops.split_planes(torch.randn(H * 1024, 512, device=dev) * 0.04, npl)
    vec = lambda n: torch.randn(n, device=dev) * 0.04  # noqa: E731
    ba, bb, wc, bc = vec(H * 512), vec(H * 512), vec(H * 512), vec(H)
    logits = torch.empty(M, H, device=dev)
    ga = ops.gate_buffer(M, H * 512, dev)
    gb = ops.gate_buffer(M, H * 512, dev)
    st = stream_ptr(dev)
    for p in (0.0, 0.25):
        for keep in (False, True):
            def run():
                call("mdl_gemm_gated", h3, M, C, C, M * C, wab, wab.shape[1] * wab.shape[2], M, H, nsplit, ba, bb, wc, bc, logits,
                     ga if keep else None, gb if keep else None, p, 1234, st)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            flops = 2.0 * M * 4096 * 512 * nsplit
            res[f"nsplit{nsplit}_p{p}_gates{int(keep)}"] = {"ms": round(ms, 4), "issued_tflops": round(flops / ms / 1e9, 1)}
print(json.dumps(res))
