"""Which resource bounds the K-major fused-pass GEMM?  Times representative shapes with operand loads and/or epilogue
writes disabled (results are wrong in those modes; this is a measurement tool only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from madeleine_b200 import ops
from madeleine_b200._lib import call

DEV = "cuda"
SHAPES = [("L1/L2 fwd", 64000, 512, 512), ("L3 fwd", 64000, 2048, 512), ("attn dgrad", 64000, 2048, 1024), ("L3 dgrad", 64000, 512, 2048)]


def timeit(fn, n=10):
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


for name, M, N, K in SHAPES:
    A = ops.split_planes(torch.randn(M, K, device=DEV), 2)
    B = ops.split_planes(torch.randn(N, K, device=DEV), 2)
    out = torch.empty(M, N, device=DEV)
    res = {}
    for flags, tag in ((0, "normal"), (1, "no_tma"), (2, "no_store"), (3, "mma_only")):
        call("mdl_gemm_debug_flags", flags)
        ms = timeit(lambda: ops.gemm_nt(A, K, (B, N, K, N * K), N, 3, out=out))
        res[tag] = ms
    call("mdl_gemm_debug_flags", 0)
    gf = 2.0 * M * N * K * 3 / 1e9
    print(f"{name:11s} M={M} N={N} K={K}: " + "  ".join(f"{t}={v*1e3:7.1f}us ({gf/v/1e3:6.0f} TF)" for t, v in res.items()))
