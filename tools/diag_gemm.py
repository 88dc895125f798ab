"""GPU diagnostic for the tcgen05 GEMM descriptors: structured inputs that localise K-advance / swizzle / major errors.
Prints compact PASS/FAIL lines; never raises (so one call surveys everything)."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from madeleine_b200 import ops
from madeleine_b200._lib import call, stream_ptr

DEV = "cuda"


def report(tag, out, ref):
    err = (out.double() - ref.double()).abs()
    scale = float(ref.abs().max()) + 1e-30
    ok = float(err.max()) <= 1e-4 * scale
    print(f"[{'PASS' if ok else 'FAIL'}] {tag}: max_err={float(err.max()):.3e} scale={scale:.3e} "
          f"nz_out={int((out != 0).sum())}/{out.numel()} nan={int(torch.isnan(out).sum())}")
    if not ok:
        bad = (err > 1e-4 * scale).nonzero()
        rows = sorted(set(bad[:, 0].tolist()))[:12]
        cols = sorted(set(bad[:, 1].tolist()))[:12]
        print(f"       first bad rows {rows} cols {cols}")
        print("       out[0,:8]", [round(float(v), 3) for v in out[0, :8]], " ref[0,:8]", [round(float(v), 3) for v in ref[0, :8]])
    return ok


def nt_case(M, N, K, nsplit, fill):
    npl = 2 if nsplit == 3 else 1
    A, B = fill(M, N, K)
    ap, bp = ops.split_planes(A, npl), ops.split_planes(B, npl)
    out = ops.gemm_nt(ap, K, (bp, N, K, N * K), N, nsplit)
    torch.cuda.synchronize()
    ref = ap.double().sum(0) @ bp.double().sum(0).t() if nsplit == 1 else (
        ap[0].double() @ bp[0].double().t() + ap[0].double() @ bp[1].double().t() + ap[1].double() @ bp[0].double().t())
    return out, ref


def tn_case(T, Mo, No, nsplit, fill):
    npl = 2 if nsplit == 3 else 1
    A, B = fill(T, Mo, No)
    ap, bp = ops.split_planes(A, npl), ops.split_planes(B, npl)
    out = torch.zeros(Mo, No, device=DEV)
    ops.gemm_tn_accum(ap, bp, out, nsplit)
    torch.cuda.synchronize()
    ref = ap[0].double().t() @ bp[0].double()
    if nsplit == 3:
        ref = ref + ap[0].double().t() @ bp[1].double() + ap[1].double().t() @ bp[0].double()
    return out, ref


def main():
    print("device", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    try:
        # 1. one-hot K positions: which k slices are picked up
        for k0 in (0, 8, 15, 16, 31, 32, 48, 63):
            def fill(M, N, K, k0=k0):
                A = torch.zeros(M, K, device=DEV); B = torch.zeros(N, K, device=DEV)
                A[:, k0] = torch.arange(1, M + 1, device=DEV).float() / 8
                B[:, k0] = torch.arange(1, N + 1, device=DEV).float() / 16
                return A, B
            out, ref = nt_case(128, 256, 64, 1, fill)
            report(f"NT onehot k0={k0}", out, ref)
        rnd = lambda M, N, K: (torch.randn(M, K, device=DEV), torch.randn(N, K, device=DEV))
        for (M, N, K, ns) in [(128, 256, 64, 1), (128, 256, 128, 1), (128, 256, 512, 1), (128, 256, 512, 3), (256, 512, 512, 3),
                              (1000, 2048, 512, 3), (200, 128, 2048, 3), (20000, 512, 512, 3)]:
            out, ref = nt_case(M, N, K, ns, rnd)
            report(f"NT rand M={M} N={N} K={K} nsplit={ns}", out, ref)
    except Exception:
        traceback.print_exc()
    try:
        for t0 in (0, 7, 8, 15, 16, 40, 63):
            def fill(T, Mo, No, t0=t0):
                A = torch.zeros(T, Mo, device=DEV); B = torch.zeros(T, No, device=DEV)
                A[t0] = torch.arange(1, Mo + 1, device=DEV).float() / 8
                B[t0] = torch.arange(1, No + 1, device=DEV).float() / 16
                return A, B
            out, ref = tn_case(64, 128, 256, 1, fill)
            report(f"TN onehot t0={t0}", out, ref)
        rnd = lambda T, Mo, No: (torch.randn(T, Mo, device=DEV), torch.randn(T, No, device=DEV))
        for (T, Mo, No, ns) in [(64, 128, 256, 1), (128, 128, 256, 1), (1000, 512, 512, 3), (5000, 128, 2048, 3), (64000, 512, 512, 3)]:
            out, ref = tn_case(T, Mo, No, ns, rnd)
            report(f"TN rand T={T} Mo={Mo} No={No} nsplit={ns}", out, ref)
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
