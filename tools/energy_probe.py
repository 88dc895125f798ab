"""Board power (NVML) and SM clock of single kernels of the bench step looped for ~1.5 s each: where the joules of a step go.
The step runs at the board's power limit (tools/ab_probe.py --power), so energy per launch, not idle time, sets its length."""
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madeleine_b200 import ops  # noqa: E402

import pynvml  # noqa: E402

dev = torch.device("cuda", 0)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
M = 64000


def looped(name, fn, seconds=1.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    e0.record(); [fn() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 5
    n = max(10, int(seconds * 1e3 / per))
    w, mhz, stop = [], [], threading.Event()

    def poll():
        while not stop.is_set():
            w.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
            mhz.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            time.sleep(0.005)

    t = threading.Thread(target=poll, daemon=True)
    e0.record()
    t.start()
    for i in range(n):
        fn()
        if i % 64 == 63:
            torch.cuda.synchronize()        # keep the launch queue short so that the sampler sees the steady state
    e1.record()
    torch.cuda.synchronize()
    stop.set(); t.join()
    ms = e0.elapsed_time(e1) / n
    half = w[len(w) // 2:]                 # second half: the power controller has settled
    pw = sorted(half)[len(half) // 2]
    clk = sorted(mhz[len(mhz) // 2:])[len(mhz) // 4]
    print(json.dumps({"kernel": name, "us_per_launch": round(ms * 1e3, 1), "power_w": round(pw, 1), "sm_mhz": clk,
                      "mJ_per_launch": round(pw * ms, 2)}))
    time.sleep(1.0)


time.sleep(1.0)
idle = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
print(json.dumps({"idle_power_w": idle, "limit_w": pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3}))

x512 = ops.split_planes(torch.randn(M, 512, device=dev), 2)
x2048 = ops.split_planes(torch.randn(M, 2048, device=dev), 2)
w3 = ops.split_planes(torch.randn(2048, 512, device=dev) / 22, 2)
bw3 = (w3, 2048, 512, 2048 * 512)
out3 = torch.empty(M, 2048, device=dev)
looped("gemm_nt L3 forward (64000 x 512 -> 2048, 3-pass)", lambda: ops.gemm_nt(x512, 512, bw3, 2048, 3, out=out3))
d4096 = ops.split_planes(torch.randn(M, 4096, device=dev), 2)
dw = torch.zeros(4096, 512, device=dev)
looped("gemm_tn attention wgrad (4 x 1024 x 512 over 64000 tokens, 3-pass)", lambda: ops.gemm_tn_accum(d4096, x2048, dw, 3, 1024, 512))
z = torch.randn(M, 2048, device=dev)
g, b = torch.ones(2048, device=dev), torch.zeros(2048, device=dev)
looped("ln_gelu_fwd<2048>", lambda: ops.ln_gelu_fwd(z, g, b, 2, 0.1, 1, 3))
planes, mean, rstd = ops.ln_gelu_fwd(z, g, b, 2, 0.1, 1, 3)
dh = torch.randn(M, 2048, device=dev)
dg, db_, dbias = (torch.zeros(2048, device=dev) for _ in range(3))
looped("ln_gelu_bwd<2048> (no pooling term)", lambda: ops.ln_gelu_bwd(z, g, b, mean, rstd, dh, None, [], 4, 2, 0.1, 1, 3, dg, db_, dbias))
big_a = torch.empty(1 << 28, dtype=torch.float32, device=dev)
big_b = torch.empty_like(big_a)
looped("torch copy 1 GiB -> 1 GiB (HBM stream)", lambda: big_b.copy_(big_a))
