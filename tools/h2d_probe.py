"""Host→device bandwidth probe for the e2e leg: how fast does THIS box move one bench batch (131 MB of pinned fp32
features), and does it depend on which NUMA node the pinned pages live on?  Prints one JSON line per experiment.

    python tools/h2d_probe.py            # cuda:0
"""
import json
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madeleine_b200.utils import hostmem  # noqa: E402

NBYTES = 16 * 2 * 2000 * 512 * 4


def h2d_gbs(host, dev_buf, reps=10):
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    with torch.cuda.stream(s):
        for _ in range(2):
            dev_buf.copy_(host, non_blocking=True)
        for _ in range(reps):
            e0.record(s)
            dev_buf.copy_(host, non_blocking=True)
            e1.record(s)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
    return host.numel() * host.element_size() / best / 1e6


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    dev_buf = torch.empty(NBYTES // 4, device=dev)
    print(json.dumps({"cpus": os.cpu_count(), "affinity_before": sorted(os.sched_getaffinity(0)),
                      "numa_nodes": sh("ls -d /sys/devices/system/node/node* | wc -l"),
                      "node_cpulists": sh("cat /sys/devices/system/node/node*/cpulist"),
                      "gpu_local_cpus": hostmem.gpu_local_cpus(0)}))
    print(sh("nvidia-smi topo -m | head -12"))
    a = torch.empty(NBYTES // 4).pin_memory()
    a.normal_()
    print(json.dumps({"experiment": "pinned, default placement", "h2d_gbs": h2d_gbs(a, dev_buf)}))
    for node_cpus in sh("cat /sys/devices/system/node/node*/cpulist").split():
        try:
            cpus = hostmem.parse_cpulist(node_cpus)
            allowed = sorted(set(cpus) & os.sched_getaffinity(0))
            if not allowed:
                print(json.dumps({"experiment": f"node cpus {node_cpus}", "skipped": "not in this process's cpuset"}))
                continue
            saved = os.sched_getaffinity(0)
            os.sched_setaffinity(0, allowed)
            b = torch.empty(NBYTES // 4).pin_memory()
            b.normal_()
            print(json.dumps({"experiment": f"pinned, first touch from cpus {node_cpus}", "h2d_gbs": h2d_gbs(b, dev_buf)}))
            os.sched_setaffinity(0, saved)
            del b
        except OSError as e:
            print(json.dumps({"experiment": f"node cpus {node_cpus}", "error": str(e)}))
    t0 = time.perf_counter()
    ok = hostmem.bind_to_gpu(0)
    c = hostmem.pinned_empty((NBYTES // 4,), torch.float32)
    c.normal_()
    print(json.dumps({"experiment": "hostmem.bind_to_gpu(0) + hostmem.pinned_empty", "bound": ok, "h2d_gbs": h2d_gbs(c, dev_buf),
                      "setup_s": time.perf_counter() - t0, "affinity_after": sorted(os.sched_getaffinity(0))}))


if __name__ == "__main__":
    main()
