"""Host-side staging memory for the H2D leg of the path (the reference's `data['feats'].to(device)`, Model.py:113).

On a two-socket B200 box pinned pages that land on the socket far from the GPU are copied at roughly half the PCIe
rate, which is what bounds the end-to-end number once the step itself takes ~5 ms.  `bind_to_gpu` narrows the calling
thread's CPU affinity to the GPU-local cores (NVML's view, falling back to sysfs) so that pinned buffers allocated and
first-touched afterwards live on the near node; `pinned_empty` allocates page-locked memory directly (no pageable
intermediate, unlike `tensor.pin_memory()`)."""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' → [0, 1, 2, 3, 8, 10, 11] (the sysfs cpulist format)."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-", 1)
            cpus.extend(range(int(lo), int(hi) + 1))
        else:
            cpus.append(int(part))
    return cpus


def _physical_index(index: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[index])
        except (ValueError, IndexError):
            return index
    return index


def gpu_local_cpus(index: int = 0) -> Optional[List[int]]:
    """CPUs on the GPU's NUMA node, or None when neither NVML nor sysfs can tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(_physical_index(index))
        words = (max(os.cpu_count() or 1, 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        if cpus:
            return cpus
    except Exception:  # noqa: BLE001
        pass
    try:
        props = torch.cuda.get_device_properties(index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            cpus = parse_cpulist(f.read())
        return cpus or None
    except Exception:  # noqa: BLE001
        return None


def bind_to_gpu(index: int = 0, cpus: Optional[Sequence[int]] = None) -> bool:
    """Restrict the calling thread to the GPU-local CPUs that its cpuset allows.  Returns True when the affinity was
    narrowed (or already was local), False when nothing is known or allowed — never raises: staging still works, only
    slower."""
    local = list(cpus) if cpus is not None else gpu_local_cpus(index)
    if not local:
        return False
    try:
        allowed = sorted(set(local) & os.sched_getaffinity(0))
        if not allowed:
            return False
        os.sched_setaffinity(0, allowed)
        return True
    except (OSError, AttributeError):
        return False


def pinned_empty(shape, dtype=torch.float32) -> torch.Tensor:
    """Page-locked host tensor; the pages are touched here so they are placed by the CURRENT thread's affinity."""
    t = torch.empty(shape, dtype=dtype, pin_memory=torch.cuda.is_available())
    if t.numel():
        t.view(-1)[:: max(1, 4096 // t.element_size())] = 0      # first touch, one write per page
    return t
