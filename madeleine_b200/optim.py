"""Fused AdamW for the train-step caller: one kernel launch per optimiser step for the whole model
instead of torch.optim.AdamW's per-tensor (or foreach) launches.  Drop-in for ``optim.AdamW(model.parameters(), lr=lr)``
(reference: madeleine/utils/setup_components.py:194-196) and compatible with torch LR schedulers (reads
``param_groups[i]['lr']`` every step)."""
from __future__ import annotations

import ctypes

import torch

from ._lib import call, stream_ptr


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._max = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._max is None:
            self._max = call("mdl_adamw_max_tensors")
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            beta1, beta2 = group["betas"]
            for p in ps:
                if p.dtype != torch.float32 or not p.is_cuda:
                    raise RuntimeError("FusedAdamW handles fp32 CUDA parameters only (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
            lr = float(group["lr"])
            # bias correction follows every tensor's OWN step count (torch.optim.AdamW tracks it per parameter): tensors that
            # first received a gradient later than the others form their own launch
            by_step = {}
            for p in ps:
                by_step.setdefault(self.state[p]["step"], []).append(p)
            for step, same in by_step.items():
                for i in range(0, len(same), self._max):
                    chunk = same[i:i + self._max]
                    n = len(chunk)
                    grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in chunk]
                    arr = ctypes.c_void_p * n
                    P = arr(*[p.data_ptr() for p in chunk])
                    G = arr(*[g.data_ptr() for g in grads])
                    M = arr(*[self.state[p]["exp_avg"].data_ptr() for p in chunk])
                    V = arr(*[self.state[p]["exp_avg_sq"].data_ptr() for p in chunk])
                    N = (ctypes.c_longlong * n)(*[p.numel() for p in chunk])
                    with torch.cuda.device(chunk[0].device):
                        call("mdl_adamw_step", n, ctypes.cast(P, ctypes.c_void_p), ctypes.cast(G, ctypes.c_void_p),
                             ctypes.cast(M, ctypes.c_void_p), ctypes.cast(V, ctypes.c_void_p), ctypes.cast(N, ctypes.c_void_p),
                             lr, beta1, beta2, group["eps"], group["weight_decay"], step, 1.0, stream_ptr(chunk[0].device))
            # the kernel wrote the parameters behind autograd's back: bump their version counters so that everything keyed
            # on them sees the update (the encoder re-packs its bf16 operand planes when a parameter's version changes)
            for p in ps:
                torch.autograd.graph.increment_version(p)
        return loss
