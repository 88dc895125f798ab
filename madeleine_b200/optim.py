"""Fused AdamW for the train-step caller: one kernel launch per optimiser step for the whole model
instead of torch.optim.AdamW's per-tensor (or foreach) launches.  Drop-in for ``optim.AdamW(model.parameters(), lr=lr)``
(reference: madeleine/utils/setup_components.py:194-196) and compatible with torch LR schedulers (reads
``param_groups[i]['lr']`` every step).

When the parameters are slices of one flat buffer (the encoder flattens them on first use, models/Model.py::_flat_master)
and their gradients are slices of the encoder's flat gradient, neighbouring tensors are merged into runs: the whole model
is then 1-3 "tensors" for the kernel and the per-step host work is a handful of pointer comparisons."""
from __future__ import annotations

import ctypes

import torch

from ._lib import call, stream_ptr


def _flat_like(ps):
    """Zero state for ``ps`` (any order): parameters that are back-to-back slices of one buffer get state that is back to
    back too (one flat allocation per run of neighbours), so that runs of parameters stay runs of state."""
    order = sorted(range(len(ps)), key=lambda i: ps[i].data_ptr())
    out = [None] * len(ps)
    i = 0
    while i < len(order):
        j = i
        while (j + 1 < len(order) and ps[order[j]].is_contiguous()
               and ps[order[j + 1]].data_ptr() == ps[order[j]].data_ptr() + 4 * ps[order[j]].numel()):
            j += 1
        run = [ps[k] for k in order[i:j + 1]]
        if len(run) == 1 or not run[-1].is_contiguous():
            for k in order[i:j + 1]:
                out[k] = torch.zeros_like(ps[k], memory_format=torch.contiguous_format)
        else:
            flat = torch.zeros(sum(p.numel() for p in run), dtype=torch.float32, device=run[0].device)
            o = 0
            for k in order[i:j + 1]:
                out[k] = flat[o:o + ps[k].numel()].view(ps[k].shape)
                o += ps[k].numel()
        i = j + 1
    return out


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._max = None
        self.last_launch_tensors = 0        # tensors (after merging runs) handed to the kernel by the last step()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._max is None:
            self._max = call("mdl_adamw_max_tensors")
        self.last_launch_tensors = 0
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            beta1, beta2 = group["betas"]
            fresh = [p for p in ps if not self.state[p]]
            for p in ps:
                if p.dtype != torch.float32 or not p.is_cuda:
                    raise RuntimeError("FusedAdamW handles fp32 CUDA parameters only (no CPU fallback)")
            if fresh:
                for p, m, v in zip(fresh, _flat_like(fresh), _flat_like(fresh)):
                    st = self.state[p]
                    st["step"] = 0
                    st["exp_avg"] = m
                    st["exp_avg_sq"] = v
            lr = float(group["lr"])
            # bias correction follows every tensor's OWN step count (torch.optim.AdamW tracks it per parameter): tensors that
            # first received a gradient later than the others form their own launch
            by_step = {}
            for p in ps:
                st = self.state[p]
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
            for step, same in by_step.items():
                # merge neighbours whose parameter, gradient and both moments continue the previous tensor's memory
                runs, keep = [], []             # [param ptr, grad ptr, m ptr, v ptr, numel]; keep: temporaries alive until launch
                for p in sorted(same, key=lambda t: t.data_ptr()):
                    g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                    keep.append(g)
                    st = self.state[p]
                    cur = [p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()]
                    if runs:
                        last = runs[-1]
                        nb = 4 * last[4]
                        if cur[0] == last[0] + nb and cur[1] == last[1] + nb and cur[2] == last[2] + nb and cur[3] == last[3] + nb:
                            last[4] += cur[4]
                            continue
                    runs.append(cur)
                self.last_launch_tensors += len(runs)
                dev = same[0].device
                for i in range(0, len(runs), self._max):
                    chunk = runs[i:i + self._max]
                    n = len(chunk)
                    arr = ctypes.c_void_p * n
                    P, G, M, V = (arr(*[r[k] for r in chunk]) for k in range(4))
                    N = (ctypes.c_longlong * n)(*[r[4] for r in chunk])
                    with torch.cuda.device(dev):
                        call("mdl_adamw_step", n, P, G, M, V, N, lr, beta1, beta2, group["eps"], group["weight_decay"], step, 1.0,
                             stream_ptr(dev))
            # the kernel wrote the parameters behind autograd's back: bump their version counters so that everything keyed
            # on them sees the update (the encoder re-packs its bf16 operand planes when a parameter's version changes)
            torch.autograd.graph.increment_version(ps)
        return loss
