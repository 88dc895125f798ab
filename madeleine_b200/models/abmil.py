"""Gated-attention head (drop-in for madeleine/models/abmil.py:8-68).

``BatchedABMIL`` keeps the reference's constructor, submodule names (``attention_a.0``, ``attention_b.0``,
``attention_c`` → identical checkpoint keys) and ``forward(x, return_raw_attention)`` contract.  Inside
``ABMILEmbedder`` the four heads are evaluated together by one tcgen05 GEMM with a fused tanh·sigmoid·w_c epilogue
(``mdl_gemm_gated``); a head called on its own goes through the same kernel with a single-head weight pack.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from .._lib import call, require_cuda, stream_ptr

_ACTIVATIONS = ("softmax", "leaky_relu", "relu", "sigmoid")


class BatchedABMIL(nn.Module):
    def __init__(self, input_dim=1024, hidden_dim=256, dropout=False, n_classes=1, n_heads=1, activation="softmax"):
        super().__init__()
        self.activation = activation
        self.input_dim, self.hidden_dim, self.n_classes = input_dim, hidden_dim, n_classes
        gate_a = [nn.Linear(input_dim, hidden_dim), nn.Tanh()]
        gate_b = [nn.Linear(input_dim, hidden_dim), nn.Sigmoid()]
        if dropout:
            gate_a.append(nn.Dropout(0.25))
            gate_b.append(nn.Dropout(0.25))
        self.attention_a = nn.Sequential(*gate_a)
        self.attention_b = nn.Sequential(*gate_b)
        self.attention_c = nn.Linear(hidden_dim, n_classes)
        self.gate_dropout = 0.25 if dropout else 0.0

    def gate_parameters(self):
        """(Wa, ba, Wb, bb, wc, bc) in the order the weight packer expects."""
        return [self.attention_a[0].weight, self.attention_a[0].bias, self.attention_b[0].weight, self.attention_b[0].bias,
                self.attention_c.weight, self.attention_c.bias]

    def forward(self, x, return_raw_attention=False):
        """x [bs, tokens, input_dim] → activated attention [bs, tokens, 1] (and the raw logits).

        Inference-only when used on its own (the fused training path lives in ABMILEmbedder)."""
        if self.activation not in _ACTIVATIONS:
            raise NotImplementedError("Activation not implemented.")
        require_cuda(x, "BatchedABMIL input")
        if self.input_dim != ops.HID or self.hidden_dim != ops.GATE or self.n_classes != 1:
            raise NotImplementedError("madeleine_b200 kernels are built for input_dim=512, hidden_dim=512, n_classes=1")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and x.requires_grad:
            raise NotImplementedError("standalone BatchedABMIL is forward-only; train through ABMILEmbedder")
        bs, T, D = x.shape
        M = bs * T
        dev = x.device
        st = stream_ptr(dev)
        precision = ops.resolve_precision(None)
        nsplit, npl = ops._nsplit(precision)
        xp = ops.split_planes(x.reshape(M, D).contiguous().float(), npl)
        wa, ba, wb, bb, wc, bc = [p.detach().float() for p in self.gate_parameters()]
        # packed gate rows: 4 groups of [128 Wa rows | 128 Wb rows]
        packed = torch.cat([torch.cat([wa[g * 128:(g + 1) * 128], wb[g * 128:(g + 1) * 128]]) for g in range(4)]).contiguous()
        wp = ops.split_planes(packed, npl)
        logits = torch.empty(M, 1, dtype=torch.float32, device=dev)
        call("mdl_gemm_gated", xp, M, D, D, M * D, wp, wp.shape[1] * wp.shape[2], M, 1, nsplit, ba.contiguous(), bb.contiguous(),
             wc.reshape(-1).contiguous(), bc.contiguous(), logits, None, None, 0.0, 0, st)
        A = logits.view(bs, T, 1)
        if self.activation == "softmax":
            act = torch.softmax(A, dim=1)
        elif self.activation == "leaky_relu":
            act = torch.nn.functional.leaky_relu(A)
        elif self.activation == "relu":
            act = torch.relu(A)
        else:
            act = torch.sigmoid(A)
        if return_raw_attention:
            return act, A
        return act
