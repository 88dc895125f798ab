"""Argument marshalling for the native step executor (``mdl_encoder_fwd`` / ``mdl_encoder_bwd``, csrc/executor.cu).

The executor takes three flat arrays indexed by the enums of include/madeleine_b200.h; the name lists below mirror those
enums (``tests/test_abi.py`` parses the header and checks the order; ``check_abi`` compares the counts with the loaded
library).  Everything that depends only on the parameter layout and the precision is prepared once per ``PackSpec``; a
call fills in the per-batch sizes, flags and pointers."""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

ENC_I = ['M', 'R', 'D_IN', 'D_IN_TOTAL', 'SE_DIM', 'N_HEADS', 'NSPLIT_FWD', 'NPL_FWD', 'NSPLIT_BWD', 'NPL_BWD', 'ACT_BF16',
         'ACTIVATION', 'KEEP', 'WANT_TOKENS', 'WANT_PROJECTOR', 'WANT_REF', 'N_VIEW_TOK', 'R2', 'N_SEL', 'SEED', 'PHASE',
         'BF_NUMEL', 'BF_W1', 'BF_W2', 'BF_W2T', 'BF_W3', 'BF_W3T', 'BF_WAB', 'BF_WABT', 'BF_TP', 'BF_TPT',
         'F32_B1', 'F32_G1', 'F32_BE1', 'F32_B2', 'F32_G2', 'F32_BE2', 'F32_B3', 'F32_G3', 'F32_BE3', 'F32_BA', 'F32_BB',
         'F32_WC', 'F32_BC', 'F32_BTP', 'F32_WP', 'F32_BP',
         'GR_NUMEL', 'GR_W1', 'GR_W2', 'GR_W3', 'GR_WAB', 'GR_TP', 'GR_B1', 'GR_G1', 'GR_BE1', 'GR_B2', 'GR_G2', 'GR_BE2',
         'GR_B3', 'GR_G3', 'GR_BE3', 'GR_BA', 'GR_BB', 'GR_WC', 'GR_BC', 'GR_BTP', 'GR_WP', 'GR_BP',
         'MASTER_NUMEL', 'MASTER_PRE0W', 'MASTER_EMB', 'GR_N', 'GR_N_EARLY', 'GR_N_LATE', 'PLANES_F16']
ENC_F = ['P_PRE', 'P_GATE']
ENC_P = ['STREAM', 'X', 'CU', 'CODES', 'MASTER', 'WBF', 'WF32', 'ARENA', 'SLIDE_HM', 'SLIDE', 'LOGITS', 'TOKENS', 'REF',
         'VIEW_TOK_IDX', 'VIEW_CU', 'VIEW_ROW2SEG', 'TOKEN_ROWS', 'TOKEN_SEL_OF_ROW', 'BWD_ARENA', 'D_SLIDE', 'D_LOGITS',
         'D_TOKENS', 'D_REF_HM', 'GMASTER', 'GR_POS', 'GR_DST', 'GR_POS_EARLY', 'GR_DST_EARLY', 'GR_POS_LATE', 'GR_DST_LATE']
PROF_TAGS = ['other', 'gemm_nt', 'gemm_gated', 'gemm_tn', 'ln_fwd', 'ln_bwd', 'gate_bwd', 'pool_weights', 'pool_fwd', 'pool_bwd',
             'skinny']

I = {n: i for i, n in enumerate(ENC_I)}
F = {n: i for i, n in enumerate(ENC_F)}
P = {n: i for i, n in enumerate(ENC_P)}
ABI = len(ENC_I) * 10000 + len(ENC_F) * 1000 + len(ENC_P)

_IArr = ctypes.c_longlong * len(ENC_I)
_FArr = ctypes.c_double * len(ENC_F)
_PArr = ctypes.c_void_p * len(ENC_P)

# names of the packed segments (ops.PackSpec) behind the offset slots
_BF = {'BF_W1': 'w1', 'BF_W2': 'w2', 'BF_W2T': 'w2T', 'BF_W3': 'w3', 'BF_W3T': 'w3T', 'BF_WAB': 'wab', 'BF_WABT': 'wabT',
       'BF_TP': 'tp', 'BF_TPT': 'tpT'}
_VEC = ['b1', 'g1', 'be1', 'b2', 'g2', 'be2', 'b3', 'g3', 'be3', 'ba', 'bb', 'wc', 'bc', 'btp', 'wp', 'bp']
_GRM = {'GR_W1': 'w1', 'GR_W2': 'w2', 'GR_W3': 'w3', 'GR_WAB': 'wab', 'GR_TP': 'tp'}


def static_iparams(spec) -> List[int]:
    """The layout-dependent entries of the integer array for one PackSpec (copied and completed per call)."""
    ip = [0] * len(ENC_I)
    ip[I['BF_NUMEL']] = spec.bf_numel
    for slot, name in _BF.items():
        ip[I[slot]] = spec.bf_segs[name].off
    for name in _VEC:
        ip[I['F32_' + name.upper()]] = spec.f32_segs[name].off
        ip[I['GR_' + name.upper()]] = spec.gr_segs[name].off
    ip[I['GR_NUMEL']] = spec.gr_numel
    for slot, name in _GRM.items():
        ip[I[slot]] = spec.gr_segs[name].off
    ip[I['MASTER_NUMEL']] = spec.master_numel
    ip[I['MASTER_PRE0W']] = spec.off("pre0.w")
    ip[I['MASTER_EMB']] = spec.off("emb.w") if spec.P["emb.w"] < len(spec.param_offsets) else 0
    ip[I['GR_N']] = spec.gr_pos.numel()
    ip[I['GR_N_EARLY']] = spec.gr_pos_early.numel()
    ip[I['GR_N_LATE']] = spec.gr_pos_late.numel()
    ip[I['N_HEADS']] = spec.n_heads
    ip[I['D_IN_TOTAL']] = spec.d_in_total
    return ip


def iarr(values: List[int]):
    return _IArr(*values)


def farr(p_pre: float, p_gate: float):
    return _FArr(float(p_pre), float(p_gate))


def parr(ptrs: Dict[str, Optional[object]]):
    """{slot name: tensor | int | None} -> void*[]; tensors contribute their data pointer."""
    arr = _PArr()
    for name, v in ptrs.items():
        if v is None:
            continue
        arr[P[name]] = v.data_ptr() if isinstance(v, torch.Tensor) else int(v)
    return arr


def check_abi(lib) -> None:
    got = lib.mdl_encoder_abi()
    if got != ABI:
        raise RuntimeError(f"madeleine_b200: executor ABI mismatch (library {got}, python {ABI}); rebuild with "
                           "`python -m madeleine_b200.build --force`")
