"""ctypes binding of libmadeleine_b200.so (C ABI declared in include/madeleine_b200.h).

There is no CPU or PyTorch fallback: if the shared library cannot be loaded, or a call fails, a RuntimeError is
raised.  The library is built in-tree by ``madeleine_b200.build`` (nvcc, sm_100a) and shipped next to this file.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmadeleine_b200.so")

_lock = threading.Lock()
_lib = None

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_ll = ctypes.c_longlong
c_f = ctypes.c_float
c_ull = ctypes.c_ulonglong
c_u = ctypes.c_uint

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "mdl_version": [],
    "mdl_built_arch": [],
    "mdl_split_planes": [c_p, c_ll, c_i, c_ll, c_p, c_ll, c_i, c_p],
    "mdl_gather_split": [c_p, c_p, c_ll, c_p, c_ll, c_i, c_p],
    "mdl_gather_f32": [c_p, c_p, c_ll, c_p, c_p],
    "mdl_scatter_f32": [c_p, c_p, c_ll, c_p, c_i, c_p],
    "mdl_row2bag": [c_p, c_i, c_p, c_ll, c_p],
    "mdl_gemm_nt": [c_p, c_ll, c_ll, c_ll, c_ll, c_p, c_ll, c_ll, c_ll, c_ll, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i,
                    c_p, c_p, c_p, c_i, c_p],
    "mdl_gemm_gated": [c_p, c_ll, c_ll, c_ll, c_ll, c_p, c_ll, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_f,
                       c_ull, c_p],
    "mdl_gemm_tn_accum": [c_p, c_ll, c_ll, c_ll, c_p, c_ll, c_ll, c_ll, c_ll, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "mdl_gemm_debug_flags": [c_i],
    "mdl_gemm_nt_simt": [c_p, c_ll, c_ll, c_p, c_ll, c_ll, c_p, c_ll, c_i, c_i, c_i, c_i, c_p],
    "mdl_gemm_tn_simt": [c_p, c_ll, c_ll, c_p, c_ll, c_ll, c_ll, c_p, c_ll, c_i, c_i, c_i, c_p],
    "mdl_ln_gelu_fwd": [c_p, c_ll, c_i, c_p, c_p, c_f, c_f, c_ull, c_u, c_p, c_ll, c_i, c_p, c_p, c_i, c_p],
    "mdl_ln_gelu_bwd": [c_p, c_ll, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_f, c_ull, c_u,
                        c_p, c_ll, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_p],
    "mdl_gate_bwd": [c_p, c_p, c_p, c_p, c_ll, c_i, c_f, c_ull, c_p, c_ll, c_i, c_p, c_p, c_p, c_p, c_p],
    "mdl_pool_tsplit": [c_i, c_ll, c_i, c_i],
    "mdl_pool_workspace_bytes": [c_i, c_i, c_i, c_i],
    "mdl_pool_weights": [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_i, c_i, c_p, c_p],
    "mdl_pool_fwd": [c_p, c_ll, c_i, c_p, c_p, c_p, c_i, c_ll, c_i, c_i, c_p, c_i, c_p, c_p],
    "mdl_pool_bwd_dlogit": [c_p, c_ll, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_ll, c_i, c_i, c_p, c_i, c_p, c_i, c_i, c_p],
    "mdl_planes_to_ref_order": [c_p, c_ll, c_i, c_ll, c_i, c_i, c_p, c_p],
    "mdl_skinny_linear_fwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p],
    "mdl_skinny_linear_bwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p],
    "mdl_stain_rowbias": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "mdl_gather_rows_planes": [c_p, c_ll, c_i, c_i, c_p, c_ll, c_p, c_ll, c_p],
    "mdl_bag_colsum_planes": [c_p, c_ll, c_i, c_i, c_p, c_i, c_p, c_p],
    "mdl_stain_rowbias_bwd": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p],
    "mdl_colsum_f32": [c_p, c_ll, c_i, c_p, c_p],
    "mdl_infonce_fwd": [c_p, c_p, c_i, c_i, c_f, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "mdl_infonce_bwd": [c_p, c_p, c_i, c_i, c_f, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "mdl_infonce_rows_workspace_floats": [c_i],
    "mdl_infonce_rows_fwd": [c_p, c_ll, c_p, c_p, c_i, c_i, c_f, c_i, c_p, c_p, c_p, c_p],
    "mdl_infonce_rows_bwd": [c_p, c_ll, c_p, c_p, c_i, c_i, c_f, c_i, c_p, c_p, c_p, c_p],
    "mdl_adamw_max_tensors": [],
    "mdl_adamw_step": [c_i, c_p, c_p, c_p, c_p, c_p, c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_p],
    "mdl_sample_gather_f32": [c_p, c_p, c_p, c_i, c_i, c_i, c_ull, c_p, c_p, c_p],
    "mdl_got_workspace_bytes": [c_i, c_i, c_i],
    "mdl_got_max_tokens": [],
    "mdl_got_force_big": [c_i],
    "mdl_got_extrema": [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p],
    "mdl_got_fwd_bwd": [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "mdl_got_main": [c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p],
    "mdl_got_finish": [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "mdl_encoder_abi": [],
    "mdl_encoder_fwd_arena_bytes": [c_p],
    "mdl_encoder_bwd_arena_bytes": [c_p],
    "mdl_encoder_fwd": [c_p, c_p, c_p],
    "mdl_encoder_bwd": [c_p, c_p, c_p],
    "mdl_permute_f32": [c_p, c_p, c_p, c_ll, c_p, c_p],
    "mdl_gram_f64": [c_p, c_ll, c_i, c_p, c_p],
    "mdl_peer_allreduce_f32": [c_p, c_p, c_i, c_i, c_ll, c_ll, c_p, c_u, c_i, c_p],
    "mdl_peer_allgather_f32": [c_p, c_p, c_i, c_i, c_ll, c_ll, c_p, c_u, c_i, c_p],
    "mdl_executor_launches": [c_i],
    "mdl_profile_enable": [c_i],
    "mdl_profile_read": [c_p, c_p, c_i],
    "mdl_set_pdl": [c_i],
}
_RESTYPES = {"mdl_got_workspace_bytes": c_ll, "mdl_pool_workspace_bytes": c_ll, "mdl_encoder_fwd_arena_bytes": c_ll,
             "mdl_encoder_bwd_arena_bytes": c_ll, "mdl_executor_launches": c_ll, "mdl_infonce_rows_workspace_floats": c_ll}
# functions that return a value rather than a status code
_VALUE_FUNCS = {"mdl_version", "mdl_built_arch", "mdl_got_workspace_bytes", "mdl_got_max_tokens", "mdl_pool_tsplit",
                "mdl_pool_workspace_bytes", "mdl_adamw_max_tensors", "mdl_encoder_abi", "mdl_encoder_fwd_arena_bytes",
                "mdl_encoder_bwd_arena_bytes", "mdl_executor_launches", "mdl_profile_enable", "mdl_profile_read", "mdl_set_pdl",
                "mdl_infonce_rows_workspace_floats"}


def exported_symbols():
    """Every symbol include/madeleine_b200.h declares (kept in sync by tests/test_abi.py)."""
    return ["mdl_last_error", *SIGNATURES.keys()]


def load():
    """Load (building first if the .so is absent and nvcc is present). Raises RuntimeError on failure."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            try:
                from . import build as _build
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(
                    f"madeleine_b200: CUDA library {LIB_PATH} is missing and could not be built ({e}). "
                    "There is no CPU fallback; run `python -m madeleine_b200.build`.") from e
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise RuntimeError(f"madeleine_b200: cannot load {LIB_PATH}: {e}. There is no CPU fallback.") from e
        lib.mdl_last_error.restype = ctypes.c_char_p
        lib.mdl_last_error.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_i)
        from . import executor as _executor
        _executor.check_abi(lib)
        _lib = lib
    return _lib


# kernels launched per entry point (memsets not counted) — used for the bench's `gpu_launches` claim
LAUNCHES = {
    "mdl_split_planes": 1, "mdl_gather_split": 1, "mdl_gather_f32": 1, "mdl_scatter_f32": 1, "mdl_row2bag": 1,
    "mdl_gemm_nt": 1, "mdl_gemm_gated": 1, "mdl_gemm_tn_accum": 1, "mdl_gemm_nt_simt": 1, "mdl_gemm_tn_simt": 1,
    "mdl_ln_gelu_fwd": 1, "mdl_ln_gelu_bwd": 1, "mdl_gate_bwd": 1, "mdl_pool_weights": 1, "mdl_pool_fwd": 1, "mdl_pool_bwd_dlogit": 1,
    "mdl_planes_to_ref_order": 1, "mdl_skinny_linear_fwd": 1, "mdl_skinny_linear_bwd": 2, "mdl_stain_rowbias": 1,
    "mdl_bag_colsum_planes": 1, "mdl_gather_rows_planes": 1, "mdl_stain_rowbias_bwd": 1, "mdl_colsum_f32": 1, "mdl_infonce_fwd": 4, "mdl_infonce_bwd": 2, "mdl_infonce_rows_fwd": 4, "mdl_infonce_rows_bwd": 2,
    "mdl_got_extrema": 2, "mdl_got_fwd_bwd": 2, "mdl_got_main": 2, "mdl_got_finish": 1, "mdl_adamw_step": 1, "mdl_sample_gather_f32": 1, "mdl_permute_f32": 1, "mdl_gram_f64": 1, "mdl_peer_allreduce_f32": 1, "mdl_peer_allgather_f32": 1,
}
# mdl_encoder_fwd / mdl_encoder_bwd issue their launches natively; they are counted by the library (mdl_executor_launches)


def native_launches(reset: bool = False) -> int:
    """Kernel entry points issued by the native executor on this thread since the last reset."""
    return int(load().mdl_executor_launches(1 if reset else 0))
launch_count = [0]
# optional per-kernel device timing: {name: [(start_event, end_event), ...]} filled when `timed_kernels` is a set of names
# (or the string "all"); the native executor's launches are timed by the library itself (mdl_profile_*)
timed_kernels = None
kernel_events = {}
_PROF_NAMES = ["other", "mdl_gemm_nt", "mdl_gemm_gated", "mdl_gemm_tn_accum", "mdl_ln_gelu_fwd", "mdl_ln_gelu_bwd", "mdl_gate_bwd",
               "mdl_pool_weights", "mdl_pool_fwd", "mdl_pool_bwd_dlogit", "mdl_skinny_linear"]


def start_timing(names="all"):
    """Time every launch: entry points called from Python whose name is in ``names`` (or all of them), and every launch the
    native executor issues.  Adds event records around each launch — use for attribution, not for the headline number."""
    global timed_kernels
    kernel_events.clear()
    timed_kernels = names
    lib = load()
    lib.mdl_profile_read(None, None, 0)          # drop stale records
    mask = 1
    if names != "all":
        mask = sum(1 << i for i, n in enumerate(_PROF_NAMES) if n in names and i > 0)
    lib.mdl_profile_enable(mask)


def stop_timing():
    """-> {entry point name: [milliseconds per launch, in launch order]}; synchronises the device."""
    global timed_kernels
    lib = load()
    timed_kernels = None
    lib.mdl_profile_enable(0)
    torch.cuda.synchronize()
    out = {name: [a.elapsed_time(b) for a, b in ev] for name, ev in kernel_events.items()}
    kernel_events.clear()
    cap = 1 << 16
    tags = (ctypes.c_int * cap)()
    ms = (ctypes.c_float * cap)()
    n = min(lib.mdl_profile_read(ctypes.cast(tags, ctypes.c_void_p), ctypes.cast(ms, ctypes.c_void_p), cap), cap)
    for i in range(n):
        out.setdefault(_PROF_NAMES[tags[i]], []).append(ms[i])
    return out


def _conv(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    if isinstance(a, ctypes.Array):
        return ctypes.addressof(a)
    return a


def call(name, *args):
    """Call a status-returning entry point with tensors converted to device pointers; raise on failure."""
    lib = load()
    if name in _VALUE_FUNCS:
        return getattr(lib, name)(*[_conv(a) for a in args])
    launch_count[0] += LAUNCHES.get(name, 0)
    if timed_kernels is not None and (timed_kernels == "all" or name in timed_kernels):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*[_conv(a) for a in args])
        e1.record()
        kernel_events.setdefault(name, []).append((e0, e1))
    else:
        rc = getattr(lib, name)(*[_conv(a) for a in args])
    if rc != 0:
        msg = lib.mdl_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"madeleine_b200.{name} failed (code {rc}): {msg}")
    return 0


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"madeleine_b200: {what} must live on a CUDA device (got {t.device}). The B200 path has no CPU fallback; "
            "use the reference implementation for CPU execution.")
