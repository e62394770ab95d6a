"""ctypes binding of ``libllp_b200.so`` (the C-ABI declared in ``include/llp_b200.h``).

PyTorch is only plumbing here: it owns device memory and the CUDA stream; every arithmetic
kernel on the hot path is a hand-written sm_100a kernel inside the shared library.  There is
no CPU fallback: importing works anywhere (so host-only tests can run), but the first compute
call on a machine without the library or without an sm_100 GPU raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libllp_b200.so")

LLP_F32, LLP_BF16 = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05, GEMM_TF32X3 = 0, 1, 2, 3


class GemmNtArgs(ctypes.Structure):
    """Mirror of ``llp_gemm_nt_args`` (include/llp_b200.h)."""

    _fields_ = [
        ("dtype", c_int), ("out_dtype", c_int), ("backend", c_int), ("relu", c_int),
        ("M", c_int64), ("N", c_int64), ("K1", c_int64), ("K2", c_int64),
        ("A1", c_void_p), ("lda1", c_int64), ("B1", c_void_p), ("ldb1", c_int64),
        ("A2", c_void_p), ("lda2", c_int64), ("B2", c_void_p), ("ldb2", c_int64),
        ("bias", c_void_p),
        ("addend", c_void_p), ("ldadd", c_int64),
        ("gate", c_void_p), ("ldgate", c_int64),
        ("gate_scale", c_float), ("dropout_p", c_float),
        ("seed", c_uint64), ("offset", c_uint64),
        ("rng_state", c_void_p),
        ("D", c_void_p), ("ldd", c_int64),
    ]


class EdgeMlpArgs(ctypes.Structure):
    """Mirror of ``llp_edge_mlp_args`` (include/llp_b200.h)."""

    _fields_ = [
        ("h", c_void_p), ("ldh", c_int64), ("u", c_void_p), ("v", c_void_p),
        ("M", c_int64), ("K", c_int64), ("N", c_int64),
        ("W1", c_void_p), ("ldw1", c_int64), ("bias1", c_void_p),
        ("relu", c_int), ("dropout_p", c_float), ("seed", c_uint64), ("offset", c_uint64), ("rng_state", c_void_p),
        ("z", c_void_p), ("ldz", c_int64), ("y", c_void_p), ("ldy", c_int64),
        ("w2", c_void_p), ("b2", c_void_p), ("prob", c_void_p),
    ]


class WeightDesc(ctypes.Structure):
    """Mirror of ``llp_weight_desc`` (include/llp_b200.h)."""

    _fields_ = [("src", c_void_p), ("rows", c_int64), ("cols", c_int64), ("dst", c_void_p), ("ld", c_int64),
                ("dst_t", c_void_p), ("ld_t", c_int64)]


# name -> (restype, argtypes); the single source for both binding and the "exports every symbol" test
PROTOTYPES = {
    "llp_version": (c_int, []),
    "llp_error_string": (c_char_p, [c_int]),
    "llp_device_supported": (c_int, []),
    "llp_launch_count": (c_int64, []),
    "llp_set_tuning": (None, [c_int, c_int]),
    "llp_debug_read": (c_int, [c_void_p, c_int]),
    "llp_csr_build_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "llp_csr_build": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_size_t, c_void_p]),
    "llp_spmm_num_chunks": (c_int64, [c_int64]),
    "llp_spmm_plan_ints": (c_int64, [c_int64]),
    "llp_spmm_plan": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "llp_spmm_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "llp_spmm": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                         c_int, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "llp_ipc_export": (c_int, [c_void_p, c_void_p, ctypes.POINTER(c_int64)]),
    "llp_ipc_open": (c_int, [c_void_p, c_int64, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p)]),
    "llp_ipc_close": (c_int, [c_void_p]),
    "llp_peer_barrier": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "llp_peer_gather_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p]),
    "llp_peer_mark_rows": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "llp_peer_reduce_rows": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "llp_spmm_peer": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_int64, c_int64,
                              c_int64, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "llp_gemm_nt": (c_int, [ctypes.POINTER(GemmNtArgs), c_void_p]),
    "llp_gemm_tn_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "llp_gemm_tn": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                            c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "llp_wgrad_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64]),
    "llp_wgrad": (c_int, [c_int, c_int, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                          c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "llp_colsum_workspace_bytes": (c_size_t, [c_int64]),
    "llp_colsum": (c_int, [c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "llp_cast2d": (c_int, [c_int, c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p]),
    "llp_weights_prep": (c_int, [c_int, ctypes.POINTER(WeightDesc), c_void_p]),
    "llp_add_act": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_float, c_uint64,
                            c_uint64, c_void_p, c_void_p, c_int64, c_void_p]),
    "llp_gate": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_int64,
                         c_void_p]),
    "llp_edge_hadamard": (c_int, [c_int, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                  c_void_p]),
    "llp_edge_plan_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "llp_edge_plan": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "llp_edge_hadamard_bwd_workspace_bytes": (c_size_t, [c_int64]),
    "llp_edge_hadamard_bwd": (c_int, [c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                      c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "llp_edge_mlp_supported": (c_int, [c_int64, c_int64]),
    "llp_edge_mlp_fused": (c_int, [ctypes.POINTER(EdgeMlpArgs), c_void_p]),
    "llp_score_head": (c_int, [c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "llp_score_head_bwd_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "llp_score_head_bwd": (c_int, [c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_float,
                                   c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "llp_loss_workspace_bytes": (c_size_t, [c_int64]),
    "llp_bce": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "llp_kd_d": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "llp_kd_r": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "llp_kd_fused_workspace_bytes": (c_size_t, [c_int64]),
    "llp_kd_fused": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_float, c_float, c_float, c_float, c_void_p, c_void_p,
                             c_void_p, c_void_p]),
    "llp_topk_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "llp_topk_desc": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "llp_count_greater": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "llp_auc_workspace_bytes": (c_size_t, [c_int64]),
    "llp_auc_pairs": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "llp_py_random_sample": (c_int, [c_void_p, c_uint64, c_int64, c_void_p]),
    "llp_negative_filter_workspace_bytes": (c_size_t, [c_int64]),
    "llp_negative_filter": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "llp_random_walk": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "llp_rng_advance": (c_int, [c_void_p, c_void_p]),
    "llp_clip_adam_workspace_bytes": (c_size_t, [c_int]),
    "llp_clip_adam": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, ctypes.POINTER(c_int64), c_int, c_float,
                              c_float, c_float, c_float, c_float, c_float, c_int64, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p]),
    "llp_sum": (c_int, [c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
}

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Load the shared library (build it first with ``__graft_entry__.build()`` or ``make -C csrc``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the LLP B200 path has no CPU/eager fallback. "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        # development aid: LLP_TUNING="key=value,key=value" applies llp_set_tuning knobs at load time (A/B runs of
        # experimental kernel variants through unmodified tests / benchmarks); results never depend on the knobs
        for item in filter(None, os.environ.get("LLP_TUNING", "").split(",")):
            k, _, v = item.partition("=")
            lib.llp_set_tuning(int(k), int(v))
        _lib = lib
    return _lib


def require_gpu() -> ctypes.CDLL:
    """Library + an sm_100 device, or RuntimeError (never a silent fallback)."""
    lib = load()
    if not torch.cuda.is_available():
        raise RuntimeError("linkless_link_prediction_b200 needs an sm_100 (B200) GPU; there is no CPU fallback")
    ok = lib.llp_device_supported()
    if ok != 1:
        raise RuntimeError(f"linkless_link_prediction_b200: current CUDA device is not sm_100 (code {ok})")
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().llp_error_string(rc)
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else rc} (code {rc})")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dtype_id(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return LLP_F32
    if dt == torch.bfloat16:
        return LLP_BF16
    raise RuntimeError(f"unsupported dtype {dt}: the LLP kernels take float32 or bfloat16")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def mat(t: torch.Tensor):
    """(pointer, leading dimension) of a 2-D row-major (possibly row-padded) CUDA tensor."""
    if t.dim() != 2 or (t.size(1) > 1 and t.stride(1) != 1):
        raise RuntimeError("expected a 2-D tensor with unit column stride")
    if not t.is_cuda:
        raise RuntimeError("expected a CUDA tensor: the LLP B200 path has no CPU fallback")
    ld = t.stride(0) if t.size(0) > 1 else max(t.stride(0), t.size(1))
    return t.data_ptr(), ld


def launch_count() -> int:
    return int(load().llp_launch_count())
