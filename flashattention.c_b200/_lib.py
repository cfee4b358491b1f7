"""ctypes binding of libfa_b200.so (include/fa_b200.h).  Fails loudly: if the library is missing or a call
returns an error there is NO fallback path — an exception is raised."""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
# FA_B200_LIB: another build of the same library (A/B and tracing variants under flashattention.c_b200/variants/)
LIB_PATH = Path(os.environ.get("FA_B200_LIB") or PKG_DIR / "libfa_b200.so")

FA_F32, FA_BF16, FA_F16 = 0, 1, 2
FA_IMPL_AUTO, FA_IMPL_TCGEN05, FA_IMPL_SIMT = 0, 1, 2
FA_FLAG_BATCH_INVARIANT = 1
FA_FLAG_PRECISE = 2

# every symbol include/fa_b200.h declares (tests check that the built library exports all of them)
EXPORTED_SYMBOLS = [
    "fa_forward", "fa_forward_ex", "fa_forward_packed_qkv", "fa_forward_packed_qkv_ex", "fa_forward_host", "fa_merge_partials", "fa_cast_f32", "fa_cast_f32_to_bf16",
    "run_flash_tiled_coarse", "run_flash_tiled_coarse_causal", "attention_forward6", "attention_forward",
    "fa_strerror", "fa_last_cuda_error", "fa_last_impl", "fa_version", "fa_launch_count",
    "fa_watchdog_info",
    "fa_p2p_alloc", "fa_p2p_open", "fa_p2p_close", "fa_p2p_free", "fa_copy_async", "fa_query_instance",
    "fa_backward",
]


class FaParams(ctypes.Structure):
    _fields_ = [
        ("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("v", ctypes.c_void_p), ("o", ctypes.c_void_p), ("lse", ctypes.c_void_p),
        ("batch", ctypes.c_int64), ("heads", ctypes.c_int64), ("n_q", ctypes.c_int64), ("n_k", ctypes.c_int64),
        ("head_dim", ctypes.c_int32), ("dtype", ctypes.c_int32), ("causal", ctypes.c_int32), ("o_f32", ctypes.c_int32),
        ("scale", ctypes.c_float),
        ("q_stride_b", ctypes.c_int64), ("q_stride_h", ctypes.c_int64), ("q_stride_n", ctypes.c_int64),
        ("k_stride_b", ctypes.c_int64), ("k_stride_h", ctypes.c_int64), ("k_stride_n", ctypes.c_int64),
        ("v_stride_b", ctypes.c_int64), ("v_stride_h", ctypes.c_int64), ("v_stride_n", ctypes.c_int64),
        ("o_stride_b", ctypes.c_int64), ("o_stride_h", ctypes.c_int64), ("o_stride_n", ctypes.c_int64),
        ("impl", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("o_acc", ctypes.c_void_p), ("lse_acc", ctypes.c_void_p),
        ("kv_heads", ctypes.c_int64),
    ]


class FaBwdParams(ctypes.Structure):
    _fields_ = (
        [(n, ctypes.c_void_p) for n in ("q", "k", "v", "o", "d_o", "lse", "dq", "dk", "dv")]
        + [(n, ctypes.c_int64) for n in ("batch", "heads", "kv_heads", "n_q", "n_k")]
        + [("head_dim", ctypes.c_int32), ("dtype", ctypes.c_int32), ("causal", ctypes.c_int32), ("scale", ctypes.c_float)]
        + [(f"{t}_stride_{a}", ctypes.c_int64) for t in ("q", "k", "v", "o", "do", "dq", "dk", "dv") for a in ("b", "h", "n")]
    )


class FaError(RuntimeError):
    pass


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FaError(
                f"{LIB_PATH} is missing: build it with `python flashattention.c_b200/build.py` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no fallback path."
            )
        L = ctypes.CDLL(str(LIB_PATH))
        vp, i64, i32, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float
        L.fa_forward.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, i32, f32, i32, i32, vp]
        L.fa_forward.restype = ctypes.c_int
        L.fa_forward_ex.argtypes = [ctypes.POINTER(FaParams), vp]
        L.fa_forward_ex.restype = ctypes.c_int
        L.fa_backward.argtypes = [ctypes.POINTER(FaBwdParams), vp]
        L.fa_backward.restype = ctypes.c_int
        L.fa_forward_packed_qkv.argtypes = [vp, vp, vp, i32, i32, i32, i32, f32, i32, vp]
        L.fa_forward_packed_qkv.restype = ctypes.c_int
        L.fa_forward_packed_qkv_ex.argtypes = [vp, vp, vp, i32, i32, i32, i32, f32, i32, i32, vp]
        L.fa_forward_packed_qkv_ex.restype = ctypes.c_int
        L.fa_forward_host.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i32, f32, i32, i32]
        L.fa_forward_host.restype = ctypes.c_int
        L.fa_merge_partials.argtypes = [vp, vp, vp, vp, i64, i32, vp]
        L.fa_merge_partials.restype = ctypes.c_int
        L.fa_cast_f32.argtypes = [vp, vp, i64, i32, vp]
        L.fa_cast_f32.restype = ctypes.c_int
        L.fa_cast_f32_to_bf16.argtypes = [vp, vp, i64, vp]
        L.fa_cast_f32_to_bf16.restype = ctypes.c_int
        L.fa_p2p_alloc.argtypes = [i64, ctypes.POINTER(vp), ctypes.c_char_p]
        L.fa_p2p_alloc.restype = ctypes.c_int
        L.fa_p2p_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
        L.fa_p2p_open.restype = ctypes.c_int
        L.fa_p2p_close.argtypes = [vp]
        L.fa_p2p_close.restype = ctypes.c_int
        L.fa_p2p_free.argtypes = [vp]
        L.fa_p2p_free.restype = ctypes.c_int
        L.fa_copy_async.argtypes = [vp, vp, i64, vp]
        L.fa_copy_async.restype = ctypes.c_int
        L.fa_query_instance.argtypes = [i32, i32]
        L.fa_query_instance.restype = ctypes.c_int
        L.fa_strerror.argtypes = [ctypes.c_int]
        L.fa_strerror.restype = ctypes.c_char_p
        L.fa_last_cuda_error.argtypes = []
        L.fa_last_cuda_error.restype = ctypes.c_char_p
        L.fa_last_impl.restype = ctypes.c_int
        L.fa_version.restype = ctypes.c_int
        L.fa_launch_count.restype = ctypes.c_int64
        L.fa_watchdog_info.argtypes = [ctypes.POINTER(ctypes.c_uint32)]
        L.fa_watchdog_info.restype = ctypes.c_int
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        L = lib()
        raise FaError(f"{what} failed ({rc}): {L.fa_strerror(rc).decode()} {L.fa_last_cuda_error().decode()}")
