"""Python face of the CPU oracle (oracle/fa_oracle.c) plus an independent numpy evaluation.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package.

  tiled(q, k, v, scale, causal)      fp32, the reference kernel's recurrence in its tile order
                                      (src/flashattention.cu:174-182, 236-290, 326-354)
  f64(q, k, v, scale, causal)        float64 softmax(scale*QK^T [+mask]) V (bench_flashattention.py:36-48 + scale)
  llmc_cpu(inp, B, T, C, NH)         the llm.c CPU loop (src/llm.c/attention_forward.cu:53-125)
  numpy_f64(q, k, v, scale, causal)  same maths as f64() written with numpy only (cross-check of the C code)
  backward_f64(q, k, v, d_o, ...)    float64 gradients (dQ, dK, dV) of that operator, written out analytically (the reference has no
                                      backward — README.md:33 — so this is pinned by central differences of numpy_f64, not by it)
  ref_llmc_cpu(inp, B, T, C, NH)     the REFERENCE's attention_forward_cpu itself, from oracle/_ref/libllmc_ref.so
                                      (only where /root/reference was compiled; returns None otherwise)
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "_build" / "libfa_oracle.so"
        if not so.exists() or so.stat().st_mtime < (_HERE / "fa_oracle.c").stat().st_mtime:
            subprocess.run(["make", "-C", str(_HERE), "oracle"], check=True, capture_output=True)
        lib = ctypes.CDLL(str(so))
        f32p, f64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
        i64, i32 = ctypes.c_int64, ctypes.c_int
        lib.fa_oracle_tiled.argtypes = [f32p, f32p, f32p, f32p, f32p, i64, i64, i64, i32, ctypes.c_float, i32]
        lib.fa_oracle_tiled.restype = None
        lib.fa_oracle_f64.argtypes = [f64p, f64p, f32p, f32p, f32p, i64, i64, i64, i32, ctypes.c_double, i32]
        lib.fa_oracle_f64.restype = None
        lib.fa_oracle_llmc_cpu.argtypes = [f32p, f32p, i32, i32, i32, i32]
        lib.fa_oracle_llmc_cpu.restype = None
        lib.fa_oracle_merge.argtypes = [f64p, f64p, f64p, f64p, i64, i32]
        lib.fa_oracle_merge.restype = None
        _LIB = lib
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _flatten(q, k, v):
    q, k, v = _f32(q), _f32(k), _f32(v)
    shape = q.shape
    d = shape[-1]
    q3, k3, v3 = q.reshape(-1, q.shape[-2], d), k.reshape(-1, k.shape[-2], d), v.reshape(-1, v.shape[-2], d)
    assert k3.shape == v3.shape and k3.shape[0] == q3.shape[0] and k3.shape[2] == d
    return q3, k3, v3, shape


def tiled(q, k, v, scale=1.0, causal=False):
    """-> (O fp32 same shape as q, LSE fp32 q.shape[:-1])"""
    q3, k3, v3, shape = _flatten(q, k, v)
    out = np.empty_like(q3)
    lse = np.empty(q3.shape[:2], dtype=np.float32)
    _lib().fa_oracle_tiled(_p(out, ctypes.c_float), _p(lse, ctypes.c_float), _p(q3, ctypes.c_float), _p(k3, ctypes.c_float),
                           _p(v3, ctypes.c_float), q3.shape[0], q3.shape[1], k3.shape[1], q3.shape[2], float(scale), int(bool(causal)))
    return out.reshape(shape), lse.reshape(shape[:-1])


def f64(q, k, v, scale=1.0, causal=False):
    """-> (O float64, LSE float64)"""
    q3, k3, v3, shape = _flatten(q, k, v)
    out = np.empty(q3.shape, dtype=np.float64)
    lse = np.empty(q3.shape[:2], dtype=np.float64)
    _lib().fa_oracle_f64(_p(out, ctypes.c_double), _p(lse, ctypes.c_double), _p(q3, ctypes.c_float), _p(k3, ctypes.c_float),
                         _p(v3, ctypes.c_float), q3.shape[0], q3.shape[1], k3.shape[1], q3.shape[2], float(scale), int(bool(causal)))
    return out.reshape(shape), lse.reshape(shape[:-1])


def llmc_cpu(inp, B, T, C, NH):
    inp = _f32(inp).reshape(B, T, 3 * C)
    out = np.empty((B, T, C), dtype=np.float32)
    _lib().fa_oracle_llmc_cpu(_p(out, ctypes.c_float), _p(inp, ctypes.c_float), B, T, C, NH)
    return out


def merge(o_acc, lse_acc, o_new, lse_new):
    """fp64 log-sum-exp merge of two partials; returns new (o, lse)."""
    o = np.array(o_acc, dtype=np.float64, copy=True, order="C")
    l = np.array(lse_acc, dtype=np.float64, copy=True, order="C")
    on = np.ascontiguousarray(o_new, dtype=np.float64)
    ln = np.ascontiguousarray(lse_new, dtype=np.float64)
    d = o.shape[-1]
    _lib().fa_oracle_merge(_p(o, ctypes.c_double), _p(l, ctypes.c_double), _p(on, ctypes.c_double), _p(ln, ctypes.c_double),
                           o.size // d, d)
    return o, l


def numpy_f64(q, k, v, scale=1.0, causal=False):
    """Independent numpy evaluation (small sizes): -> (O float64, LSE float64)."""
    q, k, v = (np.asarray(t, dtype=np.float64) for t in (q, k, v))   # (float64 inputs are used as they are)
    shape, d = q.shape, q.shape[-1]
    q3, k3, v3 = q.reshape(-1, q.shape[-2], d), k.reshape(-1, k.shape[-2], d), v.reshape(-1, v.shape[-2], d)
    assert k3.shape == v3.shape and k3.shape[0] == q3.shape[0]
    s = np.einsum("bid,bjd->bij", q3, k3) * scale
    if causal:
        n_q, n_k = q3.shape[1], k3.shape[1]
        i = np.arange(n_q)[:, None]
        j = np.arange(n_k)[None, :]
        s = np.where(j <= i + (n_k - n_q), s, -np.inf)
    mx = s.max(axis=-1, keepdims=True)
    mx = np.where(np.isfinite(mx), mx, 0.0)
    p = np.exp(s - mx)
    l = p.sum(axis=-1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        o = np.where(l > 0, np.einsum("bij,bjd->bid", p, v3) / l, 0.0)
        lse = np.where(l[..., 0] > 0, mx[..., 0] + np.log(l[..., 0]), -np.inf)
    return o.reshape(shape), lse.reshape(shape[:-1])


def backward_f64(q, k, v, d_o, scale=1.0, causal=False):
    """float64 (dQ, dK, dV) of O = softmax(scale Q K^T [+ causal mask]) V contracted with d_o.  q, d_o: [B, H, n_q, d]; k, v:
    [B, H_kv, n_k, d] with H % H_kv == 0 (query head h reads K/V head h // (H / H_kv); their gradients sum over the group).
        P = softmax(S),  D = rowsum(dO * O),  dV = P^T dO,  dS = P (dO V^T - D),  dQ = scale dS K,  dK = scale dS^T Q"""
    q, k, v, d_o = (np.asarray(t, dtype=np.float64) for t in (q, k, v, d_o))
    B, H, n_q, d = q.shape
    Hk, n_k = k.shape[1], k.shape[2]
    g = H // Hk
    kk = np.repeat(k, g, axis=1)
    vv = np.repeat(v, g, axis=1)
    s = np.einsum("bhid,bhjd->bhij", q, kk) * scale
    if causal:
        i = np.arange(n_q)[:, None]
        j = np.arange(n_k)[None, :]
        s = np.where(j <= i + (n_k - n_q), s, -np.inf)
    mx = s.max(axis=-1, keepdims=True)
    mx = np.where(np.isfinite(mx), mx, 0.0)
    p = np.exp(s - mx)
    l = p.sum(axis=-1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        p = np.where(l > 0, p / l, 0.0)
    o = np.einsum("bhij,bhjd->bhid", p, vv)
    delta = (d_o * o).sum(-1, keepdims=True)
    dp = np.einsum("bhid,bhjd->bhij", d_o, vv)
    ds = p * (dp - delta)
    dq = scale * np.einsum("bhij,bhjd->bhid", ds, kk)
    dk = scale * np.einsum("bhij,bhid->bhjd", ds, q).reshape(B, Hk, g, n_k, d).sum(2)
    dv = np.einsum("bhij,bhid->bhjd", p, d_o).reshape(B, Hk, g, n_k, d).sum(2)
    return dq, dk, dv


def packed_qkv_to_bhnd(inp, B, T, C, NH):
    """(B,T,3,NH,hs) -> q,k,v (B,NH,T,hs): what permute_kernel does (src/llm.c/attention_forward.cu:519-547)."""
    hs = C // NH
    x = _f32(inp).reshape(B, T, 3, NH, hs)
    return tuple(np.ascontiguousarray(x[:, :, i].transpose(0, 2, 1, 3)) for i in range(3))


# ---------------------------------------------------------------------------------------------
# the reference itself (only where it was compiled from /root/reference)
# ---------------------------------------------------------------------------------------------
def ref_llmc_cpu(inp, B, T, C, NH):
    """Calls the reference's own attention_forward_cpu (src/llm.c/attention_forward.cu:53) from
    oracle/_ref/libllmc_ref.so.  It materialises preatt/att (B*NH*T*T floats each) so keep T small."""
    so = _HERE / "_ref" / "libllmc_ref.so"
    if not so.exists():
        return None
    try:
        lib = ctypes.CDLL(str(so))
        fn = getattr(lib, "_Z21attention_forward_cpuPfS_S_PKfiiii")
    except (OSError, AttributeError):
        return None
    f32p = ctypes.POINTER(ctypes.c_float)
    fn.argtypes = [f32p, f32p, f32p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    fn.restype = None
    inp = _f32(inp).reshape(B, T, 3 * C)
    out = np.empty((B, T, C), dtype=np.float32)
    preatt = np.empty((B, NH, T, T), dtype=np.float32)
    att = np.empty((B, NH, T, T), dtype=np.float32)
    fn(_p(out, ctypes.c_float), _p(preatt, ctypes.c_float), _p(att, ctypes.c_float), _p(inp, ctypes.c_float), B, T, C, NH)
    return out


def ref_llmc_gpu_entry():
    """The REFERENCE's own attention_forward6(out, inp, B, T, C, NH, block_size) (src/llm.c/attention_forward.cu:1106-1179:
    permute -> flashattention kernel -> unpermute, device pointers) from oracle/_ref/libllmc_ref.so, as a ctypes function, or
    None where the reference was not compiled.  Needs a GPU; it prints a timing line per call like the reference does."""
    so = _HERE / "_ref" / "libllmc_ref.so"
    if not so.exists():
        return None
    try:
        fn = getattr(ctypes.CDLL(str(so)), "_Z18attention_forward6PfPKfiiiii")
    except (OSError, AttributeError):
        return None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5
    fn.restype = None
    return fn


def have_ref_torch_ext(d=64):
    return (_HERE / "_ref" / f"flash_ref_d{d}.so").exists()


def load_ref_torch_ext(d=64):
    """The reference torch extension (src/main.cpp + src/flashattention.cu built for sm_100a, head dim d)."""
    import importlib.util

    so = _HERE / "_ref" / f"flash_ref_d{d}.so"
    if not so.exists():
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)

    spec = importlib.util.spec_from_file_location(f"flash_ref_d{d}", str(so))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
