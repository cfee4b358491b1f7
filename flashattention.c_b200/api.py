"""Host-side operators: the reference's Python-visible surface, re-implemented over the C-ABI.

torch is used for device memory and streams only (plumbing); all arithmetic happens in libfa_b200.so.
"""
from __future__ import annotations

import ctypes
import importlib.util
import math

import torch

from . import _lib
from ._lib import FA_BF16, FA_F16, FA_F32, FaError, FaParams, check, lib


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return FA_F32
    if t.dtype == torch.bfloat16:
        return FA_BF16
    if t.dtype == torch.float16:
        return FA_F16
    raise FaError(f"unsupported dtype {t.dtype}: only float32 (tf32 tensor cores), bfloat16 and float16")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _shape4(Q, K, V):
    """-> (batch, heads, n_q, n_k, d, kv_heads).  K and V may have fewer heads than Q (grouped-query / multi-query attention,
    heads % kv_heads == 0) in the 4-D form [B, H_kv, N, d]; a 3-D [B*H, N, d] tensor carries no batch / head split, so K and V
    then need as many slices as Q."""
    if Q.dim() not in (3, 4) or K.dim() != Q.dim() or V.dim() != Q.dim():
        raise FaError("expected Q, K, V as [B*H, N, d] or [B, H, N, d]")
    if not (Q.is_cuda and K.is_cuda and V.is_cuda):
        raise FaError("Q, K, V must be CUDA tensors: this operator has no CPU path")
    if not (Q.dtype == K.dtype == V.dtype):
        raise FaError("Q, K, V dtype mismatch")
    if Q.dim() == 3:
        b, h, nq, d = 1, Q.shape[0], Q.shape[1], Q.shape[2]
        nk, hk = K.shape[1], K.shape[0]
        ok = hk == h and V.shape[0] == h and V.shape[1] == nk and K.shape[2] == d and V.shape[2] == d
    else:
        b, h, nq, d = Q.shape
        nk, hk = K.shape[2], K.shape[1]
        ok = (K.shape[0] == b and V.shape[0] == b and V.shape[1] == hk and hk > 0 and h % hk == 0 and V.shape[2] == nk
              and K.shape[3] == d and V.shape[3] == d)
    if not ok:
        raise FaError(f"shape mismatch: Q {tuple(Q.shape)} K {tuple(K.shape)} V {tuple(V.shape)}")
    return b, h, nq, nk, d, hk


def _tma_view(t: torch.Tensor) -> torch.Tensor:
    """A strided view is passed to the kernel as it is (the strides go into the TMA tensor map: sequence slices, packed or
    head-interleaved layouts need no copy) when its last axis is contiguous and base and strides are 16-byte aligned;
    anything else is made contiguous first."""
    es = t.element_size()
    ok = t.stride(-1) == 1 and t.data_ptr() % 16 == 0 and all((s * es) % 16 == 0 for s in t.stride()[:-1])
    # a broadcast axis (stride 0, size > 1: K/V expanded over query heads) has no tiled tensor map: materialise it
    ok = ok and all(s != 0 or n == 1 for s, n in zip(t.stride()[:-1], t.shape[:-1]))
    return t if ok else t.contiguous()


def _strides_bhn(t: torch.Tensor):
    """(batch, head, row) strides in elements of a [B*H, N, d] or [B, H, N, d] tensor."""
    if t.dim() == 3:
        return t.stride(0) * t.shape[0], t.stride(0), t.stride(1)
    return t.stride(0), t.stride(1), t.stride(2)


def attention(Q, K, V, causal=False, scale=None, return_lse=False, out_f32=False, impl=_lib.FA_IMPL_AUTO, out=None,
              batch_invariant=False, precise=False, acc=None):
    """O = softmax(scale * Q K^T [+ causal mask]) V.   scale defaults to 1/sqrt(d).

    Q, K, V: CUDA tensors [B*H, N, d] or [B, H, N, d], float32, bfloat16 or float16; strided views with a contiguous last axis
    (e.g. a slice of the sequence axis) are read in place.  4-D K and V may have fewer heads than Q (grouped-query / multi-query
    attention: [B, H_kv, N, d] with H % H_kv == 0; query head h reads K/V head h // (H / H_kv)).
    Returns O (same shape/dtype as Q; float32 if out_f32) and, if return_lse, LSE float32 [..., N].
    batch_invariant: FA_FLAG_BATCH_INVARIANT — a (batch, head) slice gives bit-identical results whatever else is in the
    launch (alone, in a larger batch, on another rank of a B x H sharded job); costs the split-KV tail optimisation.
    precise: FA_FLAG_PRECISE — float32 inputs only: fp32-grade contractions ("3xTF32": hi/lo operand split, three tcgen05 MMAs
    per contraction) instead of plain tf32; head dims <= 64 on the tensor cores, larger ones on the CUDA-core kernel.
    acc: (O_acc, LSE_acc) — accumulate mode: an earlier normalised partial of the same queries over other keys (float32,
    contiguous, shapes of O and LSE) is merged into this call's result by the log-sum-exp rule inside the kernel's epilogue.
    With a float32 result (float32 inputs, or out_f32) and no `out`, O_acc is updated in place and returned; LSE_acc is always
    updated in place (and returned if return_lse).  This is one step of the ring forward (ring.py).
    """
    b, h, nq, nk, d, hk = _shape4(Q, K, V)
    Q, K, V = _tma_view(Q), _tma_view(K), _tma_view(V)
    if scale is None:
        scale = 1.0 / math.sqrt(d)
    with torch.cuda.device(Q.device):
        o_dtype = torch.float32 if out_f32 else Q.dtype
        if acc is not None:
            o_acc, lse_acc = acc
            if (o_acc.dtype != torch.float32 or lse_acc.dtype != torch.float32 or o_acc.shape != Q.shape or lse_acc.shape != Q.shape[:-1]
                    or not o_acc.is_contiguous() or not lse_acc.is_contiguous() or o_acc.device != Q.device or lse_acc.device != Q.device):
                raise FaError("acc = (O_acc, LSE_acc) must be contiguous float32 CUDA tensors shaped like O and LSE")
            if out is None and o_dtype == torch.float32:
                out = o_acc
        O = out if out is not None else torch.empty(Q.shape, dtype=o_dtype, device=Q.device)
        if out is not None and (O.shape != Q.shape or O.dtype != o_dtype or not O.is_contiguous()):
            raise FaError("out tensor has the wrong shape/dtype or is not contiguous")
        if acc is not None:
            lse = acc[1]
        else:
            lse = torch.empty(Q.shape[:-1], dtype=torch.float32, device=Q.device) if return_lse else None
        p = FaParams()
        p.q, p.k, p.v, p.o = Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr()
        p.lse = lse.data_ptr() if lse is not None else None
        p.batch, p.heads, p.n_q, p.n_k, p.head_dim = b, h, nq, nk, d
        p.dtype = _dtype_code(Q)
        p.causal = 1 if causal else 0
        p.o_f32 = 1 if (out_f32 and Q.dtype != torch.float32) else 0
        p.scale = float(scale)
        (p.q_stride_b, p.q_stride_h, p.q_stride_n) = _strides_bhn(Q)
        (p.k_stride_b, p.k_stride_h, p.k_stride_n) = _strides_bhn(K)
        (p.v_stride_b, p.v_stride_h, p.v_stride_n) = _strides_bhn(V)
        p.o_stride_n, p.o_stride_h, p.o_stride_b = d, nq * d, h * nq * d
        p.impl = int(impl)
        p.flags = (_lib.FA_FLAG_BATCH_INVARIANT if batch_invariant else 0) | (_lib.FA_FLAG_PRECISE if precise else 0)
        if acc is not None:
            p.o_acc, p.lse_acc = acc[0].data_ptr(), acc[1].data_ptr()
        p.kv_heads = hk if hk != h else 0
        check(lib().fa_forward_ex(ctypes.byref(p), ctypes.c_void_p(_stream_ptr(Q.device))), "fa_forward_ex")
    return (O, lse) if return_lse else O


def backward_supported(Q) -> bool:
    """Whether fa_backward has a kernel instance for Q's dtype and head dim (bf16 / fp16, head_dim % 8 == 0, <= 128)."""
    return Q.dtype in (torch.bfloat16, torch.float16) and Q.shape[-1] % 8 == 0 and Q.shape[-1] <= 128


def attention_backward(Q, K, V, O, LSE, dO, causal=False, scale=None):
    """(dQ, dK, dV) of O = softmax(scale * Q K^T [+ causal mask]) V through the tcgen05 backward kernels (include/fa_b200.h:
    fa_backward; csrc/fa_bwd_sm100.cuh).  Q, K, V as for `attention` (K, V may have fewer heads in the 4-D form: their gradients are
    summed over the group); O and LSE are what `attention(..., return_lse=True)` returned, dO is shaped like O.  bf16 / fp16,
    head_dim <= 128; anything else raises FaError — there is no fallback inside this call."""
    b, h, nq, nk, d, hk = _shape4(Q, K, V)
    if not backward_supported(Q):
        raise FaError(f"attention_backward: no kernel instance for dtype {Q.dtype} head_dim {d} (bf16 / fp16, head_dim % 8 == 0, <= 128)")
    if O.shape != Q.shape or dO.shape != Q.shape or O.dtype != Q.dtype or dO.dtype != Q.dtype or not (O.is_cuda and dO.is_cuda):
        raise FaError("attention_backward: O and dO must match Q's shape, dtype and device")
    if LSE.dtype != torch.float32 or LSE.numel() != b * h * nq or not LSE.is_cuda:
        raise FaError("attention_backward: LSE must be the forward's fp32 [batch, heads, n_q]")
    if scale is None:
        scale = 1.0 / math.sqrt(d)
    q, k, v, o, do = (_tma_view(t) for t in (Q, K, V, O, dO))
    lse = LSE.contiguous()
    dq = torch.empty(Q.shape, dtype=Q.dtype, device=Q.device)
    dk = torch.empty(K.shape, dtype=K.dtype, device=K.device)
    dv = torch.empty(V.shape, dtype=V.dtype, device=V.device)
    p = _lib.FaBwdParams()
    p.q, p.k, p.v, p.o, p.d_o, p.lse = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr()
    p.dq, p.dk, p.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    p.batch, p.heads, p.kv_heads, p.n_q, p.n_k = b, h, hk, nq, nk
    p.head_dim, p.dtype, p.causal, p.scale = d, _dtype_code(Q), int(bool(causal)), float(scale)
    for name, t in (("q", q), ("k", k), ("v", v), ("o", o), ("do", do), ("dq", dq), ("dk", dk), ("dv", dv)):
        sb, sh, sn = _strides_bhn(t)
        setattr(p, f"{name}_stride_b", sb)
        setattr(p, f"{name}_stride_h", sh)
        setattr(p, f"{name}_stride_n", sn)
    with torch.cuda.device(Q.device):
        _lib.check(_lib.lib().fa_backward(ctypes.byref(p), _stream_ptr(Q.device)), "fa_backward")
    return dq, dk, dv


def forward(Q, K, V, causal=False):
    """Reference operator `forward(Q_d, K_d, V_d, causal) -> O` (src/main.cpp:3; src/flashattention.cu:603-617).

    Reference semantics: the scores are NOT scaled by 1/sqrt(d) (`scaling = 1.0`, src/flashattention.cu:593, 600).
    Unlike the reference, inputs are validated, 4-D [B, H, N, d] is accepted as well as 3-D [B*H, N, d], the
    launch is asynchronous on the current stream, and errors raise instead of being dropped.
    """
    return attention(Q, K, V, causal=bool(causal), scale=1.0)


def attention_host(q, k, v, causal=False, scale=None, out=None):
    """End-to-end call on HOST tensors (pinned or pageable): H2D copies, kernel, D2H copy, synchronise —
    all inside libfa_b200.so (fa_forward_host).  Mirrors what bench_flashattention.py:31-33,70 does around forward()."""
    if q.is_cuda or k.is_cuda or v.is_cuda:
        raise FaError("attention_host takes host tensors")
    if q.dim() not in (3, 4) or k.dim() != q.dim() or v.dim() != q.dim():
        raise FaError("expected Q, K, V as [B*H, N, d] or [B, H, N, d]")
    if not (q.dtype == k.dtype == v.dtype):
        raise FaError("Q, K, V dtype mismatch")
    if q.dim() == 3:
        b, h, nq, d = 1, q.shape[0], q.shape[1], q.shape[2]
        nk = k.shape[1]
    else:
        b, h, nq, d = q.shape
        nk = k.shape[2]
    # the library reads and writes raw host pointers: every size it derives from (b, h, nq, nk, d) must be backed by the tensors
    if tuple(k.shape) != tuple(q.shape[:-2]) + (nk, d) or tuple(v.shape) != tuple(k.shape):
        raise FaError(f"shape mismatch: Q {tuple(q.shape)} K {tuple(k.shape)} V {tuple(v.shape)}")
    if scale is None:
        scale = 1.0 / math.sqrt(d)
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    o = out if out is not None else torch.empty_like(q)
    if o.is_cuda or o.shape != q.shape or o.dtype != q.dtype or not o.is_contiguous():
        raise FaError("out must be a contiguous host tensor of Q's shape and dtype")
    check(lib().fa_forward_host(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), b, h, nq, nk, d, float(scale),
                                1 if causal else 0, _dtype_code(q)), "fa_forward_host")
    return o


def attention_forward6(out, inp, B, T, C, NH, block_size=256, return_lse=False, precise=True):
    """llm.c-style entry (src/llm.c/attention_forward.cu:1106-1179): inp (B,T,3C) packed QKV fp32 on the device,
    out (B,T,C); causal, scale 1/sqrt(C/NH).  The packed layout is consumed in place through strided TMA maps —
    no permute/unpermute kernels, no temporaries.  `block_size` is accepted and ignored (it only sized the
    reference's layout kernels).  precise=True (the default, like the C symbol): fp32-grade contractions, so the harness's
    own validate_result(out, 1e-4f) gate (src/llm.c/attention_forward.cu:1262) holds; precise=False: plain tf32."""
    if C % NH:
        raise FaError("C must be divisible by NH")
    hs = C // NH
    if inp.dtype != torch.float32 or out.dtype != torch.float32 or not inp.is_cuda or not out.is_cuda:
        raise FaError("attention_forward expects float32 CUDA tensors")
    if inp.numel() != B * T * 3 * C or out.numel() != B * T * C or not inp.is_contiguous() or not out.is_contiguous():
        raise FaError("attention_forward: bad tensor sizes")
    with torch.cuda.device(inp.device):
        lse = torch.empty((B, NH, T), dtype=torch.float32, device=inp.device) if return_lse else None
        check(lib().fa_forward_packed_qkv_ex(inp.data_ptr(), out.data_ptr(), lse.data_ptr() if lse is not None else None, B, T, NH, hs,
                                             1.0 / math.sqrt(hs), 1, _lib.FA_FLAG_PRECISE if precise else 0,
                                             ctypes.c_void_p(_stream_ptr(inp.device))), "fa_forward_packed_qkv_ex")
    return (out, lse) if return_lse else out


def attention_forward(kernel_num, out, inp, B, T, C, NH, block_size=256, precise=True):
    """Kernel-number dispatch of the llm.c harness (src/llm.c/attention_forward.cu:1183-1211).  Only kernel 6 — the
    reference author's flash kernel — is on the hot path; 1-5 are upstream llm.c comparison kernels (out of scope)."""
    if kernel_num != 6:
        raise FaError("Invalid kernel number (only kernel 6, the flash attention kernel, is provided)")
    return attention_forward6(out, inp, B, T, C, NH, block_size, precise=precise)


def merge_partials(o_acc, lse_acc, o_new, lse_new):
    """In-place log-sum-exp merge of two attention partials (fp32 O [..., d], fp32 LSE [...])."""
    for t in (o_acc, lse_acc, o_new, lse_new):
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise FaError("merge_partials expects contiguous float32 CUDA tensors")
    d = o_acc.shape[-1]
    rows = o_acc.numel() // d
    with torch.cuda.device(o_acc.device):
        check(lib().fa_merge_partials(o_acc.data_ptr(), lse_acc.data_ptr(), o_new.data_ptr(), lse_new.data_ptr(), rows, d,
                                      ctypes.c_void_p(_stream_ptr(o_acc.device))), "fa_merge_partials")
    return o_acc, lse_acc


def cast_to_16(src, dtype, dst=None):
    """fp32 -> bfloat16 / float16 on the current stream (the final cast of the ring accumulator)."""
    if dst is None:
        dst = torch.empty(src.shape, dtype=dtype, device=src.device)
    with torch.cuda.device(src.device):
        check(lib().fa_cast_f32(src.data_ptr(), dst.data_ptr(), src.numel(), _dtype_code(dst), ctypes.c_void_p(_stream_ptr(src.device))),
              "fa_cast_f32")
    return dst


def cast_to_bf16(src, dst=None):
    return cast_to_16(src, torch.bfloat16, dst)


def last_impl() -> int:
    return lib().fa_last_impl()


def launch_count() -> int:
    return lib().fa_launch_count()


_ext = None


def load_extension():
    """The prebuilt pybind module with the reference's `forward(Q, K, V, causal)` (csrc/torch_binding.cpp) — what
    `load(name='flash', sources=['src/main.cpp', 'src/flashattention.cu'])` returns in bench_flashattention.py:10."""
    global _ext
    if _ext is None:
        so = _lib.PKG_DIR / "flash_b200.so"
        if not so.exists():
            raise FaError(f"{so} is missing: run `python flashattention.c_b200/build.py`")
        lib()  # load libfa_b200.so first so the extension resolves it
        spec = importlib.util.spec_from_file_location("flash_b200", str(so))
        _ext = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_ext)
    return _ext
