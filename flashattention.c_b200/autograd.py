"""Autograd support for the forward operator (NOT part of the reference, which is forward only, nor of the hot path this
repository rebuilds: SURVEY.md 8(f) lists a backward pass under "generality the reference lacks").

`attention_autograd(Q, K, V, causal, scale)` runs the tcgen05 forward kernel and keeps (Q, K, V, O, LSE).  For bf16 / fp16 tensors
with head_dim <= 128 the backward is the tcgen05 backward (api.attention_backward -> fa_backward, csrc/fa_bwd_sm100.cuh: a
statistics pass, a dK/dV launch and a dQ launch, atomics-free).  Other cases (fp32 tensors, head_dim 256) recompute the probabilities
from the saved LSE — no N x N tensor is ever kept — in blocks of query rows with plain torch matmuls (library GEMMs: host-side
plumbing around the forward kernel, not a kernel of this repository).  Both compute

    P  = exp(scale * Q K^T - LSE)            D  = rowsum(dO * O)
    dV = P^T dO                              dS = P * (dO V^T - D)
    dQ = scale * dS K                        dK = scale * dS^T Q
"""
from __future__ import annotations

import math

import torch

from . import api


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal, scale, precise):
        o, lse = api.attention(q, k, v, causal=causal, scale=scale, return_lse=True, precise=precise)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.causal, ctx.scale = causal, scale
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, o, lse = ctx.saved_tensors
        causal, scale = ctx.causal, ctx.scale
        if api.backward_supported(q):
            dq, dk, dv = api.attention_backward(q, k, v, o, lse, d_o, causal=causal, scale=scale)
            return dq, dk, dv, None, None, None
        dq, dk, dv = recompute_backward(q, k, v, o, lse, d_o, causal, scale)
        return dq, dk, dv, None, None, None


def recompute_backward(q, k, v, o, lse, d_o, causal, scale):
    """(dQ, dK, dV) by blockwise recomputation from the saved LSE with torch matmuls (module docstring): the path of fp32 tensors and
    head_dim 256, and an independent cross-check of the backward kernels at lengths no N x N reference fits."""
    if k.shape[:-2] != q.shape[:-2]:
        raise api.FaError("the recomputation backward (fp32 tensors, head_dim > 128) needs K and V with Q's head count")
    shape = q.shape
    d = shape[-1]
    q3, k3, v3, o3, do3 = (t.reshape(-1, t.shape[-2], d) for t in (q, k, v, o, d_o.contiguous()))
    lse3 = lse.reshape(-1, lse.shape[-1])
    bh, n_q, n_k = q3.shape[0], q3.shape[1], k3.shape[1]
    dq = torch.empty_like(q3)
    dk = torch.zeros(k3.shape, dtype=torch.float32, device=k.device)
    dv = torch.zeros(v3.shape, dtype=torch.float32, device=v.device)
    delta = (do3.float() * o3.float()).sum(-1)                              # D
    # query rows per block: keep the fp32 [bh, rows, n_k] work tensors around 256 MiB
    rows = max(16, min(n_q, (256 << 20) // max(1, 4 * bh * n_k)))
    kt = k3.transpose(1, 2)
    vt = v3.transpose(1, 2)
    key_idx = torch.arange(n_k, device=q.device)
    for r0 in range(0, n_q, rows):
        r1 = min(n_q, r0 + rows)
        s = torch.matmul(q3[:, r0:r1], kt).float() * scale
        p = torch.exp(s - lse3[:, r0:r1, None])
        if causal:                                                           # bottom-right aligned, like the kernel
            visible = key_idx[None, :] <= (torch.arange(r0, r1, device=q.device)[:, None] + (n_k - n_q))
            p = p * visible
        p = torch.nan_to_num(p, nan=0.0)                                     # rows without a visible key: LSE = -inf
        dp = torch.matmul(do3[:, r0:r1], vt).float()
        ds = p * (dp - delta[:, r0:r1, None])
        p_lo, ds_lo = p.to(q.dtype), ds.to(q.dtype)
        dv += torch.matmul(p_lo.transpose(1, 2), do3[:, r0:r1]).float()
        dk += torch.matmul(ds_lo.transpose(1, 2), q3[:, r0:r1]).float() * scale
        dq[:, r0:r1] = (torch.matmul(ds_lo, k3).float() * scale).to(q.dtype)
    return dq.reshape(shape), dk.to(k.dtype).reshape(k.shape), dv.to(v.dtype).reshape(v.shape)


def attention_autograd(Q, K, V, causal=False, scale=None, precise=False):
    """`attention` with gradients: O = softmax(scale * Q K^T [+ causal mask]) V through the tcgen05 forward kernel, backward by
    the tcgen05 backward kernels (bf16 / fp16, head_dim <= 128) or blockwise recomputation from the saved LSE (module docstring).
    Q, K, V: CUDA tensors [B*H, N, d] or [B, H, N, d]; K, V may have fewer heads in the 4-D form when the backward kernel applies."""
    if V.shape != K.shape:
        raise api.FaError("attention_autograd: K and V shapes differ")
    if K.shape[:-2] != Q.shape[:-2] and not api.backward_supported(Q):
        raise api.FaError("attention_autograd: grouped K/V heads need the backward kernel (bf16 / fp16, head_dim <= 128); expand K/V first")
    if scale is None:
        scale = 1.0 / math.sqrt(Q.shape[-1])
    return _Attention.apply(Q, K, V, bool(causal), float(scale), bool(precise))
