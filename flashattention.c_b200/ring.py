"""Multi-GPU forms of the forward (one process per GPU, torch.distributed for the plumbing).

Neither exists in the reference (single GPU only) — their oracle is the single-GPU result.

  B x H sharding (configs 3, 4): every (batch, head) is independent (`batch_offset = batch_stride * blockIdx.x`,
      src/flashattention.cu:144), so ranks take contiguous slices of the flattened B*H axis and run the local
      forward.  NO collective is on the data path; `gather_bh` exists only for parity checks.

  Ring attention (config 5): Q, K, V are partitioned along the sequence.  In step s rank r attends its Q shard to
      the K/V shard that started on rank (r - s) mod P while the next shard moves r -> r+1 with NCCL send/recv
      on a side stream (double-buffered, overlapped with the tile loop of the local kernel); partial (O, LSE)
      pairs are merged with the log-sum-exp rule by fa_merge_partials.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def bh_shard_range(total_bh: int, rank: int, world: int):
    """Contiguous slice [start, stop) of the flattened B*H axis owned by `rank` (ceil split; trailing ranks may be empty)."""
    per = (total_bh + world - 1) // world
    start = min(total_bh, rank * per)
    return start, min(total_bh, start + per)


def shard_bh(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """[B*H, N, d] or [B, H, N, d] -> this rank's [bh_local, N, d] slice (a view when possible)."""
    x3 = x.reshape(-1, x.shape[-2], x.shape[-1])
    s, e = bh_shard_range(x3.shape[0], rank, world)
    return x3[s:e]


def sharded_attention(Q, K, V, causal=False, scale=None, group=None, already_sharded=False):
    """B x H-sharded forward: returns this rank's slice of O ([bh_local, N, d]).  No communication."""
    from .api import attention

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if not already_sharded:
        Q, K, V = (shard_bh(t, rank, world) for t in (Q, K, V))
    if Q.shape[0] == 0:
        return Q.new_empty(Q.shape)
    return attention(Q.contiguous(), K.contiguous(), V.contiguous(), causal=causal, scale=scale)


def gather_bh(o_local: torch.Tensor, total_bh: int, group=None) -> torch.Tensor:
    """Parity-check helper: all-gather the per-rank O slices back into [B*H, N, d]."""
    world = dist.get_world_size(group)
    per = (total_bh + world - 1) // world
    pad = o_local.new_zeros((per,) + tuple(o_local.shape[1:]))
    pad[: o_local.shape[0]] = o_local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat(outs, dim=0)[:total_bh]


def ring_schedule(rank: int, world: int):
    """[(step, source_rank_of_the_kv_shard_processed_in_that_step)] for `rank`."""
    return [(s, (rank - s) % world) for s in range(world)]


def ring_attention(q, k, v, causal=False, scale=None, group=None, _attn=None, _merge=None, _finalize=None):
    """Sequence-partitioned forward.  q, k, v: this rank's shards [B, H, N/P, d] (or [B*H, N/P, d]), rank r holding
    sequence positions [r*N/P, (r+1)*N/P).  Returns this rank's shard of O in q's dtype and the fp32 LSE.

    `_attn`, `_merge`, `_finalize` are test seams (the gloo/CPU tests inject the oracle to exercise the rotation and the
    merge without a GPU); the product path leaves them None and runs the CUDA kernels.
    """
    from . import api

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    d = q.shape[-1]
    if scale is None:
        scale = 1.0 / math.sqrt(d)
    attn = _attn or (lambda q_, k_, v_, c_: api.attention(q_, k_, v_, causal=c_, scale=scale, return_lse=True, out_f32=True))
    merge = _merge or api.merge_partials
    on_gpu = q.is_cuda

    if world == 1:
        o, lse = attn(q, k, v, causal)
        return (_finalize(o) if _finalize else (api.cast_to_bf16(o) if (on_gpu and q.dtype == torch.bfloat16) else o)), lse

    nxt = dist.get_global_rank(group, (rank + 1) % world) if group is not None else (rank + 1) % world
    prv = dist.get_global_rank(group, (rank - 1) % world) if group is not None else (rank - 1) % world
    k_cur, v_cur = k.contiguous(), v.contiguous()
    spare = [(torch.empty_like(k_cur), torch.empty_like(v_cur)) for _ in range(2)]  # receive buffers, ping-pong
    main = torch.cuda.current_stream(q.device) if on_gpu else None
    comm = torch.cuda.Stream(device=q.device) if on_gpu else None
    o_acc = lse_acc = None

    for step, src in ring_schedule(rank, world):
        reqs = []
        if step < world - 1:
            k_nxt, v_nxt = spare[step % 2]
            ops = [dist.P2POp(dist.isend, k_cur, nxt, group), dist.P2POp(dist.isend, v_cur, nxt, group),
                   dist.P2POp(dist.irecv, k_nxt, prv, group), dist.P2POp(dist.irecv, v_nxt, prv, group)]
            if on_gpu:
                comm.wait_stream(main)  # the buffer being overwritten was last read by the previous step's kernel
                with torch.cuda.stream(comm):
                    reqs = dist.batch_isend_irecv(ops)
            else:
                reqs = dist.batch_isend_irecv(ops)
        # local tile loop on the shard that is resident now (overlaps the transfer above)
        if causal and src > rank:
            pass  # every key of this shard is in the future of every local query
        else:
            o_s, lse_s = attn(q, k_cur, v_cur, bool(causal and src == rank))
            if o_acc is None:
                o_acc, lse_acc = o_s, lse_s
            else:
                o_acc, lse_acc = merge(o_acc, lse_acc, o_s, lse_s)
        if step < world - 1:
            for r in reqs:
                r.wait()
            if on_gpu:
                main.wait_stream(comm)
            k_cur, v_cur = k_nxt, v_nxt
    if _finalize is not None:
        return _finalize(o_acc), lse_acc
    if on_gpu and q.dtype == torch.bfloat16:
        return api.cast_to_bf16(o_acc), lse_acc
    return o_acc, lse_acc
