"""Multi-GPU forms of the forward (one process per GPU, torch.distributed for the plumbing).

Neither exists in the reference (single GPU only) — their oracle is the single-GPU result.

  B x H sharding (configs 3, 4): every (batch, head) is independent (`batch_offset = batch_stride * blockIdx.x`,
      src/flashattention.cu:144), so ranks take contiguous slices of the flattened B*H axis and run the local
      forward.  NO collective is on the data path; `gather_bh` exists only for parity checks.

  Ring attention (config 5): Q, K, V are partitioned along the sequence.  In step s rank r attends its Q shard to
      the K/V shard that started on rank (r - s) mod P while the next shard moves r -> r+1 with NCCL send/recv
      on a side stream (double-buffered, overlapped with the tile loop of the local kernel); partial (O, LSE)
      pairs are merged with the log-sum-exp rule by fa_merge_partials.

  Causal ring, balanced (zig-zag): with contiguous sequence shards a causal ring is lopsided — rank 0 has one shard
      of visible keys, rank P-1 has P.  `zigzag=True` cuts the sequence into 2P chunks and gives rank r chunks r and
      2P-1-r, so every rank has the same N^2/(2P) visible (query, key) pairs and every ring step costs the same on
      every rank (`zigzag_shard` / `zigzag_unshard` convert between the layouts).
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


def sharded_attention(Q, K, V, causal=False, scale=None, group=None, already_sharded=False, batch_invariant=False):
    """B x H-sharded forward: returns this rank's slice of O ([bh_local, N, d]).  No communication.
    batch_invariant=True makes the gathered result bit-identical to the unsharded forward (FA_FLAG_BATCH_INVARIANT)."""
    from .api import attention

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if not already_sharded:
        Q, K, V = (shard_bh(t, rank, world) for t in (Q, K, V))
    if Q.shape[0] == 0:
        return Q.new_empty(Q.shape)
    return attention(Q.contiguous(), K.contiguous(), V.contiguous(), causal=causal, scale=scale, batch_invariant=batch_invariant)


def gather_bh(o_local: torch.Tensor, total_bh: int, group=None) -> torch.Tensor:
    """Parity-check helper: all-gather the per-rank O slices back into [B*H, N, d]."""
    world = dist.get_world_size(group)
    per = (total_bh + world - 1) // world
    pad = o_local.new_zeros((per,) + tuple(o_local.shape[1:]))
    pad[: o_local.shape[0]] = o_local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat(outs, dim=0)[:total_bh]


def zigzag_shard(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Full-sequence [..., N, d] -> rank's zig-zag shard [..., N/P, d]: chunk `rank` followed by chunk `2P-1-rank` of the
    2P equal chunks of the sequence axis."""
    n = x.shape[-2]
    if n % (2 * world):
        raise ValueError(f"sequence length {n} is not a multiple of 2 * world = {2 * world}")
    c = n // (2 * world)
    lo, hi = rank, 2 * world - 1 - rank
    return torch.cat([x[..., lo * c:(lo + 1) * c, :], x[..., hi * c:(hi + 1) * c, :]], dim=-2)


def zigzag_unshard(shards, world: int) -> torch.Tensor:
    """Inverse of zigzag_shard over the list of all ranks' shards ([..., N/P, d] each, rank order)."""
    c = shards[0].shape[-2] // 2
    chunks = [None] * (2 * world)
    for r, sh in enumerate(shards):
        chunks[r] = sh[..., :c, :]
        chunks[2 * world - 1 - r] = sh[..., c:, :]
    return torch.cat(chunks, dim=-2)


def zigzag_step_plan(rank: int, src: int):
    """What rank `rank` computes while it holds the K/V shard of rank `src` (both in the zig-zag layout [lo | hi]).
    A list of (query half, key range, causal) with query half 0 = lo / 1 = hi, key range 'lo' (first half) or 'all';
    causal is bottom-right aligned.  Every entry list covers exactly 2 c^2 (query, key) pairs (c = chunk length), or
    c^2/2 + 3 c^2/2 on the diagonal step."""
    if src == rank:   # [Q_lo | Q_hi] against its own keys: lo sees lo causally; hi sees all of lo and hi causally
        return [(0, "lo", True), (1, "all", True)]
    if src < rank:    # K_lo(src) is in the past of both query chunks, K_hi(src) in the future of both
        return [(0, "lo", False), (1, "lo", False)]
    return [(1, "all", False)]   # src > rank: both key chunks are in the past of Q_hi and in the future of Q_lo


def ring_schedule(rank: int, world: int):
    """[(step, source_rank_of_the_kv_shard_processed_in_that_step)] for `rank`."""
    return [(s, (rank - s) % world) for s in range(world)]


def ring_attention(q, k, v, causal=False, scale=None, group=None, zigzag=False, _attn=None, _merge=None, _finalize=None):
    """Sequence-partitioned forward.  q, k, v: this rank's shards [B, H, N/P, d] (or [B*H, N/P, d]), rank r holding
    sequence positions [r*N/P, (r+1)*N/P) — or, with zigzag=True (causal only), chunks r and 2P-1-r of 2P
    (zigzag_shard).  Returns this rank's shard of O in q's dtype and the fp32 LSE, in the same layout as q.

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
    if zigzag and not causal:
        raise ValueError("zigzag sharding only makes sense for the causal ring")
    if zigzag and q.shape[-2] % 2:
        raise ValueError("a zig-zag shard holds two chunks of equal length")
    zz = zigzag and world > 1
    c = q.shape[-2] // 2
    # zig-zag: separate accumulators for the two query chunks (each step touches one or both)
    acc = [[None, None], [None, None]]

    def zz_accumulate(half, o_s, lse_s):
        if acc[half][0] is None:
            acc[half] = [o_s, lse_s]
        else:
            acc[half] = list(merge(acc[half][0], acc[half][1], o_s, lse_s))

    if world == 1:
        o, lse = attn(q, k, v, causal)
        return (_finalize(o) if _finalize else (api.cast_to_16(o, q.dtype) if (on_gpu and q.dtype != torch.float32) else o)), lse

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
        if zz:
            for half, keys, cz in zigzag_step_plan(rank, src):
                q_h = q[..., half * c:(half + 1) * c, :]
                k_s, v_s = (k_cur, v_cur) if keys == "all" else (k_cur[..., :c, :], v_cur[..., :c, :])
                zz_accumulate(half, *attn(q_h, k_s, v_s, cz))
        elif causal and src > rank:
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
    if zz:
        o_acc = torch.cat([acc[0][0], acc[1][0]], dim=-2)
        lse_acc = torch.cat([acc[0][1], acc[1][1]], dim=-1)
    if _finalize is not None:
        return _finalize(o_acc), lse_acc
    if on_gpu and q.dtype != torch.float32:
        return api.cast_to_16(o_acc, q.dtype), lse_acc
    return o_acc, lse_acc
