"""Multi-GPU forms of the forward (one process per GPU, torch.distributed for the plumbing).

Neither exists in the reference (single GPU only) — their oracle is the single-GPU result.

  B x H sharding (configs 3, 4): every (batch, head) is independent (`batch_offset = batch_stride * blockIdx.x`,
      src/flashattention.cu:144), so ranks take contiguous slices of the flattened B*H axis and run the local
      forward.  NO collective is on the data path; `gather_bh` exists only for parity checks.

  Ring attention (config 5): Q, K, V are partitioned along the sequence.  In step s rank r attends its Q shard to
      the K/V shard that started on rank (r - s) mod P while the next shard is on its way into a ping-pong staging
      buffer; partial (O, LSE) pairs are merged with the log-sum-exp rule by fa_merge_partials.  Two transports:
        "p2p"  (default on GPUs) every rank publishes its K/V shard in a CUDA-IPC-exported buffer and PULLS the shard
               it needs next straight from its owner with a device-to-device copy (fa_copy_async): a copy-engine
               transfer over NVLink / NVSwitch that needs no SM, so it really runs under the persistent attention
               kernel, which owns all 148 SMs for the whole step.  The shards never change, so there is no
               rotation chain: one barrier after publishing, one before the buffers may be reused.
        "nccl" K/V rotate r -> r+1 with NCCL send/recv on a side stream.  An NCCL kernel needs SMs; launched next
               to a persistent kernel it runs before or after it, not under it: measured, ring time = compute +
               transfer (profiles/r01_ring_nccl_settings_2gpu.log).  Kept for comparison and for the CPU (gloo) tests.

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


class _P2PState:
    """This rank's exported K/V buffer, the mappings of every peer's, the staging ping-pong and the side stream — cached per
    (process group, shard shape, device) and reused by every call.  The five methods publish / prefetch / acquire / release /
    finish are the transport protocol `_ring_pull` drives (the gloo tests drive the same schedule through a CPU stand-in)."""

    def __init__(self, like, group):
        import ctypes

        from ._lib import FaError, check, lib

        self.L, self.check = lib(), check
        self.group, self.dev = group, like.device
        self.half = like.numel() * like.element_size()     # bytes of K (= bytes of V)
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.local, self.peer = None, []

        def agree(ok, what):
            """Every rank learns whether the step worked everywhere, so that a failure on one rank (no CUDA IPC in this
            container, out of memory, ...) is an exception on all of them and not a hang in the next collective."""
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                self.unmap_peers()
                self.free_local()
                raise FaError(f"ring p2p transport unavailable: {what} failed on at least one rank (use transport='nccl')")

        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        rc = self.L.fa_p2p_alloc(2 * self.half, ctypes.byref(ptr), handle)
        if rc == 0:
            self.local = ptr.value
        agree(rc == 0, "fa_p2p_alloc")
        mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=self.dev)
        handles = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(handles, mine, group=group)
        ok = True
        for r in range(world):
            if r == rank:
                self.peer.append(self.local)
            else:
                pp = ctypes.c_void_p()
                ok = ok and self.L.fa_p2p_open(bytes(handles[r].cpu().tolist()), ctypes.byref(pp)) == 0
                self.peer.append(pp.value)
        agree(ok, "fa_p2p_open")
        self.side = torch.cuda.Stream(device=self.dev)
        self.flag = torch.zeros(1, device=self.dev)
        # staging ping-pong [2][K | V], shaped like the caller's shards
        self.stage = [torch.empty((2,) + tuple(like.shape), dtype=like.dtype, device=self.dev) for _ in range(2)]
        self.ev_copy, self.ev_free, self.ev_start = {}, [None, None], None

    def _copy(self, dst_ptr, src_ptr, nbytes, stream):
        import ctypes

        self.check(self.L.fa_copy_async(ctypes.c_void_p(dst_ptr), ctypes.c_void_p(src_ptr), nbytes, ctypes.c_void_p(stream.cuda_stream)),
                   "fa_copy_async")

    def _barrier(self):
        """Stream-ordered barrier on the current stream: completes on a rank only when every rank's stream has reached it."""
        dist.all_reduce(self.flag, group=self.group)

    def publish(self, kc, vc):
        """Copy this rank's shard into its exported buffer and wait (on the stream) until every rank has done so."""
        main = torch.cuda.current_stream(self.dev)
        self._copy(self.local, kc.data_ptr(), self.half, main)
        self._copy(self.local + self.half, vc.data_ptr(), self.half, main)
        self._barrier()
        self.ev_start = torch.cuda.Event()
        self.ev_start.record(main)
        self.ev_copy, self.ev_free = {}, [None, None]

    def prefetch(self, i, src):
        """Start pulling rank `src`'s shard into staging buffer i % 2 on the side stream (copy engine, no SM)."""
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_start)
            if self.ev_free[i % 2] is not None:
                self.side.wait_event(self.ev_free[i % 2])    # the kernels that read this buffer two steps ago are done
            self._copy(self.stage[i % 2].data_ptr(), self.peer[src], 2 * self.half, self.side)
            self.ev_copy[i] = torch.cuda.Event()
            self.ev_copy[i].record(self.side)

    def acquire(self, i, src):
        """(K, V) of pull i, valid for work enqueued on the current stream after this call."""
        torch.cuda.current_stream(self.dev).wait_event(self.ev_copy.pop(i))
        return self.stage[i % 2][0], self.stage[i % 2][1]

    def release(self, i):
        self.ev_free[i % 2] = torch.cuda.Event()
        self.ev_free[i % 2].record(torch.cuda.current_stream(self.dev))

    def finish(self):
        self._barrier()       # nobody is still pulling from a buffer the next call will overwrite

    def unmap_peers(self):
        import ctypes

        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        for r, pp in enumerate(self.peer):
            if r != rank and pp:
                self.L.fa_p2p_close(ctypes.c_void_p(pp))
        self.peer = []

    def free_local(self):
        import ctypes

        if self.local:
            self.L.fa_p2p_free(ctypes.c_void_p(self.local))
        self.local = None


_p2p_states = {}


def ring_p2p_release(group=None):
    """Drop the cached p2p transport state (exported buffers, peer mappings, staging buffers) of `group` — or of every group
    with group=None.  Collective: call it on every rank, before destroying the process group; the next ring_attention call
    rebuilds the state.  Order: nobody pulls any more -> every rank unmaps its peers' buffers -> every rank frees its own (an
    exported buffer must not be freed while another process still has it mapped)."""
    keys = [k for k in _p2p_states if group is None or k[0] == (id(group) if group is not None else 0)]
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier(group)
    for key in keys:
        _p2p_states[key].unmap_peers()
    if dist.is_initialized():
        dist.barrier(group)
    for key in keys:
        _p2p_states.pop(key).free_local()


def _p2p_state(k, group):
    key = (id(group) if group is not None else 0, k.device.index, tuple(k.shape), k.dtype)
    st = _p2p_states.get(key)
    if st is None:
        st = _p2p_states[key] = _P2PState(k, group)
    return st


def ring_calls(rank, world, causal, zz):
    """Every attention call of rank `rank`'s ring forward, in order: (source rank of the K/V shard, accumulator 0/1, query half
    or None for the whole shard, key range 'lo'/'all', causal flag, last call into that accumulator).  A causal ring with
    contiguous shards never looks at later ranks' shards; the zig-zag ring keeps one accumulator per query chunk."""
    calls = []
    for _, src in ring_schedule(rank, world):
        if zz:
            calls += [(src, hq, hq, keys, cz) for hq, keys, cz in zigzag_step_plan(rank, src)]
        elif not causal or src <= rank:
            calls.append((src, 0, None, "all", bool(causal and src == rank)))
    last = {}
    for i, cl in enumerate(calls):
        last[cl[1]] = i
    return [cl + (last[cl[1]] == i,) for i, cl in enumerate(calls)]


class _Partials:
    """The running (O, LSE) of a ring forward, one per accumulator.  Two ways to fold a step in:
    fused (the product path on GPUs) — the step's kernel takes the running partial as its accumulate input and merges in
        its epilogue; steps before the last keep O in fp32 in place, the last one writes O in the caller's dtype: no merge
        launches, no extra pass over O, no final cast;
    unfused — `attn` produces an fp32 partial and `merge` folds it in (the CPU test seams; fa_merge_partials on a GPU)."""

    def __init__(self, attn, merge, fused_scale=None):
        self.attn, self.merge, self.fused_scale = attn, merge, fused_scale
        self.acc = [None, None]

    def step(self, slot, q_, k_, v_, causal, is_last):
        from . import api

        if self.fused_scale is None:
            o_s, lse_s = self.attn(q_, k_, v_, causal)
            self.acc[slot] = [o_s, lse_s] if self.acc[slot] is None else list(self.merge(self.acc[slot][0], self.acc[slot][1], o_s, lse_s))
            return
        f32_out = not is_last and q_.dtype != torch.float32
        if self.acc[slot] is None:
            self.acc[slot] = list(api.attention(q_, k_, v_, causal=causal, scale=self.fused_scale, return_lse=True, out_f32=f32_out))
        else:
            self.acc[slot][0] = api.attention(q_, k_, v_, causal=causal, scale=self.fused_scale, out_f32=f32_out, acc=tuple(self.acc[slot]))

    def result(self, zz):
        if zz:
            return torch.cat([self.acc[0][0], self.acc[1][0]], dim=-2), torch.cat([self.acc[0][1], self.acc[1][1]], dim=-1)
        return self.acc[0][0], self.acc[0][1]


def _ring_pull(st, q, k, v, causal, zz, c, parts, rank, world):
    """The ring forward over a pull transport `st` (see _P2PState).  Returns O (fp32 accumulator, or the final dtype when the
    merges are fused into the kernels) and the LSE."""
    kc, vc = k.contiguous(), v.contiguous()
    st.publish(kc, vc)
    calls = ring_calls(rank, world, causal, zz)
    # the remote shards this rank needs, in ring order
    remote = []
    for cl in calls:
        if cl[0] != rank and cl[0] not in remote:
            remote.append(cl[0])

    def compute(src, k_s, v_s):
        for s_, slot, hq, keys, cz, is_last in calls:
            if s_ != src:
                continue
            q_ = q if hq is None else q[..., hq * c:(hq + 1) * c, :]
            kk, vv = (k_s, v_s) if keys == "all" else (k_s[..., :c, :], v_s[..., :c, :])
            parts.step(slot, q_, kk, vv, cz, is_last)

    if remote:
        st.prefetch(0, remote[0])
    compute(rank, kc, vc)                      # the local shard, while the first remote shard is on its way
    for i, src in enumerate(remote):
        if i + 1 < len(remote):
            st.prefetch(i + 1, remote[i + 1])
        compute(src, *st.acquire(i, src))
        st.release(i)
    st.finish()
    return parts.result(zz)


def ring_attention(q, k, v, causal=False, scale=None, group=None, zigzag=False, transport=None, fuse_merge=True, _attn=None,
                   _merge=None, _finalize=None):
    """Sequence-partitioned forward.  q, k, v: this rank's shards [B, H, N/P, d] (or [B*H, N/P, d]), rank r holding
    sequence positions [r*N/P, (r+1)*N/P) — or, with zigzag=True (causal only), chunks r and 2P-1-r of 2P
    (zigzag_shard).  Returns this rank's shard of O in q's dtype and the fp32 LSE, in the same layout as q.
    transport: "p2p" (copy-engine pulls from CUDA-IPC-mapped peer buffers; the default for CUDA tensors), "nccl"
    (send/recv rotation; the default for the CPU test seams), or an object with the pull-transport protocol of _P2PState.

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
    # the product path folds every step's partial into the running one inside the attention kernel's epilogue; with injected
    # arithmetic (the CPU tests) or fuse_merge=False the partials are merged by `merge` and cast at the end
    fused = on_gpu and _attn is None and _merge is None and _finalize is None and fuse_merge
    parts = _Partials(attn, merge, scale if fused else None)
    if zigzag and not causal:
        raise ValueError("zigzag sharding only makes sense for the causal ring")
    if zigzag and q.shape[-2] % 2:
        raise ValueError("a zig-zag shard holds two chunks of equal length")
    zz = zigzag and world > 1
    c = q.shape[-2] // 2

    def finish(o_acc, lse_acc):
        if _finalize is not None:
            return _finalize(o_acc), lse_acc
        if on_gpu and o_acc.dtype != q.dtype:
            return api.cast_to_16(o_acc, q.dtype), lse_acc
        return o_acc, lse_acc

    if world == 1:
        o, lse = attn(q, k, v, causal)
        return (_finalize(o) if _finalize else (api.cast_to_16(o, q.dtype) if (on_gpu and q.dtype != torch.float32) else o)), lse
    if transport is None:
        transport = "p2p" if (on_gpu and _attn is None) else "nccl"
    if isinstance(transport, str) and transport not in ("p2p", "nccl"):
        raise ValueError(f"unknown ring transport {transport!r}")
    if transport != "nccl":
        if transport == "p2p" and not on_gpu:
            raise ValueError("the p2p transport needs CUDA tensors")
        # a transport object (publish / prefetch / acquire / release / finish) is the CPU tests' stand-in for _P2PState
        st = _p2p_state(k, group) if transport == "p2p" else transport
        return finish(*_ring_pull(st, q, k, v, causal, zz, c, parts, rank, world))

    nxt = dist.get_global_rank(group, (rank + 1) % world) if group is not None else (rank + 1) % world
    prv = dist.get_global_rank(group, (rank - 1) % world) if group is not None else (rank - 1) % world
    k_cur, v_cur = k.contiguous(), v.contiguous()
    spare = [(torch.empty_like(k_cur), torch.empty_like(v_cur)) for _ in range(2)]  # receive buffers, ping-pong
    main = torch.cuda.current_stream(q.device) if on_gpu else None
    comm = torch.cuda.Stream(device=q.device) if on_gpu else None
    calls = ring_calls(rank, world, causal, zz)

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
        # local tile loop on the shard that is resident now (overlaps the transfer above); a causal ring with contiguous
        # shards has no call for a later rank's shard: every key of it is in the future of every local query
        for s_, slot, hq, keys, cz, is_last in calls:
            if s_ != src:
                continue
            q_ = q if hq is None else q[..., hq * c:(hq + 1) * c, :]
            k_s, v_s = (k_cur, v_cur) if keys == "all" else (k_cur[..., :c, :], v_cur[..., :c, :])
            parts.step(slot, q_, k_s, v_s, cz, is_last)
        if step < world - 1:
            for r in reqs:
                r.wait()
            if on_gpu:
                main.wait_stream(comm)
            k_cur, v_cur = k_nxt, v_nxt
    return finish(*parts.result(zz))
