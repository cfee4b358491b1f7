"""world_size-2 (and 4) CPU tests of the multi-process host logic over gloo: the ring rotation + log-sum-exp merge and
the B x H sharding.  The attention arithmetic is injected from the oracle (this is a test seam; the product path runs
the CUDA kernels), so what is exercised here is exactly the part that has no GPU dependence: who sends what to
whom in which step, buffer ping-pong, the merge order, causal shard skipping, and the shard arithmetic."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_attn(scale):
    from oracle import fa_oracle

    def attn(q, k, v, causal):
        o, lse = fa_oracle.f64(q.numpy(), k.numpy(), v.numpy(), scale, causal)
        return torch.from_numpy(o), torch.from_numpy(lse)

    def merge(o_acc, lse_acc, o_new, lse_new):
        o, l = fa_oracle.merge(o_acc.numpy(), lse_acc.numpy(), o_new.numpy(), lse_new.numpy())
        return torch.from_numpy(o), torch.from_numpy(l)

    return attn, merge


def _zigzag_worker(rank, world, port, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import flashattention_c_b200 as fab
        from oracle import fa_oracle

        B, H, N, d = 1, 2, 16 * 2 * world, 16
        rng = np.random.default_rng(200)
        q, k, v = (torch.from_numpy(rng.standard_normal((B, H, N, d), dtype=np.float32)) for _ in range(3))
        scale = 0.25
        attn, merge = _oracle_attn(scale)
        qs, ks, vs = (fab.zigzag_shard(t, rank, world).contiguous() for t in (q, k, v))
        o_loc, lse_loc = fab.ring_attention(qs, ks, vs, causal=True, scale=scale, zigzag=True, _attn=attn, _merge=merge,
                                            _finalize=lambda o: o)
        o_full, lse_full = fa_oracle.f64(q.numpy(), k.numpy(), v.numpy(), scale, True)
        o_ref = fab.zigzag_shard(torch.from_numpy(o_full), rank, world).numpy()
        lse_ref = fab.zigzag_shard(torch.from_numpy(lse_full).unsqueeze(-1), rank, world).squeeze(-1).numpy()
        # all ranks' shards put back together give the full-sequence result
        gathered = [torch.empty_like(o_loc) for _ in range(world)]
        dist.all_gather(gathered, o_loc.contiguous())
        err_full = float(np.abs(fab.zigzag_unshard(gathered, world).numpy() - o_full).max())
        ret[rank] = (float(np.abs(o_loc.numpy() - o_ref).max()), float(np.abs(lse_loc.numpy() - lse_ref).max()), err_full)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_zigzag_causal_ring_over_gloo(world):
    """Balanced causal ring: rank r holds chunks r and 2P-1-r; rotation, per-step plan and the two-accumulator merge."""
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_zigzag_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err_o, err_l, err_full = ret[rank]
        assert err_o < 1e-12 and err_l < 1e-12 and err_full < 1e-12, (rank, err_o, err_l, err_full)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_zigzag_plan_covers_the_causal_mask_exactly_and_is_balanced(world):
    """Host logic only: over all ring steps the plan of every rank visits each visible (query, key) pair exactly once and
    no masked pair, and every step of every rank costs the same number of pairs."""
    sys.path.insert(0, str(ROOT))
    import flashattention_c_b200 as fab

    c = 4
    n = 2 * world * c
    pos = [list(range(r * c, (r + 1) * c)) + list(range((2 * world - 1 - r) * c, (2 * world - r) * c)) for r in range(world)]
    for rank in range(world):
        seen = np.zeros((n, n), dtype=np.int64)
        for _, src in fab.ring.ring_schedule(rank, world):
            pairs = 0
            for half, keys, causal in fab.zigzag_step_plan(rank, src):
                qpos = pos[rank][half * c:(half + 1) * c]
                kpos = pos[src] if keys == "all" else pos[src][:c]
                for i, qp in enumerate(qpos):
                    for j, kp in enumerate(kpos):
                        if causal and j > i + (len(kpos) - len(qpos)):
                            continue
                        seen[qp, kp] += 1
                        pairs += 1
            assert pairs in (2 * c * c, c * (c + 1) // 2 + c * c + c * (c + 1) // 2), (rank, src, pairs)
        for qp in pos[rank]:
            expect = (np.arange(n) <= qp).astype(np.int64)
            assert (seen[qp] == expect).all(), (rank, qp)


def _worker(rank, world, port, causal, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import flashattention_c_b200 as fab
        from oracle import fa_oracle

        B, H, N, d = 1, 2, 32 * world, 16
        rng = np.random.default_rng(100)
        q, k, v = (torch.from_numpy(rng.standard_normal((B, H, N, d), dtype=np.float32)) for _ in range(3))
        scale = 0.25
        # ---- ring: sequence shards ----
        n_loc = N // world
        sl = slice(rank * n_loc, (rank + 1) * n_loc)
        attn, merge = _oracle_attn(scale)
        o_loc, lse_loc = fab.ring_attention(q[:, :, sl].contiguous(), k[:, :, sl].contiguous(), v[:, :, sl].contiguous(),
                                            causal=causal, scale=scale, _attn=attn, _merge=merge, _finalize=lambda o: o)
        o_full, lse_full = fa_oracle.f64(q.numpy(), k.numpy(), v.numpy(), scale, causal)
        err_o = float(np.abs(o_loc.numpy() - o_full[:, :, sl]).max())
        err_l = float(np.abs(lse_loc.numpy() - lse_full[:, :, sl]).max())
        # ---- B x H sharding: slices are disjoint, cover everything, and need no communication ----
        bh = B * H
        s, e = fab.bh_shard_range(bh, rank, world)
        counts = [None] * world
        dist.all_gather_object(counts, (s, e))
        ret[rank] = (err_o, err_l, counts)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,causal", [(2, False), (2, True), (4, False)])
def test_ring_attention_over_gloo(world, causal):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, causal, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err_o, err_l, counts = ret[rank]
        assert err_o < 1e-12, (rank, err_o)
        assert err_l < 1e-12, (rank, err_l)
        covered = [i for (s, e) in counts for i in range(s, e)]
        assert covered == list(range(2))


class _GatherPull:
    """CPU stand-in for the p2p transport (flashattention_c_b200.ring._P2PState): publishing = all-gathering every rank's K/V
    shard (on the GPU every owner's buffer is readable through its IPC mapping), a pull = indexing the gathered list.  It
    checks the protocol `_ring_pull` must follow: staging buffer i % 2 is free when pull i starts, every pull is acquired
    exactly once with the source it was started with, nothing is in flight at the end."""

    def __init__(self):
        self.busy, self.inflight, self.pulled = [None, None], {}, []

    def publish(self, kc, vc):
        world = dist.get_world_size()
        self.k_all = [torch.empty_like(kc) for _ in range(world)]
        self.v_all = [torch.empty_like(vc) for _ in range(world)]
        dist.all_gather(self.k_all, kc)
        dist.all_gather(self.v_all, vc)

    def prefetch(self, i, src):
        assert self.busy[i % 2] is None, f"staging buffer {i % 2} still in use by pull {self.busy[i % 2]}"
        self.busy[i % 2] = i
        self.inflight[i] = src

    def acquire(self, i, src):
        assert self.inflight.pop(i) == src
        self.pulled.append(src)
        return self.k_all[src], self.v_all[src]

    def release(self, i):
        assert self.busy[i % 2] == i
        self.busy[i % 2] = None

    def finish(self):
        assert not self.inflight and self.busy == [None, None]
        dist.barrier()


def _pull_worker(rank, world, port, causal, zigzag, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import flashattention_c_b200 as fab
        from oracle import fa_oracle

        B, H, N, d = 1, 2, 16 * 2 * world, 16
        rng = np.random.default_rng(300)
        q, k, v = (torch.from_numpy(rng.standard_normal((B, H, N, d), dtype=np.float32)) for _ in range(3))
        scale = 0.25
        attn, merge = _oracle_attn(scale)
        if zigzag:
            qs, ks, vs = (fab.zigzag_shard(t, rank, world).contiguous() for t in (q, k, v))
        else:
            n_loc = N // world
            qs, ks, vs = (t[:, :, rank * n_loc:(rank + 1) * n_loc].contiguous() for t in (q, k, v))
        tr = _GatherPull()
        for _ in range(2):   # a second call reuses the transport object, as the cached GPU state is reused
            o_loc, lse_loc = fab.ring_attention(qs, ks, vs, causal=causal, scale=scale, zigzag=zigzag, transport=tr, _attn=attn,
                                                _merge=merge, _finalize=lambda o: o)
        o_full, lse_full = fa_oracle.f64(q.numpy(), k.numpy(), v.numpy(), scale, causal)
        if zigzag:
            o_ref = fab.zigzag_shard(torch.from_numpy(o_full), rank, world).numpy()
            lse_ref = fab.zigzag_shard(torch.from_numpy(lse_full).unsqueeze(-1), rank, world).squeeze(-1).numpy()
        else:
            o_ref, lse_ref = o_full[:, :, rank * n_loc:(rank + 1) * n_loc], lse_full[:, :, rank * n_loc:(rank + 1) * n_loc]
        ret[rank] = (float(np.abs(o_loc.numpy() - o_ref).max()), float(np.abs(lse_loc.numpy() - lse_ref).max()), list(tr.pulled))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,causal,zigzag", [(2, False, False), (4, False, False), (4, True, False), (2, True, True), (4, True, True)])
def test_pull_transport_schedule_over_gloo(world, causal, zigzag):
    """The p2p ring's schedule (which owner is pulled in which step, prefetch one step ahead into a two-buffer ping-pong,
    causal shard skipping, zig-zag step plan on pulled shards) with the transport replaced by a CPU stand-in."""
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_pull_worker, args=(world, port, causal, zigzag, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err_o, err_l, pulled = ret[rank]
        assert err_o < 1e-12 and err_l < 1e-12, (rank, err_o, err_l)
        per_call = pulled[:len(pulled) // 2]
        assert pulled == per_call * 2
        if causal and not zigzag:
            assert per_call == [(rank - s) % world for s in range(1, world) if (rank - s) % world < rank]   # earlier ranks only
        else:
            assert per_call == [(rank - s) % world for s in range(1, world)]                                   # every other owner, ring order
