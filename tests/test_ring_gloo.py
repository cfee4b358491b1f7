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
