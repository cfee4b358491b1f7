#!/usr/bin/env python
"""e2e (fa_forward_host: pinned host Q/K/V -> O in pinned host memory) under N ranks sharing one host, by chunk schedule.
torchrun --nproc-per-node N scripts/e2e_multi.py  — every rank runs the C2 workload on its own GPU; times are max over ranks."""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import flashattention_c_b200 as fab  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
cores = sorted(os.sched_getaffinity(0))
per = max(1, len(cores) // world)
os.sched_setaffinity(0, cores[local * per:(local + 1) * per] or cores)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = torch.Generator().manual_seed(rank)
q, k, v = (torch.randn(2, 8, 8192, 64, generator=g).pin_memory() for _ in range(3))
o = torch.empty_like(q).pin_memory()
runs = [("default (decay .4)", {}), ("equal x1", {"FA_B200_HOST_CHUNKS": "1"}), ("equal x2", {"FA_B200_HOST_CHUNKS": "2"}),
        ("equal x3", {"FA_B200_HOST_CHUNKS": "3"}), ("equal x4", {"FA_B200_HOST_CHUNKS": "4"}), ("decay .6", {"FA_B200_HOST_DECAY": ".6"}),
        ("sched 8,6,2", {"FA_B200_HOST_SCHED": "8,6,2"}), ("sched 10,5,1", {"FA_B200_HOST_SCHED": "10,5,1"})]
for name, env in runs:
    for kk in ("FA_B200_HOST_CHUNKS", "FA_B200_HOST_DECAY", "FA_B200_HOST_SCHED"):
        os.environ.pop(kk, None)
    os.environ.update(env)
    for _ in range(3):
        fab.attention_host(q, k, v, out=o)
    res = []
    for sync_each in (False, True):
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            if sync_each:
                dist.barrier()
            fab.attention_host(q, k, v, out=o)
        t = torch.tensor([(time.perf_counter() - t0) / 10], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res.append(float(t[0]) * 1e3)
    if rank == 0:
        print(f"world {world}  {name:20s} e2e ms/step: free-running {res[0]:.3f}   barrier before every step {res[1]:.3f}", flush=True)
dist.destroy_process_group()
