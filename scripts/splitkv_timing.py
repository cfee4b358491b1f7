#!/usr/bin/env python
"""Decode-like launches (few query rows, long key sequence): split-KV across CTAs off (FA_B200_KV_SPLIT=0) vs automatic.
CUDA events, L2 flushed before every repetition, median of 20."""
import math
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import flashattention_c_b200 as fab  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t_ms(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for dtype, d, bh, nq, nk, causal in ((torch.float32, 64, 4, 128, 8192, False), (torch.float32, 64, 16, 128, 8192, False),
                                     (torch.bfloat16, 128, 8, 1, 16384, False), (torch.bfloat16, 128, 8, 128, 16384, False),
                                     (torch.bfloat16, 128, 32, 128, 8192, False), (torch.bfloat16, 128, 32, 1, 131072, False),
                                     (torch.bfloat16, 128, 4, 512, 32768, True), (torch.float32, 64, 16, 1024, 1024, False)):
    q = torch.randn(bh, nq, d, device=dev).to(dtype)
    k, v = (torch.randn(bh, nk, d, device=dev).to(dtype) for _ in range(2))
    out = torch.empty_like(q)
    res = {}
    for mode in ("0", "auto"):
        if mode == "auto":
            os.environ.pop("FA_B200_KV_SPLIT", None)
        else:
            os.environ["FA_B200_KV_SPLIT"] = mode
        before = fab.launch_count()
        fab.attention(q, k, v, causal=causal, out=out)
        res[mode + "_launches"] = fab.launch_count() - before
        res[mode] = t_ms(lambda: fab.attention(q, k, v, causal=causal, out=out))
    fl = 4.0 * bh * nq * nk * d * (0.5 if causal else 1.0)
    print(f"{str(dtype).split('.')[-1]:9s} d={d:3d} B*H={bh:3d} n_q={nq:5d} n_k={nk:6d} causal={int(causal)}: unsplit {res['0'] * 1e3:8.1f} us   "
          f"auto {res['auto'] * 1e3:8.1f} us ({res['auto_launches']} launches)   x{res['0'] / res['auto']:.2f}   "
          f"K+V bytes {2 * bh * nk * d * k.element_size() / 1e6:.0f} MB -> {2 * bh * nk * d * k.element_size() / res['auto'] / 1e6:.0f} GB/s", flush=True)
