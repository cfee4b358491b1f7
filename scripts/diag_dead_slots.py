"""Diagnostic: which (head, 128-row tile) of a many-item launch with dead slots is wrong, and what it holds instead."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import flashattention_c_b200 as fab  # noqa: E402

dev = torch.device("cuda:0")


def run(bh, n, d, dtype, causal=False, precise=False, **env):
    for k_, v_ in env.items():
        os.environ[k_] = v_
    g = torch.Generator(device="cpu").manual_seed(n + d)
    q, k, v = (torch.randn(bh, n, d, generator=g).to(dtype).to(dev) for _ in range(3))
    o = fab.attention(q, k, v, causal=causal, precise=precise)
    o_s = fab.attention(q, k, v, causal=causal, impl=fab.FA_IMPL_SIMT)
    torch.cuda.synchronize()
    err = (o.float() - o_s.float()).abs()
    tiles = (n + 127) // 128
    bad = []
    for h in range(bh):
        for t in range(tiles):
            e = float(err[h, t * 128:(t + 1) * 128].max())
            if e > 0.05:
                bad.append((h, t, round(e, 3)))
    print(f"bh={bh} n={n} d={d} {dtype} causal={causal} precise={precise} env={env}: max err {float(err.max()):.3e}, bad tiles {len(bad)} of {bh * tiles}")
    print("   first bad:", bad[:24])
    if bad:
        h, t, _ = bad[0]
        rows = slice(t * 128, (t + 1) * 128)
        badrows = (err[h, rows].amax(dim=-1) > 0.05).nonzero().flatten().tolist()
        print("   bad rows in first bad tile:", badrows[:10], "...", len(badrows), "rows")
        r = t * 128 + badrows[0]
        print("   got ", o[h, r, :6].float().tolist())
        print("   want", o_s[h, r, :6].float().tolist())
        # is the wrong row the right row of ANOTHER (head, row)?  search exact matches against the SIMT result
        tgt = o[h, r].float()
        dist = (o_s.float() - tgt).abs().amax(dim=-1)
        idx = torch.nonzero(dist < 2e-3)
        print("   matches elsewhere in the correct output:", idx[:5].tolist())
    for k_ in env:
        os.environ.pop(k_, None)


import ctypes
try:
    for rep in range(3):
        run(160, 384, 64, torch.float32)
except Exception as e:
    info = (ctypes.c_uint32 * 4)()
    fab.lib().fa_watchdog_info(info)
    print("FAILED:", str(e).splitlines()[0], "watchdog {tag, block, thread, parity} =", list(info))
    sys.exit(1)
run(160, 384, 64, torch.float32, causal=True)
run(160, 384, 64, torch.float32, FA_B200_TAIL_SPLIT="0")
run(160, 640, 64, torch.float32)
run(160, 384, 128, torch.float32)
run(160, 640, 128, torch.float32, causal=True)
run(160, 384, 32, torch.float32)
run(40, 1152, 64, torch.float32, precise=True)
run(40, 1152, 64, torch.float32, precise=True, causal=True)
run(160, 384, 128, torch.bfloat16)
run(160, 640, 256, torch.bfloat16)
