"""Bring-up check of the backward kernels on a GPU: every case against the fp64 oracle (test infrastructure), errors printed."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import fa_oracle as oracle  # noqa: E402
import flashattention_c_b200 as fab  # noqa: E402

CASES = [
    # B, H, Hk, nq, nk, d, causal, dtype
    (1, 1, 1, 128, 128, 64, False, torch.bfloat16),
    (1, 1, 1, 128, 128, 128, False, torch.bfloat16),
    (1, 2, 2, 256, 256, 64, True, torch.bfloat16),
    (1, 2, 2, 200, 333, 64, True, torch.bfloat16),
    (2, 4, 2, 192, 320, 128, False, torch.bfloat16),
    (1, 2, 1, 77, 130, 128, True, torch.float16),
    (1, 3, 3, 300, 100, 64, True, torch.bfloat16),
    (1, 2, 2, 640, 640, 96, True, torch.bfloat16),
    (1, 2, 2, 1024, 1024, 32, False, torch.float16),
    (1, 2, 2, 1100, 900, 128, False, torch.bfloat16),
]


def run(case, seed=0):
    B, H, Hk, nq, nk, d, causal, dt = case
    g = torch.Generator().manual_seed(1234 + seed)
    q = torch.randn(B, H, nq, d, generator=g).to(dt)
    k = torch.randn(B, Hk, nk, d, generator=g).to(dt)
    v = torch.randn(B, Hk, nk, d, generator=g).to(dt)
    do = torch.randn(B, H, nq, d, generator=g).to(dt)
    scale = 1.0 / d ** 0.5
    qd, kd, vd, dod = (t.cuda() for t in (q, k, v, do))
    o, lse = fab.attention(qd, kd, vd, causal=causal, scale=scale, return_lse=True)
    dq, dk, dv = fab.attention_backward(qd, kd, vd, o, lse, dod, causal=causal, scale=scale)
    torch.cuda.synchronize()
    rq, rk, rv = oracle.backward_f64(q.float().numpy(), k.float().numpy(), v.float().numpy(), do.float().numpy(), scale=scale, causal=causal)
    out = []
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        gotn = got.float().cpu().numpy().astype(np.float64)
        err = np.abs(gotn - ref).max() / max(np.abs(ref).max(), 1e-30)
        out.append((name, err, bool(np.isfinite(gotn).all())))
    return out


if __name__ == "__main__":
    bad = 0
    for c in CASES:
        t0 = time.time()
        try:
            res = run(c)
        except Exception as e:  # noqa: BLE001
            print("CASE", c, "EXC", repr(e)[:300], flush=True)
            import ctypes
            info = (ctypes.c_uint32 * 4)()
            fab.lib().fa_watchdog_info(info)
            print("  watchdog", list(info), flush=True)
            bad += 1
            break
        tol = 3e-2 if c[7] == torch.bfloat16 else 6e-3
        ok = all(e < tol and fin for _, e, fin in res)
        bad += 0 if ok else 1
        print("CASE", c[:7], str(c[7]).split(".")[-1], " ".join(f"{n}={e:.2e}{'' if fin else '(nonfinite)'}" for n, e, fin in res), "OK" if ok else "FAIL",
              f"{time.time() - t0:.1f}s", flush=True)
    print("BAD", bad)
    sys.exit(1 if bad else 0)
