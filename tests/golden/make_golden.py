"""Generates tests/golden/ref_kernel_*.npz ON A B200: outputs of the REFERENCE CUDA kernel (src/main.cpp +
src/flashattention.cu compiled for sm_100a from /root/reference into oracle/_ref/flash_ref_d{32,64,128}.so by
oracle/Makefile) on seeded inputs.  Run via `gpurun -- python tests/golden/make_golden.py`; the files land in
gpurun_out/golden/ and are then copied into tests/golden/ and committed.

The reference forward() hard-codes scaling = 1.0 (src/flashattention.cu:593, 600) and expects [B*H, N, d] fp32 with
N % 32 == 0 (Appendix A #4 of SURVEY.md), so every case obeys that.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import fa_oracle  # noqa: E402

CASES = [  # name, d, bh, n, causal, input std, seed
    ("d64_n128_full", 64, 2, 128, False, 1.0, 101),
    ("d64_n128_causal", 64, 2, 128, True, 1.0, 102),
    ("d64_n160_soft", 64, 1, 160, False, 0.35, 103),
    ("d32_n128_full", 32, 2, 128, False, 1.0, 104),
    ("d32_n96_causal", 32, 2, 96, True, 0.5, 105),
    ("d128_n96_causal", 128, 1, 96, True, 0.5, 106),
    ("d128_n64_full", 128, 2, 64, False, 0.3, 107),
]


def main():
    out = ROOT / "gpurun_out" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    for name, d, bh, n, causal, std, seed in CASES:
        ext = fa_oracle.load_ref_torch_ext(d)
        assert ext is not None, f"oracle/_ref/flash_ref_d{d}.so missing"
        rng = np.random.default_rng(seed)
        q, k, v = ((rng.standard_normal((bh, n, d), dtype=np.float32) * np.float32(std)).astype(np.float32) for _ in range(3))
        o = ext.forward(torch.from_numpy(q).cuda(), torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda(), causal)
        torch.cuda.synchronize()
        o = o.cpu().numpy()
        o64, _ = fa_oracle.f64(q, k, v, 1.0, causal)
        print(f"{name}: reference kernel vs fp64 oracle max abs err {np.abs(o - o64).max():.3e}")
        np.savez(out / f"ref_kernel_{name}.npz", q=q, k=k, v=v, o=o, causal=np.array(causal), d=np.array(d), seed=np.array(seed))


if __name__ == "__main__":
    main()
