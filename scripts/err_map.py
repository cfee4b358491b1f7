#!/usr/bin/env python
"""Where does the tf32 error of the causal C1 case live?  Prints the error by row block and the worst locations,
for the tcgen05 path under a few settings and for the CUDA-core checker."""
import math
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def run_case(tag):
    import flashattention_c_b200 as fab
    from oracle import fa_oracle

    B, H, N, d = 2, 8, 1024, 64
    rng = lambda s: np.random.default_rng(s).standard_normal((B, H, N, d), dtype=np.float32)
    q, k, v = rng(11), rng(12), rng(13)
    tq, tk, tv = (torch.from_numpy(x).cuda() for x in (q, k, v))
    for causal in (True, False):
        ref, _ = fa_oracle.f64(q, k, v, 1 / math.sqrt(d), causal)
        for impl, name in ((1, "tcgen05"), (2, "simt")):
            o = fab.attention(tq, tk, tv, causal=causal, impl=impl).cpu().numpy()
            err = np.abs(o - ref)
            per_row = err.max(axis=(0, 1, 3))
            blocks = [float(per_row[i:i + 128].max()) for i in range(0, N, 128)]
            idx = np.unravel_index(np.argsort(err, axis=None)[-4:], err.shape)
            print(f"[{tag}] causal={causal} {name}: max {err.max():.3e}; by 128-row block: " + " ".join(f"{b:.1e}" for b in blocks))
            if impl == 1:
                for b_, h_, r_, c_ in zip(*idx):
                    print(f"      worst at b={b_} h={h_} row={r_} col={c_}: got {o[b_, h_, r_, c_]:+.6f} ref {ref[b_, h_, r_, c_]:+.6f} |ref|row max {np.abs(ref[b_, h_, r_]).max():.3f}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for tag, env in (("default", {}), ("no_tma_tf32", {"FA_B200_TMA_TF32": "0"}), ("no_split_wave", {"FA_B200_NO_SPLIT_WAVE": "1"})):
            e = dict(os.environ)
            e.update(env)
            subprocess.run([sys.executable, __file__, tag], env=e, check=False)
