#!/usr/bin/env python
"""CTA-level milestones of a FA_TRACE dump (FA_TRACE_MISC rows) plus per-step S-ready times: where a short-sequence
CTA spends its life (setup, first loads, K/V loop, epilogue)."""
import sys
import numpy as np
for f in sys.argv[1:]:
    a = np.loadtxt(f, dtype=np.uint64).reshape(4, -1, 8).astype(np.int64)
    t0 = a[a > 0].min()
    r = np.where(a > 0, a - t0, -1)
    print('#', f)
    print('  softmax A misc [entry, setup, loop done, O final, staged, store read, store done]:', r[0, -1, :7].tolist())
    print('  softmax B misc:', r[1, -1, :7].tolist())
    print('  MMA warp  misc [entry, Q full, K0 full, S0 issued]:', r[2, -1, :4].tolist())
    for role in (0, 1):
        x = r[role, :-1]
        n = int((x[:, 1] > 0).sum())
        print(f'  softmax {"AB"[role]}: {n} steps; S ready at', x[:n, 1].tolist(), '; last P arrive at', x[:n, 6].tolist())
