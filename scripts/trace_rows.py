#!/usr/bin/env python
"""Raw view of a FA_TRACE dump: per K/V step, the softmax slots' and the MMA warp's time stamps relative to the first one.
softmax (roles 0/1): 0 wait S | 1 S ready | 2 max done | 3,4 P piece 0 written, arrived | 5,6 P piece 1 written, arrived
MMA warp (roles 2/3 = slot A/B): 0 before P wait | 1 piece 0 seen | 2 P V piece 0 issued | 3 piece 1 seen | 6 P V issued + commits | 7 S_t(step) issued + committed"""
import sys

import numpy as np

a = np.loadtxt(sys.argv[1], dtype=np.uint64).reshape(4, -1, 8).astype(np.int64)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8, 20)
t0 = a[a > 0].min()
r = np.where(a > 0, a - t0, -1)
for j in range(lo, hi):
    print(f"step {j:2d}  smA {r[0, j, :7].tolist()}  smB {r[1, j, :7].tolist()}")
    print(f"         mmA {r[2, j].tolist()}  mmB {r[3, j].tolist()}")
per = [np.diff(r[x, lo:hi, 1]).mean() for x in (0, 1)]
print("period (S ready -> S ready): slot A %.0f, slot B %.0f" % tuple(per))
for x, nm in ((0, "A"), (1, "B")):
    print(f"softmax {nm}: wait S {np.mean(r[x, lo:hi, 1] - r[x, lo:hi, 0]):.0f}  S ready -> P1 arrived {np.mean(r[x, lo:hi, 6] - r[x, lo:hi, 1]):.0f}")
    print(f"   S_{nm}(j) committed -> softmax {nm}(j) sees it {np.mean(r[x, lo:hi, 1] - r[2 + x, lo:hi, 7]):.0f};  P1 arrived -> MMA sees it {np.mean(r[2 + x, lo:hi, 3] - r[x, lo:hi, 6]):.0f}")
