#!/usr/bin/env python
"""Summarise a FA_TRACE timeline dump (see FA_TRACE in csrc/fa_fwd_sm100.cuh): per-KV-tile phase durations in SM cycles
for CTA 0 — softmax warpgroups A/B (roles 0/1) and the MMA thread's view of tiles A/B (roles 2/3)."""
import sys

import numpy as np


def main(path, lo=8, hi=40):
    a = np.loadtxt(path, dtype=np.uint64).reshape(4, -1, 8).astype(np.int64)
    t0 = a[a > 0].min()
    r = np.where(a > 0, a - t0, -1)
    sm_names = ["wait_S(0->1)", "ld+max(1->2)", "exp_h0(2->3)", "st_wait+arrive0(3->4)", "exp_h1(4->5)", "st_wait+arrive1(5->6)"]
    print(f"# {path}: steady-state means over KV tiles [{lo},{hi}) in SM cycles")
    for role, name in ((0, "softmax A"), (1, "softmax B")):
        x = r[role, lo:hi]
        per = np.diff(r[role, lo:hi + 1, 1]).mean()
        print(f"{name}: period {per:.0f}")
        for k, nm in enumerate(sm_names):
            if (x[:, k + 1] >= 0).all() and (x[:, k] >= 0).all():
                print(f"   {nm:26s} {np.mean(x[:, k + 1] - x[:, k]):8.0f}")
    mm_names = ["wait_P0(0->1)", "issue_PV_h0(1->2)", "wait_P1(2->3)", "issue_PV_h1(3->4)", "to_S_issue(4->5)", "issue_S+commit(5->6)"]
    for role, name in ((2, "MMA thread, tile A"), (3, "MMA thread, tile B")):
        x = r[role, lo:hi]
        print(name)
        for k, nm in enumerate(mm_names):
            if (x[:, k + 1] >= 0).all() and (x[:, k] >= 0).all():
                print(f"   {nm:26s} {np.mean(x[:, k + 1] - x[:, k]):8.0f}")
    # cross-role latencies: P arrive (softmax slot 6, or 4 for half 0) -> MMA sees it (slot 3 / 1); S commit issue (MMA slot 6) -> softmax wake (next step slot 1)
    for t in (0, 1):
        sm, mm = r[t], r[2 + t]
        if (sm[lo:hi, 6] >= 0).all() and (mm[lo:hi, 3] >= 0).all():
            print(f"tile {'AB'[t]}: last P arrive -> MMA thread resumes      {np.mean(mm[lo:hi, 3] - sm[lo:hi, 6]):8.0f}")
        if (mm[lo:hi, 6] >= 0).all():
            print(f"tile {'AB'[t]}: S(j+1) issued+committed -> softmax wakes {np.mean(sm[lo + 1:hi + 1, 1] - mm[lo:hi, 6]):8.0f}   (= MMA execution + commit + wake-up)")
            print(f"tile {'AB'[t]}: last P arrive -> softmax wakes on S(j+1) {np.mean(sm[lo + 1:hi + 1, 1] - sm[lo:hi, 6]):8.0f}")
    # MMA warp between two K/V steps: end of B's step (role 3 slot 6) -> ring slots released (role 3 slot 7) -> next step's
    # K/V landed (role 2 slot 7) -> step start (role 2 slot 0)
    if (r[3, lo:hi, 7] >= 0).all() and (r[2, lo + 1:hi + 1, 7] >= 0).all():
        print(f"MMA warp: B step end -> slots released {np.mean(r[3, lo:hi, 7] - r[3, lo:hi, 6]):6.0f} | released -> next K/V landed "
              f"{np.mean(r[2, lo + 1:hi + 1, 7] - r[3, lo:hi, 7]):6.0f} | landed -> A step start {np.mean(r[2, lo + 1:hi + 1, 0] - r[2, lo + 1:hi + 1, 7]):6.0f}")
        if (r[2, lo + 1:hi + 1, 5] >= 0).all() and (r[3, lo + 1:hi + 1, 5] >= 0).all():
            print(f"MMA warp: released -> V_j wait done {np.mean(r[2, lo + 1:hi + 1, 5] - r[3, lo:hi, 7]):6.0f} | -> K_(j+1) wait done "
                  f"{np.mean(r[3, lo + 1:hi + 1, 5] - r[2, lo + 1:hi + 1, 5]):6.0f} | -> fence done {np.mean(r[2, lo + 1:hi + 1, 7] - r[3, lo + 1:hi + 1, 5]):6.0f}")
        print(f"MMA warp: A step end -> B step start {np.mean(r[3, lo:hi, 0] - r[2, lo:hi, 6]):6.0f} | A step {np.mean(r[2, lo:hi, 6] - r[2, lo:hi, 0]):6.0f} | B step {np.mean(r[3, lo:hi, 6] - r[3, lo:hi, 0]):6.0f}")
    # producer: TMA issue time of K_j (role 2 slot 4, row j) / V_j (role 3 slot 4, row j) vs the MMA warp seeing V_j and K_{j+1}
    # landed (role 2 slot 7, row j)
    if (r[2, lo:hi + 1, 4] >= 0).all() and (r[2, lo:hi, 7] >= 0).all():
        print(f"producer: K_(j+1) TMA issued -> MMA warp has V_j and K_(j+1) {np.mean(r[2, lo:hi, 7] - r[2, lo + 1:hi + 1, 4]):6.0f} | "
              f"V_j issued -> same {np.mean(r[2, lo:hi, 7] - r[3, lo:hi, 4]):6.0f} | slots of step j released -> K_(j+3) issued "
              f"{np.mean(r[2, lo + 3:hi + 3, 4] - r[3, lo:hi, 7]):6.0f}")
    print("raw, first rows (role, step, slots):")
    for role in range(4):
        for j in range(lo, lo + 3):
            print(role, j, r[role, j].tolist())


if __name__ == "__main__":
    main(sys.argv[1])
