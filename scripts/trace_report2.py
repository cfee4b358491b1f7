#!/usr/bin/env python
"""Compact per-phase report of a FA_TRACE dump for the split-P, single-issuer kernel (see FA_TRACE in fa_fwd_sm100.cuh)."""
import sys
import numpy as np
for f in sys.argv[1:]:
    a=np.loadtxt(f,dtype=np.uint64).reshape(4,-1,8).astype(np.int64)
    t0=a[a>0].min(); r=np.where(a>0,a-t0,-1)
    lo,hi=8,40
    print('#',f)
    for role,name in ((0,'softmax A'),(1,'softmax B')):
        x=r[role]
        per=np.diff(x[lo:hi+1,1]).mean()
        d=lambda a_,b_: np.mean(x[lo:hi,b_]-x[lo:hi,a_])
        print(f"{name}: period {per:.0f} | wait_S {d(0,1):.0f} ld+max {d(1,2):.0f} exp_h0 {d(2,3):.0f} arrive0 {d(3,4):.0f} exp_h1 {d(4,5):.0f} arrive1 {d(5,6):.0f} tail->next {np.mean(x[lo+1:hi+1,0]-x[lo:hi,6]):.0f}")
    for role,name in ((2,'MMA view tile A'),(3,'MMA view tile B')):
        x=r[role]
        d=lambda a_,b_: np.mean(x[lo:hi,b_]-x[lo:hi,a_])
        print(f"{name}: wait_P0 {d(0,1):.0f} issue_PV0 {d(1,2):.0f} wait_P1 {d(2,3):.0f} issue_PV1+S {d(3,6):.0f}")
    print('MMA warp gap B(j)->A(j+1): %.0f   A(j)->B(j): %.0f' % (np.mean(r[2,lo+1:hi+1,0]-r[3,lo:hi,6]), np.mean(r[3,lo:hi,0]-r[2,lo:hi,6])))
    for t in (0,1):
        sm,mm=r[t],r[2+t]
        print(f"tile {'AB'[t]}: P0 arrive -> MMA resumes {np.mean(mm[lo:hi,1]-sm[lo:hi,4]):.0f}; P1 arrive -> MMA resumes {np.mean(mm[lo:hi,3]-sm[lo:hi,6]):.0f}; S committed -> softmax wakes {np.mean(sm[lo+1:hi+1,1]-mm[lo:hi,6]):.0f}")
    print('raw step 10/11:')
    for role in range(4): print(' ', role, r[role,10].tolist(), r[role,11].tolist())
