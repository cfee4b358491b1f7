"""Timeline of one CTA of a backward launch from a FA_BWD_TRACE dump: python scripts/bwd_trace_report.py trace.txt
compute slots: 0 S ready, 1 exps done, 2 P handed over, 3 dP ready, 4 dS handed over
MMA slots:     0 loop top, 1 P seen, 2 dV + S(i+1) issued, 3 dS seen, 4 dK (dQ) + dP(i+1) issued
"""
import sys

rows = [[int(x) for x in l.split()] for l in open(sys.argv[1]) if l.strip()]
R = lambda role, step: rows[role * 32 + step]   # noqa: E731
t0 = min(x for r in rows for x in r if x)
rel = lambda x: (x - t0) if x else -1   # noqa: E731
print("step | compute h0: S ready, exps done, P handed, dP ready, dS handed | MMA: top, P seen, dV+S(i+1) issued, dS seen, dK+dP(i+1) issued | producer: slot free")
prev = None
n = 0
for s in range(32):
    c0, m, pr = R(0, s), R(2, s), R(3, s)
    if not c0[0]:
        break
    n += 1
    line = f"{s:3d} | " + " ".join(f"{rel(x):7d}" for x in c0[:5]) + " | " + " ".join(f"{rel(x):7d}" for x in m[:5]) + " | " + f"{rel(pr[0]):7d}"
    if prev is not None:
        line += f"   period {c0[0] - prev}"
    prev = c0[0]
    print(line)
rng = range(4, min(12, n - 1))
print("\ncompute h0 per step: S ready -> exps done", [R(0, s)[1] - R(0, s)[0] for s in rng], " -> P handed", [R(0, s)[2] - R(0, s)[1] for s in rng],
      " wait for dP", [R(0, s)[3] - R(0, s)[2] for s in rng], " dP ready -> dS handed", [R(0, s)[4] - R(0, s)[3] for s in rng],
      " dS handed -> next S ready", [R(0, s + 1)[0] - R(0, s)[4] for s in rng])
print("MMA thread per step: wait for P", [R(2, s)[1] - R(2, s)[0] for s in rng], " issue dV + S(i+1)", [R(2, s)[2] - R(2, s)[1] for s in rng],
      " wait for dS", [R(2, s)[3] - R(2, s)[2] for s in rng], " issue dK + dP(i+1)", [R(2, s)[4] - R(2, s)[3] for s in rng])
print("P handed -> seen by MMA thread", [R(2, s)[1] - R(0, s)[2] for s in rng], "  dS handed -> seen", [R(2, s)[3] - R(0, s)[4] for s in rng])
