"""Timeline of one CTA of a backward launch from a FA_BWD_TRACE dump: python scripts/bwd_trace_report.py trace.txt"""
import sys

rows = [[int(x) for x in l.split()] for l in open(sys.argv[1]) if l.strip()]
R = lambda role, step: rows[role * 32 + step]   # noqa: E731
t0 = min(x for r in rows for x in r if x)
rel = lambda x: (x - t0) if x else -1   # noqa: E731
print("step | compute h0: S ready, ld done, math done, st done, arrived | h1: S ready .. arrived | MMA: top, SD(i+1) issued, P seen, acc issued | producer: slot free")
prev = None
for s in range(32):
    c0, c1, m, pr = R(0, s), R(1, s), R(2, s), R(3, s)
    if not c0[0]:
        break
    line = f"{s:3d} | " + " ".join(f"{rel(x):7d}" for x in c0[:5]) + " | " + " ".join(f"{rel(x):7d}" for x in c1[:5]) + " | " + \
        " ".join(f"{rel(x):7d}" for x in m[:4]) + " | " + f"{rel(pr[0]):7d}"
    if prev is not None:
        line += f"   period {c0[0] - prev}"
    prev = c0[0]
    print(line)
print("\nper step (h0): wait->ld", [R(0, s)[1] - R(0, s)[0] for s in range(4, 12)], " math", [R(0, s)[2] - R(0, s)[1] for s in range(4, 12)],
      " st", [R(0, s)[3] - R(0, s)[2] for s in range(4, 12)])
print("P arrive -> MMA sees P:", [R(2, s)[2] - max(R(0, s)[4], R(1, s)[4]) for s in range(4, 12)])
print("acc issued -> next S ready at compute (h0):", [R(0, s + 2)[0] - R(2, s)[3] for s in range(4, 12)])
print("MMA thread, issue of S, dP of step s: waited for the stage's tiles (cycles), then issued 16 MMAs + commit (cycles):",
      [(R(2, s)[4] - R(2, s - 1)[0], R(2, s - 1)[1] - R(2, s)[4]) for s in range(5, 13)])
