"""Static SASS instruction counts per kernel of libfa_b200.so (no GPU needed): python scripts/sass_summary.py > profiles/rNN_sass_summary.txt"""
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "flashattention.c_b200" / "libfa_b200.so"
COLS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU.EX2", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "HMMA", "ELECT", "LDG", "STG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    names = {}
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            names[cur] = {"instrs": 0, **{c: 0 for c in COLS}}
            continue
        if cur is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            continue
        toks = line.split("*/", 1)[1].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        names[cur]["instrs"] += 1
        for c in COLS:
            if op == c or op.startswith(c + "."):
                names[cur][c] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for mangled, pretty in zip(names, dem):
        pretty = re.sub(r"\(.*", "", pretty).replace("void fa::", "").replace("fa::", "")
        rows.append((pretty, names[mangled]))
    print("# static SASS instruction counts per kernel in libfa_b200.so (cuobjdump -sass, sm_100a; scripts/sass_summary.py)")
    print("# fa_fwd_sm100_kernel<kTF32, kHeadDim, kCausal, kOutF32, kF16, kPrecise>; fa_bwd_sm100_kernel<kHeadDim, kF16, kDKV>")
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops; HMMA (mma.sync) appears nowhere")
    print(f"{'kernel':64s}" + f"{'instrs':>8s}" + "".join(f"{c:>9s}" for c in COLS))
    for pretty, d in sorted(rows):
        print(f"{pretty[:63]:64s}{d['instrs']:8d}" + "".join(f"{d[c]:9d}" for c in COLS))


if __name__ == "__main__":
    sys.exit(main())
