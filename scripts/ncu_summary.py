#!/usr/bin/env python
"""Turn an `ncu --set full` report into a short markdown summary for profiles/ (run here, no GPU needed):

    python scripts/ncu_summary.py gpurun_out/prof_c4.ncu-rep "C4 bf16 d=128 N=8192" >> profiles/r01_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
]


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True, check=True).stdout


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    print(f"\n### {title}\n")
    print(f"kernel: `{vals[hdr.index('Kernel Name')][:120]}`\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {vals[i]} | {units[i]} |")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    h = src[1]
    data = src[2:]
    ix = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
    agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
    print("\nwarp-stall samples (all warps of the CTA, incl. the single-lane producer / MMA warps):\n")
    print("| reason | share |\n|---|---|")
    for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        print(f"| {s} | {100 * v / tot:.1f}% |")
    print("\nhottest SASS instructions:\n")
    print("| samples | share | instruction |\n|---|---|---|")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:8]:
        n = int(r[ix["# Samples"]] or 0)
        print(f"| {n} | {100 * n / tot:.1f}% | `{r[ix['Source']].strip()[:80]}` |")
    sass = [r[ix["Source"]].split()[0:2] for r in data]
    ops = {}
    for r in data:
        toks = r[ix["Source"]].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    want = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "MUFU", "HMMA", "FMNMX3"]
    print("\nSASS evidence (static instruction counts): " + ", ".join(f"{k}={ops.get(k, 0)}" for k in want))


if __name__ == "__main__":
    main()
