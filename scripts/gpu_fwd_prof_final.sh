#!/bin/bash
# ncu --set full of the forward kernel on C2 / C3 / C4 with the shipped library (summarised on the box)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/prof
S=gpurun_out/fwd_ncu_summary_final.md
echo "# forward kernel — ncu --set full (B200, --clock-control none), shipped library (tf32 instances hand P over at key 96)" > $S
NCU="ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 2 -c 1 -f"
prof() { name=$1; title=$2; shift 2; timeout 300 $NCU -o /tmp/prof/$name python scripts/fwd_one.py "$@" 4 > /dev/null 2>&1; python scripts/ncu_summary.py /tmp/prof/$name.ncu-rep "$title" >> $S 2>/dev/null; }
prof c2 "C2: fp32->tf32 d=64 B2 H8 N=8192 non-causal (headline)" 2 8 8192 64 f32 0
prof c3 "C3: fp32->tf32 d=32 B8 H16 N=1024 non-causal" 8 16 1024 32 f32 0
prof c4 "C4: bf16 d=128 B4 H32 N=8192 non-causal" 4 32 8192 128 bf16 0
grep "^###\|gpu__time_duration\|tensor_cycles_active\|dram__bytes" $S
