#!/bin/bash
# backward kernels, evidence run: smoke, ncu --set full of both launches (summarised on the box), compute-sanitizer on a small problem
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/prof
L=gpurun_out/bwd_final.log
S=gpurun_out/bwd_ncu_summary.md
echo "== smoke" > $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "== launch durations, C4 shape (B4 H32 N8192 d128 bf16), 2 backward passes" >> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fa_bwd --csv --log-file gpurun_out/bwd_launches.csv python scripts/bwd_one.py 4 32 8192 128 0 2 >> $L 2>&1
grep -v "^==" gpurun_out/bwd_launches.csv | awk -F'","' '{print substr($5,1,60), $NF}' | tail -6 >> $L
echo "# backward kernels — ncu --set full (B200, --clock-control none), B4 H32 N8192 d128 bf16 non-causal (the C4 shape)" > $S
NCU="ncu --set full --clock-control none --import-source on -k regex:fa_bwd_sm100 -c 1 -f"
timeout 400 $NCU -s 0 -o /tmp/prof/bwd_dkv python scripts/bwd_one.py 4 32 8192 128 0 >> $L 2>&1
python scripts/ncu_summary.py /tmp/prof/bwd_dkv.ncu-rep "dK/dV launch (kDKV = true)" >> $S 2>> $L
timeout 400 $NCU -s 1 -o /tmp/prof/bwd_dq python scripts/bwd_one.py 4 32 8192 128 0 >> $L 2>&1
python scripts/ncu_summary.py /tmp/prof/bwd_dq.ncu-rep "dQ launch (kDKV = false)" >> $S 2>> $L
cp /tmp/prof/bwd_dkv.ncu-rep gpurun_out/bwd_dkv.ncu-rep
echo "== compute-sanitizer memcheck + racecheck, B1 H2 N300/260 d128 causal" >> $L
timeout 600 compute-sanitizer --tool memcheck python scripts/bwd_one.py 1 2 300 128 1 2>&1 | tail -4 >> $L
timeout 600 compute-sanitizer --tool racecheck python scripts/bwd_one.py 1 2 300 64 1 2>&1 | tail -4 >> $L
grep -v "^==PROF==\|^==WARNING==\|^$" $L | cut -c1-250 | tail -n 30
