#!/bin/bash
# backward kernels on one GPU: tests, per-launch durations at the C4 shape, ncu --set full of the dK/dV and the dQ launch
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/prof
L=gpurun_out/bwd_prof.log
S=gpurun_out/bwd_ncu_summary.md
echo "== tests" > $L
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -m gpu > gpurun_out/pytest_bwd.log 2>&1; tail -4 gpurun_out/pytest_bwd.log >> $L
echo "== launch durations, C4 shape (B4 H32 N8192 d128 bf16), 2 backward passes" >> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fa_bwd --csv --log-file gpurun_out/bwd_launches.csv python scripts/bwd_one.py 4 32 8192 128 0 2 >> $L 2>&1
grep -v "^==" gpurun_out/bwd_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8 >> $L
echo "# backward kernels — ncu --set full (B200, --clock-control none), B1 H16 N4096 d128 bf16 non-causal" > $S
NCU="ncu --set full --clock-control none --import-source on -k regex:fa_bwd_sm100 -c 1 -f"
timeout 300 $NCU -s 0 -o /tmp/prof/bwd_dkv python scripts/bwd_one.py 1 16 4096 128 0 >> $L 2>&1
python scripts/ncu_summary.py /tmp/prof/bwd_dkv.ncu-rep "dK/dV launch (kDKV = true)" >> $S 2>> $L
timeout 300 $NCU -s 1 -o /tmp/prof/bwd_dq python scripts/bwd_one.py 1 16 4096 128 0 >> $L 2>&1
python scripts/ncu_summary.py /tmp/prof/bwd_dq.ncu-rep "dQ launch (kDKV = false)" >> $S 2>> $L
cp /tmp/prof/bwd_dkv.ncu-rep gpurun_out/bwd_dkv.ncu-rep
grep -v "^==PROF==\|^==WARNING==\|^$" $L | cut -c1-250 | tail -n 24
