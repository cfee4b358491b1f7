#!/bin/bash
# pytest -m gpu (full), smoke, bench (ours + reference), ncu launch list + full captures of the C2 and C4 kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
echo "== pytest -m gpu" > gpurun_out/pytest.log
timeout 900 python -m pytest tests -q -m gpu >> gpurun_out/pytest.log 2>&1
tail -n 25 gpurun_out/pytest.log
echo "== smoke" > gpurun_out/round.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/round.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2>> gpurun_out/round.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/round.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu.json 2>> gpurun_out/round.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 2 -c 1 -f -o gpurun_out/prof_c2 \
   $H/fa_check f32 64 16 8192 0 0 2 0 >> gpurun_out/round.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 2 -c 1 -f -o gpurun_out/prof_c4 \
   $H/fa_check bf16 128 128 8192 0 0 2 0 >> gpurun_out/round.log 2>&1
grep -v "^==PROF\|^==WARN" gpurun_out/round.log | tail -n 12 | cut -c1-600
cat gpurun_out/bench_ours.json | cut -c1-900
cat gpurun_out/bench_ref.json | cut -c1-600
