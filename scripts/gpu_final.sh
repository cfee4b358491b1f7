#!/bin/bash
# Round-end GPU session: pytest -m gpu, smoke, bench (own arm + reference arm), ncu launch list of the bench command,
# ncu --set full captures of the C2, C4 and fp32 d=128 (one-slot) kernels (read back with scripts/ncu_summary.py).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/final.log
echo "== pytest -m gpu" > $L
timeout 900 python -m pytest tests -q -m gpu >> $L 2>&1
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "== bench ours" >> $L
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2>> $L
cat gpurun_out/bench_ours.json >> $L
echo "== bench reference" >> $L
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>> $L
cat gpurun_out/bench_ref.json >> $L
echo "== ncu launch list" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu.json 2>> $L
echo "== ncu full (C2, fp32 d=128 one-slot, C4 kernels)" >> $L
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 2 -c 1 -f -o gpurun_out/prof_c2 \
   $H/fa_check f32 64 16 8192 0 0 2 0 >> $L 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 2 -c 1 -f -o gpurun_out/prof_f32_d128 \
   $H/fa_check f32 128 16 8192 0 0 2 0 >> $L 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 2 -c 1 -f -o gpurun_out/prof_c4 \
   $H/fa_check bf16 128 128 8192 0 0 2 0 >> $L 2>&1
grep -v "^==PROF==" $L | cut -c1-400 | tail -n 40
