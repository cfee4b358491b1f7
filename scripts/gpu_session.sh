#!/bin/bash
# One lean GPU session: pytest -m gpu, smoke, bench (own arm + reference arm), ncu launch list of the bench command,
# harness timings of the wide-row (one-slot) instances against the CUDA-core kernel.  Everything lands in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/session.log
echo "== pytest -m gpu" > $L
timeout 900 python -m pytest tests -q -m gpu --durations=8 -x >> $L 2>&1
echo "pytest exit $?" >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "== bench ours" >> $L
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2>> $L
cat gpurun_out/bench_ours.json >> $L
echo "== bench reference" >> $L
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>> $L
cat gpurun_out/bench_ref.json >> $L
echo "== wide-row instances (tcgen05 one-slot) vs CUDA-core kernel" >> $L
for a in "f32 128 16 8192 0" "f32 128 16 8192 1" "bf16 256 16 8192 0" "f32 96 16 4096 0" "bf16 96 64 4096 0" "bf16 32 128 1024 0"; do
  timeout 120 $H/fa_check $a 0 5 1 >> $L 2>&1 || echo "  (exit $?)" >> $L
done
FA_CHECK_TIME_IMPL=2 timeout 120 $H/fa_check f32 128 16 8192 0 0 3 0 >> $L 2>&1   # the CUDA-core kernel on the same shape
echo "== ncu launch list" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu.json 2>> $L
grep -v "^==PROF==" $L | cut -c1-600 | tail -n 70
