#!/bin/bash
# Full GPU session: golden vectors from the reference kernel, pytest -m gpu, smoke, bench (ours + reference), ncu captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
echo "== golden" > gpurun_out/round.log
timeout 300 python tests/golden/make_golden.py >> gpurun_out/round.log 2>&1
mkdir -p tests/golden && cp gpurun_out/golden/*.npz tests/golden/ 2>/dev/null
echo "== pytest -m gpu" >> gpurun_out/round.log
timeout 900 python -m pytest tests -q -m gpu >> gpurun_out/round.log 2>&1
echo "== smoke" >> gpurun_out/round.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/round.log 2>&1
echo "== TMA tf32 conversion experiment" >> gpurun_out/round.log
for e in 0 1; do
  FA_B200_TMA_TF32=$e timeout 120 $H/fa_check f32 64 16 1024 0 0 5 >> gpurun_out/round.log 2>&1
  FA_B200_TMA_TF32=$e timeout 120 $H/fa_check f32 64 4 512 1 1.0 5 >> gpurun_out/round.log 2>&1
done
echo "== bench ours" >> gpurun_out/round.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2>> gpurun_out/round.log
cat gpurun_out/bench_ours.json >> gpurun_out/round.log
echo "== bench reference" >> gpurun_out/round.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/round.log
cat gpurun_out/bench_ref.json >> gpurun_out/round.log
echo "== ncu launch list" >> gpurun_out/round.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu.json 2>> gpurun_out/round.log
echo "== ncu full (C2 kernel)" >> gpurun_out/round.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 2 -c 1 -f -o gpurun_out/prof_c2 \
   $H/fa_check f32 64 16 8192 0 0 2 0 >> gpurun_out/round.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 2 -c 1 -f -o gpurun_out/prof_c4 \
   $H/fa_check bf16 128 128 8192 0 0 2 0 >> gpurun_out/round.log 2>&1
tail -n 60 gpurun_out/round.log
