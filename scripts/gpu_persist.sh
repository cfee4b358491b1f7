#!/bin/bash
# A/B of the scheduling modes (FA_B200_PERSISTENT=1: one CTA per SM + atomic item counter, 0: one CTA per item),
# then pytest -m gpu, smoke and the bench lines.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/persist_ab.log
: > $L
run() { timeout 120 $H/fa_check "$@" >> $L 2>&1 || echo "  (exit $?)" >> $L; }
for m in 1 0; do
  export FA_B200_PERSISTENT=$m
  echo "#### FA_B200_PERSISTENT=$m" >> $L
  run f32 64 3 1000 1 0 3
  run f32 64 3 384 0 0 3
  run bf16 128 2 777 1 0 3
  run f32 32 4 512 0 0 3
  run bf16 64 2 300 1 0 3
  run f32 64 40 1024 1 0 3
  run bf16 64 300 512 0 0 3
  run f32 64 16 1024 0 0 30 0
  run f32 64 16 1024 1 0 30 0
  run f32 64 16 8192 0 0 20 0
  run f32 64 16 8192 1 0 20 0
  run f32 32 128 1024 0 0 30 0
  run f32 32 128 1024 1 0 30 0
  run bf16 64 128 1024 0 0 30 0
  run bf16 128 128 8192 0 0 10 0
  run bf16 128 128 8192 1 0 10 0
done
unset FA_B200_PERSISTENT
cut -c1-60,150-400 $L
if [ "$1" != "quick" ]; then
  echo "== pytest -m gpu" > gpurun_out/pytest.log
  timeout 900 python -m pytest tests -q -m gpu -x >> gpurun_out/pytest.log 2>&1
  tail -n 15 gpurun_out/pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
  timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2>> gpurun_out/pytest.log
  cat gpurun_out/bench_ours.json
fi
