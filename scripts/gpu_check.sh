#!/bin/bash
# Correctness + timing sweep with the torch-free harness, then (unless "quick") pytest -m gpu and the bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/check3.log
: > $L
run() { echo "== fa_check $*" >> $L; timeout 120 $H/fa_check "$@" >> $L 2>&1 || echo "  (exit $?)" >> $L; }
run f32 64 2 256 0 0 5
run f32 64 2 256 1 0 5
run f32 32 4 512 1 0 5
run bf16 128 2 512 0 0 5
run bf16 64 2 512 1 0 5
run f32 64 3 1000 0 0 5
run f32 64 3 1000 1 1.0 5
run bf16 128 2 777 1 0 5
run f32 64 3 65 1 0 5
run f32 64 16 1024 0 0 20
run f32 64 16 8192 0 0 20
run f32 64 16 8192 1 0 20
run f32 32 128 1024 0 0 20
run bf16 128 128 8192 0 0 10 0
run bf16 128 128 8192 1 0 10 0
run bf16 128 8 8192 0 0 5 1
cat $L
if [ "$1" != "quick" ]; then
  echo "== pytest -m gpu" > gpurun_out/pytest.log
  timeout 900 python -m pytest tests -q -m gpu >> gpurun_out/pytest.log 2>&1
  tail -n 30 gpurun_out/pytest.log
  timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2>> gpurun_out/pytest.log
  cat gpurun_out/bench_ours.json
fi
