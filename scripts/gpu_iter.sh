#!/bin/bash
# quick iteration: correctness + timing sweep (default build), traces of C1/C3 (variants/T), optional pytest
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/iter.log
: > $L
run() { timeout 120 $H/fa_check "$@" >> $L 2>&1 || echo "  (exit $?)" >> $L; }
run f32 64 3 1000 1 0 3
run bf16 128 2 777 1 0 3
run f32 32 4 512 0 0 3
run bf16 64 300 512 1 0 3
run f32 64 16 1024 0 0 30 0
run f32 64 16 1024 1 0 30 0
run f32 64 16 8192 0 0 20 0
run f32 64 16 8192 1 0 20 0
run f32 32 128 1024 0 0 30 0
run f32 32 128 1024 1 0 30 0
run bf16 64 128 1024 0 0 30 0
run bf16 128 128 8192 0 0 10 0
run bf16 128 128 8192 1 0 10 0
cut -c1-60,150-400 $L
if [ -d flashattention.c_b200/variants/T ]; then
  export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/T
  FA_B200_TRACE=gpurun_out/trace_c1.txt timeout 120 $H/fa_check f32 64 16 1024 0 0 2 0 > /dev/null
  FA_B200_TRACE=gpurun_out/trace_c3.txt timeout 120 $H/fa_check f32 32 128 1024 0 0 2 0 > /dev/null
  python scripts/trace_misc.py gpurun_out/trace_c1.txt gpurun_out/trace_c3.txt
  unset LD_LIBRARY_PATH
fi
if [ "$1" = "full" ]; then
  timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
  timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2>gpurun_out/bench_err.log
  cat gpurun_out/bench_ours.json
fi
