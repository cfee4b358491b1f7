#!/bin/bash
# A/B test of kernel variants built into flashattention.c_b200/variants/<name>/libfa_b200.so (+ timeline trace of variant T)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/ab.log
: > $L
for v in $(ls flashattention.c_b200/variants); do
  export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/$v
  if [[ "$v" == T* ]]; then
    FA_B200_TRACE=gpurun_out/trace_c4.txt timeout 120 $H/fa_check bf16 128 128 8192 0 0 2 0 > /dev/null
    FA_B200_TRACE=gpurun_out/trace_c2.txt timeout 120 $H/fa_check f32 64 16 8192 0 0 2 0 > /dev/null
    continue
  fi
  echo "#### variant $v" >> $L
  run() { timeout 120 $H/fa_check "$@" >> $L 2>&1 || echo "  (exit $?)" >> $L; }
  run f32 64 3 1000 1 0 3
  run bf16 128 2 777 1 0 3
  run f32 32 4 512 0 0 3
  run bf16 64 2 300 1 0 3
  run f32 64 16 1024 0 0 20 0
  run f32 64 16 8192 0 0 20 0
  run f32 64 16 8192 1 0 20 0
  run f32 32 128 1024 0 0 20 0
  run bf16 128 128 8192 0 0 10 0
  run bf16 128 128 8192 1 0 10 0
done
cat $L
