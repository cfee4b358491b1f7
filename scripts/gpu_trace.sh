#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
for v in T_split T_plain; do
  export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/$v
  FA_B200_TRACE=gpurun_out/trace_${v}_c4.txt timeout 120 $H/fa_check bf16 128 128 8192 0 0 2 0
  FA_B200_TRACE=gpurun_out/trace_${v}_c2.txt timeout 120 $H/fa_check f32 64 16 8192 0 0 2 0
done
ls -la gpurun_out/trace_*
