#!/bin/bash
# timeline traces (FA_TRACE build in variants/T) of the short-sequence shapes C1 and C3
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/T
FA_B200_TRACE=gpurun_out/trace_c1.txt timeout 120 $H/fa_check f32 64 16 1024 0 0 2 0 > /dev/null
FA_B200_TRACE=gpurun_out/trace_c3.txt timeout 120 $H/fa_check f32 32 128 1024 0 0 2 0 > /dev/null
FA_B200_TRACE=gpurun_out/trace_c3b.txt timeout 120 $H/fa_check bf16 64 128 1024 0 0 2 0 > /dev/null
python scripts/trace_misc.py gpurun_out/trace_c1.txt gpurun_out/trace_c3.txt gpurun_out/trace_c3b.txt
