#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/T1
FA_B200_TRACE=gpurun_out/trace_c2.txt timeout 120 $H/fa_check f32 64 16 8192 0 0 2 0 | tail -1
FA_B200_TRACE=gpurun_out/trace_c3.txt timeout 120 $H/fa_check f32 32 128 1024 0 0 2 0 | tail -1
python scripts/trace_rows.py gpurun_out/trace_c2.txt 10 18
