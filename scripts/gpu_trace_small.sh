#!/bin/bash
# Timeline traces (FA_TRACE build in flashattention.c_b200/variants/T1) of the small BASELINE shapes C3 and C1:
# CTA 0's first 48 K/V steps = several whole items, so the item-to-item gaps are visible.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/T1
FA_B200_TRACE=gpurun_out/trace_c3.txt timeout 120 $H/fa_check f32 32 128 1024 0 0 2 0
FA_B200_TRACE=gpurun_out/trace_c1.txt timeout 120 $H/fa_check f32 64 16 1024 0 0 2 0
FA_B200_TRACE=gpurun_out/trace_c3_bigger.txt timeout 120 $H/fa_check f32 32 512 1024 0 0 2 0
FA_B200_TRACE=gpurun_out/trace_bf16_d64_n1024.txt timeout 120 $H/fa_check bf16 64 128 1024 0 0 2 0
ls -la gpurun_out/trace_*.txt
