#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/T1
FA_B200_TRACE=gpurun_out/trace_c2.txt timeout 120 $H/fa_check f32 64 16 8192 0 0 2 0 | tail -1 | cut -c1-200
FA_B200_TRACE=gpurun_out/trace_c4.txt timeout 120 $H/fa_check bf16 128 128 8192 0 0 2 0 | tail -1 | cut -c1-200
python scripts/trace_report.py gpurun_out/trace_c2.txt | head -48
echo ==== C4
python scripts/trace_report.py gpurun_out/trace_c4.txt | head -40
