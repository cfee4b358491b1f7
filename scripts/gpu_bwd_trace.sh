#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export FA_B200_LIB=$PWD/flashattention.c_b200/variants/BT/libfa_b200.so
FA_B200_BWD_TRACE=gpurun_out/bwd_trace_dkv.txt timeout 120 python scripts/bwd_one.py 4 32 8192 128 0
FA_B200_BWD_TRACE=gpurun_out/bwd_trace_dq.txt FA_B200_BWD_TRACE_DQ=1 timeout 120 python scripts/bwd_one.py 4 32 8192 128 0
echo "=== dK/dV launch"; python scripts/bwd_trace_report.py gpurun_out/bwd_trace_dkv.txt | head -40
echo "=== dQ launch"; python scripts/bwd_trace_report.py gpurun_out/bwd_trace_dq.txt | head -40
