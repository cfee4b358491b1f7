#!/bin/bash
# CTA timelines of the two backward launches (FA_BWD_TRACE build in variants/BT): scripts/gpu_bwd_trace.sh [B H N d causal]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export FA_B200_LIB=$PWD/flashattention.c_b200/variants/BT/libfa_b200.so
ARGS="${@:-4 32 8192 128 0}"
FA_B200_BWD_TRACE=gpurun_out/bwd_trace_dkv.txt timeout 120 python scripts/bwd_one.py $ARGS
FA_B200_BWD_TRACE=gpurun_out/bwd_trace_dq.txt FA_B200_BWD_TRACE_DQ=1 timeout 120 python scripts/bwd_one.py $ARGS
echo "=== dK/dV launch ($ARGS)"; python scripts/bwd_trace_report.py gpurun_out/bwd_trace_dkv.txt | head -40
echo "=== dQ launch ($ARGS)"; python scripts/bwd_trace_report.py gpurun_out/bwd_trace_dq.txt | head -40
