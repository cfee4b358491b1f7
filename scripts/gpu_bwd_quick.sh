#!/bin/bash
# backward kernels, quick loop: bring-up check, tests, timing
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/bwd_check.py > gpurun_out/bwd_check.log 2>&1; tail -3 gpurun_out/bwd_check.log
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -m gpu > gpurun_out/pytest_bwd.log 2>&1; tail -3 gpurun_out/pytest_bwd.log
timeout 300 python scripts/bwd_timing.py > gpurun_out/bwd_timing.log 2>&1; cat gpurun_out/bwd_timing.log
