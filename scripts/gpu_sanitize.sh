#!/bin/bash
# compute-sanitizer memcheck over the in-process GPU parity tests (global / shared out-of-bounds, misaligned accesses)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x \
  -k "not bench_script and not harness and not multi_gpu and not graph_capture and not full_size and not randomised and not back_to_back" \
  > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"
grep "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds" gpurun_out/sanitize_memcheck.log | sort | uniq -c | head -20
