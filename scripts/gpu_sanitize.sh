#!/bin/bash
# compute-sanitizer memcheck over the edge-case parity tests (global / shared out-of-bounds, misaligned accesses)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x \
  -k "ragged or cross_lengths or accumulate_mode or split_kv_across or causal_with_more or zero_padded or precise_mode_vs_oracle or tail_cta or strided_views" \
  > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"
grep -c "ERROR SUMMARY" gpurun_out/sanitize_memcheck.log
grep "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds" gpurun_out/sanitize_memcheck.log | sort | uniq -c | head -20
