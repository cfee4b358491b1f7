#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
H=flashattention.c_b200/harness
for tool in racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 10 $H/fa_check f32 64 160 384 1 0 1 1 > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -8
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 10 $H/fa_check bf16 128 40 1152 0 0 1 1 >> gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -8
done
