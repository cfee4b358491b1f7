#!/bin/bash
# First-contact GPU session: primitive probes, then end-to-end checks of the forward kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
: > gpurun_out/probe.log
for cfg in "tf32 64" "tf32 32" "bf16 128" "bf16 64"; do
  for mode in ss ts; do
    for v in 0 1 2 3; do
      timeout 30 $H/umma_probe $cfg $mode $v >> gpurun_out/probe.log 2>&1 || echo "  (exit $?)" >> gpurun_out/probe.log
    done
  done
done
: > gpurun_out/check.log
run() { echo "== fa_check $*" >> gpurun_out/check.log; timeout 120 $H/fa_check "$@" >> gpurun_out/check.log 2>&1 || echo "  (exit $?)" >> gpurun_out/check.log; }
run f32 64 2 256 0 0 5
run f32 64 2 256 1 0 5
run f32 64 2 1024 0 0 5
run f32 32 4 512 0 0 5
run bf16 128 2 512 0 0 5
run bf16 64 2 512 1 0 5
run f32 64 3 1000 0 0 5
run f32 64 3 1000 1 1.0 5
run bf16 128 2 777 1 0 5
run f32 64 16 1024 0 0 20
run f32 64 16 8192 0 0 20
run f32 64 16 8192 1 0 20
run f32 32 128 1024 0 0 20
run bf16 128 128 8192 0 0 10 0
run bf16 128 128 8192 1 0 10 0
run bf16 128 8 8192 0 0 5 1
echo "== test harness" >> gpurun_out/check.log
timeout 120 $H/test >> gpurun_out/check.log 2>&1 || echo "  (exit $?)" >> gpurun_out/check.log
tail -n 80 gpurun_out/probe.log
cat gpurun_out/check.log
