#!/bin/bash
# tf32 P*V bring-up: which shared-memory layout does tcgen05 accept for a 32-bit MN-major B operand,
# and does A-from-TMEM work for kind::tf32?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
: > gpurun_out/probe2.log
for cfg in "tf32 64" "tf32 32"; do
  for mode in sv ts; do
    for v in 0 1 2 3 4; do
      timeout 30 $H/umma_probe $cfg $mode $v >> gpurun_out/probe2.log 2>&1 || echo "  (exit $?)" >> gpurun_out/probe2.log
    done
  done
done
timeout 30 $H/umma_probe bf16 128 sv 0 >> gpurun_out/probe2.log 2>&1
: > gpurun_out/check2.log
run() { echo "== [V_VARIANT=${FA_B200_V_VARIANT:-default}] fa_check $*" >> gpurun_out/check2.log; timeout 120 $H/fa_check "$@" >> gpurun_out/check2.log 2>&1 || echo "  (exit $?)" >> gpurun_out/check2.log; }
for vv in 1 2 3; do
  export FA_B200_V_VARIANT=$vv
  run f32 64 2 256 0 0 3
  run f32 32 2 256 0 0 3
done
unset FA_B200_V_VARIANT
run f32 64 2 256 1 0 5
run f32 64 3 1000 1 1.0 5
run f32 32 4 512 0 0 5
run f32 64 16 1024 0 0 20
run f32 64 16 8192 0 0 20
run f32 64 16 8192 0 1.0 20
run f32 64 16 8192 1 0 20
run f32 32 128 1024 0 0 20
echo "== test harness" >> gpurun_out/check2.log
timeout 120 $H/test >> gpurun_out/check2.log 2>&1 || echo "  (exit $?)" >> gpurun_out/check2.log
cat gpurun_out/probe2.log | grep PROBE
cat gpurun_out/check2.log
