#!/bin/bash
# A/B of variants (names X*: timing; T*: FA_TRACE builds -> timeline reports, incl. a slot-A-only run that shows the
# softmax phases without the other slot's interference)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/ab.log
: > $L
for v in $(ls flashattention.c_b200/variants); do
  export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/$v
  if [[ "$v" == T* ]]; then
    FA_B200_TRACE=gpurun_out/trace_${v}_c4.txt timeout 120 $H/fa_check bf16 128 128 8192 0 0 2 0 > /dev/null
    FA_B200_TRACE=gpurun_out/trace_${v}_c2.txt timeout 120 $H/fa_check f32 64 16 8192 0 0 2 0 > /dev/null
    FA_B200_TAIL_SPLIT=0 FA_B200_TRACE=gpurun_out/trace_${v}_c4_slotA.txt timeout 120 $H/fa_check bf16 128 2 8192 0 0 2 0 > /dev/null
    for f in c4 c2 c4_slotA; do python scripts/trace_report.py gpurun_out/trace_${v}_$f.txt > gpurun_out/trace_${v}_${f}_report.txt 2>&1; done
    continue
  fi
  echo "#### variant $v" >> $L
  run() { timeout 120 $H/fa_check "$@" >> $L 2>&1 || echo "  (exit $?)" >> $L; }
  run f32 64 3 1000 1 0 3
  run bf16 128 2 777 1 0 3
  run f32 64 16 1024 0 0 20 0
  run f32 64 16 8192 0 0 20 0
  run f32 32 128 1024 0 0 20 0
  run bf16 128 128 8192 0 0 10 0
  run bf16 128 128 8192 1 0 10 0
done
cut -c1-70,150-400 $L
head -30 gpurun_out/trace_T*_report.txt
