#!/bin/bash
# A/B: rotating S buffers (shipped lib) vs -DFA_OPT_ROT_S=0 (variants/R0) over sequence lengths: per-step vs per-item cost
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
for v in rot R0; do
  if [ $v = R0 ]; then export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/R0; fi
  for args in "f32 64 148 256 0 0" "f32 64 148 512 0 0" "f32 64 148 1024 0 0" "f32 64 148 2048 0 0" "f32 64 74 4096 0 0" "f32 64 16 8192 0 0" "f32 32 128 1024 0 0" "f32 32 148 2048 0 0"; do
    echo -n "$v $args : "
    timeout 120 $H/fa_check $args 20 0 | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['ms_median'], j['ms_min'])"
  done
done
