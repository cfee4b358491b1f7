#!/bin/bash
# A/B of every build in flashattention.c_b200/variants against the shipped library: fa_check timings on a few shapes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
for v in shipped $(ls flashattention.c_b200/variants); do
  if [ $v = shipped ]; then unset LD_LIBRARY_PATH; else export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/$v; fi
  for args in "f32 64 16 8192 0 0" "f32 32 128 1024 0 0" "f32 64 16 1024 0 0" "bf16 64 64 4096 0 0" "bf16 128 128 8192 0 0"; do
    echo -n "$v | $args : "
    timeout 120 $H/fa_check $args 20 0 | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['ms_median'], j['ms_min'], 'err', j['err_tc_vs_fp64'])"
  done
done
