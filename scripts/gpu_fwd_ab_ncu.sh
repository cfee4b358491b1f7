#!/bin/bash
# per-launch durations (ncu, gpu__time_duration) of the forward kernel for build variants: scripts/gpu_fwd_ab_ncu.sh "B H N d dtype causal" v1 v2 ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SHAPE="$1"; shift
for v in "$@"; do
  FA_B200_LIB=$PWD/flashattention.c_b200/variants/$v/libfa_b200.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fa_fwd_sm100 --csv --log-file gpurun_out/fwd_ab_ncu_$v.csv python scripts/fwd_one.py $SHAPE 4 > /dev/null 2>&1
  echo "== $v ($SHAPE): forward launch durations in us" >> gpurun_out/fwd_ab_ncu.log
  grep -v "^==" gpurun_out/fwd_ab_ncu_$v.csv | awk -F'","' 'NR>1 {gsub(/"/,"",$NF); printf "%s ", $NF/1000} END {print ""}' >> gpurun_out/fwd_ab_ncu.log
done
