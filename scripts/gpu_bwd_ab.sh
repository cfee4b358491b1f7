#!/bin/bash
# A/B of backward build variants (flashattention.c_b200/variants/<name>/libfa_b200.so): scripts/gpu_bwd_ab.sh name1 name2 ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/bwd_ab.log
for rep in 1 2; do
for v in "$@"; do
  echo "== variant $v (pass $rep)" >> gpurun_out/bwd_ab.log
  FA_B200_LIB=$PWD/flashattention.c_b200/variants/$v/libfa_b200.so BWD_NO_TORCH=1 timeout 200 python scripts/bwd_timing.py 5 >> gpurun_out/bwd_ab.log 2>&1
done
done
python - <<'PY'
import json
cur = None
for l in open("gpurun_out/bwd_ab.log"):
    if l.startswith("=="):
        cur = l.strip()
    elif l.startswith("{"):
        j = json.loads(l)
        print(f'{cur:28s} {j["config"]:36s} bwd {j["bwd_ms"]:8.4f} ms  {j["bwd_tflops_algorithmic_5gemm"]:7.1f} TF alg')
PY
