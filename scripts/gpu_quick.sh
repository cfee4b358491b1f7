#!/bin/bash
# quick A/B: fa_check timings on the BASELINE shapes (+ a few neighbours), then the GPU test suite
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/quick.log
: > $L
for args in "f32 64 16 8192 0 0" "f32 64 16 8192 1 0" "f32 32 128 1024 0 0" "f32 64 16 1024 0 0" "bf16 64 128 1024 0 0" "bf16 64 64 4096 0 0" "f32 32 64 4096 0 0" "bf16 128 128 8192 0 0" "f32 64 72 4096 1 0.125"; do
  timeout 120 $H/fa_check $args 20 1 >> $L 2>&1
done
python - <<'PY'
import json
for l in open("gpurun_out/quick.log"):
    if l.startswith("{"):
        j = json.loads(l)
        print(f'{j["check"]:48s} ms_med {j["ms_median"]:.4f} min {j["ms_min"]:.4f}  TF {j["tflops_median"]:7.1f}  err {j["err_tc_vs_fp64"]:.2e} tc_vs_simt {j["tc_vs_simt"]:.2e} lse {j["lse_err"]:.2e}')
    elif "error" in l.lower():
        print(l.strip())
PY
if [ "$1" != "notest" ]; then
  timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -6
fi
