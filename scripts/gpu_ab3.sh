#!/bin/bash
# A/B of the builds in flashattention.c_b200/variants: correctness spot checks (tcgen05 vs CUDA-core kernel vs fp64 rows),
# then timings on C2, a C4-shaped problem with fewer heads (37 x 32 items = 8 waves; cheap host-side input generation), its
# causal form, C3 and bf16 d=64.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/ab3.log
: > $L
for v in $(ls flashattention.c_b200/variants); do
  export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/$v
  echo "#### variant $v" >> $L
  run() { timeout 120 $H/fa_check "$@" >> $L 2>&1 || echo "  (exit $?)" >> $L; }
  run bf16 128 3 777 1 0 3
  run bf16 64 5 1000 0 0 3
  run f16 128 2 600 0 0 3
  run bf16 256 2 515 1 0 3
  run f32 64 16 8192 0 0 20 0
  run bf16 128 37 8192 0 0 12 0
  run bf16 128 37 8192 1 0 12 0
  run f32 32 128 1024 0 0 30 0
  run bf16 64 128 1024 0 0 30 0
done
python - <<'PY' >> $L
import json, re, collections
rows = collections.defaultdict(dict)
v = None
for line in open("gpurun_out/ab3.log"):
    m = re.match(r"#### variant (\S+)", line)
    if m: v = m.group(1); continue
    if line.startswith("{"):
        j = json.loads(line)
        rows[j["check"]][v] = (j["ms_median"], j["err_tc_vs_fp64"], j["tc_vs_simt"])
print("== summary: median ms (err vs fp64 rows / vs CUDA-core kernel)")
for k, d in rows.items():
    print(k[:40].ljust(42), "  ".join(f"{vv}:{a:.4f} ({b:.0e}/{c:.0e})" for vv, (a, b, c) in sorted(d.items())))
PY
grep -c "exit\|failed" $L
tail -n 11 $L | cut -c1-330
