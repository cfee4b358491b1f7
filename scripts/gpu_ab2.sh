#!/bin/bash
# A/B of the variants in flashattention.c_b200/variants (timing sweep over the BASELINE shapes with the torch-free harness,
# each shape twice), then the full GPU test suite on the default build.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/ab2.log
: > $L
for rep in 1 2; do
for v in $(ls flashattention.c_b200/variants); do
  export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/$v
  echo "#### variant $v (pass $rep)" >> $L
  run() { timeout 120 $H/fa_check "$@" >> $L 2>&1 || echo "  (exit $?)" >> $L; }
  if [ $rep = 1 ]; then
    run f32 32 3 1000 1 0 3
    run bf16 64 5 777 1 0 3
  fi
  run f32 64 16 1024 0 0 30 0
  run f32 32 128 1024 0 0 30 0
  run bf16 64 128 1024 0 0 30 0
  run f32 64 16 8192 0 0 20 0
  run bf16 128 128 8192 0 0 10 0
  run bf16 128 128 8192 1 0 10 0
done
done
unset LD_LIBRARY_PATH
echo "== pytest -m gpu (default build)" >> $L
timeout 900 python -m pytest tests -q -m gpu -x >> $L 2>&1
echo "pytest exit $?" >> $L
python - <<'PY' >> $L
import json, re, collections
rows = collections.defaultdict(dict)
v = None
for line in open("gpurun_out/ab2.log"):
    m = re.match(r"#### variant (\S+)", line)
    if m: v = m.group(1); continue
    if line.startswith("{"):
        j = json.loads(line)
        rows[j["check"]].setdefault(v, []).append(j["ms_median"])
print("== summary: median ms per variant (min over passes)")
for k, d in rows.items():
    print(k[:44].ljust(46), "  ".join(f"{vv}:{min(x):.4f}" for vv, x in sorted(d.items())))
PY
tail -n 25 $L | cut -c1-300
