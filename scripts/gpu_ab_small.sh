#!/bin/bash
# A/B of the builds in flashattention.c_b200/variants on the small / narrow shapes only (cheap: no 8192-long inputs)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/ab_small.log
: > $L
for v in $(ls flashattention.c_b200/variants); do
  export LD_LIBRARY_PATH=$PWD/flashattention.c_b200/variants/$v
  echo "#### variant $v" >> $L
  run() { timeout 120 $H/fa_check "$@" >> $L 2>&1 || echo "  (exit $?)" >> $L; }
  run f32 32 3 1000 1 0 3
  run bf16 64 5 777 1 0 3
  run f32 32 128 1024 0 0 40 0
  run bf16 64 128 1024 0 0 40 0
  run bf16 64 64 4096 0 0 20 0
  run f32 32 64 4096 0 0 20 0
done
python - <<'PY' >> $L
import json, re, collections
rows = collections.defaultdict(dict)
v = None
for line in open("gpurun_out/ab_small.log"):
    m = re.match(r"#### variant (\S+)", line)
    if m: v = m.group(1); continue
    if line.startswith("{"):
        j = json.loads(line)
        rows[j["check"]].setdefault(v, []).append((j["ms_median"], j["ms_min"], j["err_tc_vs_fp64"]))
print("== summary: median / min ms per variant, max err")
for k, d in rows.items():
    print(k[:44].ljust(46), "  ".join(f"{vv}:{min(a for a, _, _ in x):.4f}/{min(b for _, b, _ in x):.4f} ({max(c for _, _, c in x):.0e})" for vv, x in sorted(d.items())))
PY
tail -n 8 $L | cut -c1-300
