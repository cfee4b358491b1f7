#!/bin/bash
# round 2, call F (1 GPU): final-state check — GPU suite, smoke, both bench arms with the driver's default flags, reference table
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=5 > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -12 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/r2f_ref_n1.json 2> gpurun_out/r2f_ref_n1.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"
timeout 900 python scripts/ref_table.py > gpurun_out/r2f_ref_table.jsonl 2> gpurun_out/r2f_ref_table.err; echo "ref_table rc=$?"
tail -2 gpurun_out/r2f_ref_table.jsonl | cut -c1-600
python - <<'PY'
import json
for f in ("r2f_bench_n1", "r2f_ref_n1"):
    j = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
    print(f, "value", j.get("value"), "ms", j.get("ms_per_step"), "steps", j.get("steps"), "e2e", j.get("e2e", {}).get("ms_per_step"), "clocks", j.get("clocks"))
    if "other_configs" in j: print("   other:", json.dumps(j["other_configs"])[:900])
PY
