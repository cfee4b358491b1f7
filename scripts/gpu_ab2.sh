#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
bash scripts/gpu_ab.sh > /dev/null
unset LD_LIBRARY_PATH
echo "== pytest -m gpu" > gpurun_out/pytest.log
timeout 900 python -m pytest tests -q -m gpu >> gpurun_out/pytest.log 2>&1
tail -n 15 gpurun_out/pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2>> gpurun_out/pytest.log
cat gpurun_out/bench_ours.json
