#!/bin/bash
# round 2, call E (8 GPUs): e2e by chunk schedule with 8 ranks on one host; bench (ours) again for the ring's clock samples
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29561 scripts/e2e_multi.py > gpurun_out/r2e_e2e_multi_n8.log 2> gpurun_out/r2e_e2e_multi_n8.err
cat gpurun_out/r2e_e2e_multi_n8.log
timeout 600 $TR --nproc-per-node 2 --master-port 29562 scripts/e2e_multi.py > gpurun_out/r2e_e2e_multi_n2.log 2>> gpurun_out/r2e_e2e_multi_n8.err
cat gpurun_out/r2e_e2e_multi_n2.log
timeout 900 $TR --nproc-per-node 8 --master-port 29563 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n8.json 2> gpurun_out/r2e_bench_n8.err
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/r2e_bench_n8.json") if l.startswith("{")][-1])
print(json.dumps(j.get("c5_ring"), indent=0)[:1500])
PY
