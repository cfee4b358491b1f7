#!/bin/bash
# bench.py (own arm) under torchrun at N = $1 (default 2) and a summary of its multi-GPU sections
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -c 300 gpurun_out/bench_n$N.err
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
print("value", j["value"], "e2e", j["e2e"]["ms_per_step"], "floor", j["e2e"]["host_copy_floor_ms"])
for k in ("c4_sharded", "c3_sharded", "c5_ring"):
    x = j.get(k, {})
    print(k, json.dumps({kk: vv for kk, vv in x.items() if kk not in ("workload", "timing", "parity", "transport")})[:700], "parity", x.get("parity", {}).get("ok"))
PY
