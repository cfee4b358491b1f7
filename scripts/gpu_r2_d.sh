#!/bin/bash
# round 2, call D (8 GPUs): the driver's SCALE commands at N=8 (both arms) and N=4 (ours), topology
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1
lscpu | grep -i "numa\|model name\|^CPU(s)" >> gpurun_out/r2d_topo.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n8.json 2> gpurun_out/r2d_bench_n8.err
echo "bench n8 rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29552 bench.py --impl reference --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2d_ref_n8.json 2> gpurun_out/r2d_ref_n8.err
echo "ref n8 rc=$?"
timeout 900 $TR --nproc-per-node 4 --master-port 29553 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n4.json 2> gpurun_out/r2d_bench_n4.err
echo "bench n4 rc=$?"
FA_BENCH_NUMA=spread timeout 600 $TR --nproc-per-node 8 --master-port 29554 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/r2d_bench_n8_spread.json 2> gpurun_out/r2d_bench_n8_spread.err
echo "bench n8 spread rc=$?"
tail -c 400 gpurun_out/r2d_bench_n8.err
python - <<'PY'
import json
for f in ("r2d_bench_n8", "r2d_bench_n4", "r2d_ref_n8", "r2d_bench_n8_spread"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
    except Exception as e:
        print(f, "NO LINE", e); continue
    print(f, "value", j.get("value"), "ms", j.get("ms_per_step"), "e2e", j.get("e2e", {}).get("ms_per_step"), "floor", j.get("e2e", {}).get("host_copy_floor_ms"), j.get("e2e", {}).get("host_binding"))
    for k in ("c4_sharded", "c3_sharded", "c5_ring"):
        if k in j:
            x = j[k]
            print("   ", k, {kk: x.get(kk) for kk in ("ms", "tflops_total", "efficiency", "speedup_vs_one_gpu", "overlap", "ms_nccl_transport", "frac_sustained_peak", "error")}, "parity", x.get("parity", {}).get("ok"))
PY
