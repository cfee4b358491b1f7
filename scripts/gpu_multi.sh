#!/bin/bash
# multi-GPU session: parity of B x H sharding + ring attention over NCCL, ring timing, bench.py scaling line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/multi_${N}gpu.log
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multi_gpu_check world=$N" > $L
timeout 600 $TR --master-port 29511 scripts/multi_gpu_check.py >> $L 2>&1
echo "== bench --gpus $N" >> $L
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2>> $L
cat gpurun_out/bench_${N}gpu.json >> $L
echo "== bench --gpus $N --workload C4" >> $L
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --workload C4 > gpurun_out/bench_${N}gpu_c4.json 2>> $L
cat gpurun_out/bench_${N}gpu_c4.json >> $L
grep -v "^W\|^\[W\|Warning\|warn" $L | tail -n 40
