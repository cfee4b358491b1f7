#!/bin/bash
# ring-attention timing under different NCCL transport settings (2 GPUs)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/ring_env_${N}gpu.log
: > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { echo "== $*" >> $L; env "$@" timeout 300 $TR --master-port 29521 scripts/multi_gpu_check.py --reps 5 2>&1 | grep ring_timing >> $L; }
run FOO=default
run NCCL_MAX_NCHANNELS=2
run NCCL_MAX_NCHANNELS=4
run NCCL_P2P_USE_CUDA_MEMCPY=1
run NCCL_MAX_CTAS=4
cat $L
