#!/bin/bash
# round 2, call B (2 GPUs): GPU suite incl. the multi-GPU test, bench at N=1 and N=2 (both arms)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2b_topo.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=8 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -25 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
echo "bench n1 rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
echo "bench n2 rc=$?"
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2b_ref_n2.json 2> gpurun_out/r2b_ref_n2.err
echo "ref n2 rc=$?"
tail -c 600 gpurun_out/r2b_bench_n2.err
