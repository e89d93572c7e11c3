#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi2.log 2>&1; echo "multi2 rc=$?"; tail -15 gpurun_out/multi2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; cat gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench1 rc=$?"; cat gpurun_out/bench_1gpu.json; tail -3 gpurun_out/bench_1gpu.err
