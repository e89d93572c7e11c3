#!/bin/bash
N=$1; shift
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi$N.log 2>&1; echo "multi$N rc=$?"; grep -c " ok" gpurun_out/multi$N.log; grep "FAIL\|differs\|Error" gpurun_out/multi$N.log | head
bash scripts/gpu_bench_n.sh "$@"
