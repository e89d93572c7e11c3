#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi$N.log 2>&1; echo "multi$N rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/multi$N.log | tail -20
for ex in nccl p2p; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 3 --exchange $ex > gpurun_out/bench_${N}gpu_$ex.json 2> gpurun_out/bench_${N}gpu_$ex.err; echo "bench$N $ex rc=$?"; python - <<PY
import json
l=[x for x in open("gpurun_out/bench_${N}gpu_$ex.json") if x.startswith("{")]
d=json.loads(l[-1]); print("$ex", d["n_gpus"], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "k1 frac", round(d["roofline"]["frac"],3), "share", round(d["roofline"]["share_of_step"],3), "launches", d["gpu_launches"])
PY
tail -3 gpurun_out/bench_${N}gpu_$ex.err
done
