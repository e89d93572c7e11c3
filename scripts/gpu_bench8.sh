#!/bin/bash
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_${N}gpu_p2p.json 2> gpurun_out/bench_${N}gpu_p2p.err; echo "bench$N rc=$?"; python - <<PY
import json
l=[x for x in open("gpurun_out/bench_${N}gpu_p2p.json") if x.startswith("{")]
d=json.loads(l[-1]); print(d["n_gpus"], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "k1 frac", round(d["roofline"]["frac"],3), "share", round(d["roofline"]["share_of_step"],3), "launches", d["gpu_launches"], d["clocks"]["sm_mhz"])
PY
grep -v "^\*\|OMP_NUM" gpurun_out/bench_${N}gpu_p2p.err | tail -3
