#!/bin/bash
# 1 -> 8 GPU scaling of config 2 (bench.py, strong scaling, fused peer-memory exchange) and config 3 at 8 GPUs
mkdir -p gpurun_out
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_${N}gpu_p2p.json 2> gpurun_out/bench_${N}gpu_p2p.err; echo "bench$N rc=$?"; python - <<PY
import json
l=[x for x in open("gpurun_out/bench_${N}gpu_p2p.json") if x.startswith("{")]
d=json.loads(l[-1]); print(d["n_gpus"], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "k1 frac", round(d["roofline"]["frac"],3), "share", round(d["roofline"]["share_of_step"],3), "launches", d["gpu_launches"], d["clocks"]["sm_mhz"])
PY
grep -v "^\*\|OMP_NUM" gpurun_out/bench_${N}gpu_p2p.err | tail -3
done
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench1 rc=$?"; python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/bench_1gpu.json") if x.startswith("{")][-1]); print(d["n_gpus"], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "k1 frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 scripts/config_runs.py c3 --iters 30 > gpurun_out/c3_8gpu.jsonl 2> gpurun_out/c3_8gpu.err; echo "c3x8 rc=$?"; grep "^{" gpurun_out/c3_8gpu.jsonl; grep -v "^\*\|OMP_NUM" gpurun_out/c3_8gpu.err | tail -3
