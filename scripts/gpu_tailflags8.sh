#!/bin/bash
mkdir -p gpurun_out
N=8
for mode in 1 0; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$mode bench.py --gpus $N --steps 30 --warmup 3 --tail-flags $mode > gpurun_out/bench_8gpu_flags$mode.json 2> gpurun_out/bench_8gpu_flags$mode.err; echo "mode $mode rc=$?"; python - <<PY
import json
l=[x for x in open("gpurun_out/bench_8gpu_flags$mode.json") if x.startswith("{")]
d=json.loads(l[-1]); print("tail_flags", $mode, "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "k1 ms", round(d["roofline"]["avg_launch_ms"],4), "tail_us", round(d["cg_tail_avg_launch_us"],1), "passes", d["passes_over_A_per_step"])
PY
done
