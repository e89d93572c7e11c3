#!/bin/bash
mkdir -p gpurun_out
for tb in 148 96 64 40 24; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --tail-blocks $tb > gpurun_out/bench_tb$tb.json 2> gpurun_out/bench_tb.err; python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/bench_tb$tb.json") if x.startswith("{")][-1]); print("tail_blocks", $tb, "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "k1 ms", round(d["roofline"]["avg_launch_ms"],4), "tail_us", round(d["cg_tail_avg_launch_us"],2), d["clocks"]["sm_mhz"])
PY
done
