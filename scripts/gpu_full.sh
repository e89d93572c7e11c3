#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench1 rc=$?"; python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/bench_1gpu.json") if x.startswith("{")][-1]); print(d["n_gpus"], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "k1 frac", round(d["roofline"]["frac"],3), "share", round(d["roofline"]["share_of_step"],3), "whole", round(d["roofline"]["whole_iteration_frac"],3), "launches", d["gpu_launches"])
PY
tail -3 gpurun_out/bench_1gpu.err
