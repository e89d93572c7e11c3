#!/bin/bash
# final 1-GPU validation of round 2: full GPU test suite, smoke, default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_gputest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_gputest_final.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/r2_bench_final.json") if x.startswith("{")][-1])
print("value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "frac", round(d["roofline"]["frac"],3), "tte", round(d["time_to_eps"]["seconds"],2), d["time_to_eps"]["status"], "parity pass", d["parity"]["pass"], "launches", d["gpu_launches"], "cpu", round(d["cpu_baseline"]["value"],3))
PY
