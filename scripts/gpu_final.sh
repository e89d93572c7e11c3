#!/bin/bash
# round-end evidence: launch list of the bench command (shares), smoke, default bench
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1; echo "ncu list rc=$?"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/bench_default.json") if x.startswith("{")][-1])
print({k:d[k] for k in ("value","ms_per_step","steps","gpu_launches")}, "frac", round(d["roofline"]["frac"],3), "share", round(d["roofline"]["share_of_step"],3), "cpu", d["cpu_baseline"]["value"], "e2e", round(d["e2e"]["value"],2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "tail_us", d.get("cg_tail_avg_launch_us"))
PY
