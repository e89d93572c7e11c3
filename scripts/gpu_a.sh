#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py -m gpu -q -k psd --timeout 300 > gpurun_out/pytest_psd.log 2>&1; echo "pytest psd rc=$?"; tail -2 gpurun_out/pytest_psd.log
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; echo "bench default rc=$?"; python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/bench_default.json") if x.startswith("{")][-1])
print({k:d[k] for k in ("value","ms_per_step","steps","gpu_launches")}, d["roofline"]["frac"], d["cpu_baseline"], d["e2e"]["value"], d["clocks"])
PY
tail -3 gpurun_out/bench_default.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | grep real; cat gpurun_out/bench_reference.json | cut -c1-600
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:k_batch_solve -c 1 -s 1 -o gpurun_out/prof_batch_r1b -f python scripts/config_runs.py c5 --nprob 296 --iters 6 --warmup 2 > gpurun_out/ncu_batch.log 2>&1; echo "ncu batch rc=$?"; tail -2 gpurun_out/ncu_batch.log | cut -c1-300
