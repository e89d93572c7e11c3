#!/bin/bash
# first GPU contact: build check, probe (incl. compute-sanitizer on the small shapes), tests, small bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 300 python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?"; tail -25 gpurun_out/probe.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/gpu_probe.py --small > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -15 gpurun_out/sanitizer.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench rc=$?"; cat gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
