#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_units.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu2.log
bash scripts/gpu_bench_n.sh 1
