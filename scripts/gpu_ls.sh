#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solvers.py -m gpu -q -k "linesearch or direct or feasibility" --timeout 300 > gpurun_out/pytest_ls.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_ls.log
