#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py -m gpu -q --timeout 300 > gpurun_out/pytest_units.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_units.log
