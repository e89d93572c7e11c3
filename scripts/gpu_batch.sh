#!/bin/bash
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_batch.py -m gpu -q -x -k "lockstep and FISTA and 40" --timeout 280 > gpurun_out/batch_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -15 gpurun_out/batch_sanitizer.log
timeout 1200 python -m pytest tests/test_gpu_batch.py -m gpu -q --timeout 300 > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_batch.log
