#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py tests/test_gpu_solvers.py -m gpu -q -k "psd or sdp" --timeout 600 > gpurun_out/pytest_psd.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_psd.log
timeout 600 python scripts/config_runs.py c4 --iters 30 > gpurun_out/c4.jsonl 2> gpurun_out/c4.err; echo "c4 rc=$?"; cat gpurun_out/c4.jsonl; tail -3 gpurun_out/c4.err
timeout 600 python scripts/config_runs.py c4 --iters 200 --warmup 100 > gpurun_out/c4_late.jsonl 2> gpurun_out/c4.err; echo "c4 late rc=$?"; cat gpurun_out/c4_late.jsonl; tail -3 gpurun_out/c4.err
