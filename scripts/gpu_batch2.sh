#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_batch.py -m gpu -q --timeout 300 > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_batch.log
timeout 900 python scripts/config_runs.py c5 --nprob 1024 --iters 100 > gpurun_out/c5.jsonl 2> gpurun_out/c5.err; echo "c5 rc=$?"; cat gpurun_out/c5.jsonl; tail -5 gpurun_out/c5.err
