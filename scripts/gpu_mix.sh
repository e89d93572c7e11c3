#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/psd_probe.py > gpurun_out/psd_probe.jsonl 2> gpurun_out/psd_probe.err; echo "probe rc=$?"; cut -c1-220 gpurun_out/psd_probe.jsonl; tail -5 gpurun_out/psd_probe.err
timeout 600 python scripts/config_runs.py c4 --iters 30 > gpurun_out/c4.jsonl 2> gpurun_out/c4.err; echo "c4 rc=$?"; cat gpurun_out/c4.jsonl; tail -5 gpurun_out/c4.err
