#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py -m gpu -q -k "psd" --timeout 300 > gpurun_out/pytest_psd.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_psd.log
timeout 600 python scripts/psd_probe.py > gpurun_out/psd_probe.jsonl 2> gpurun_out/psd_probe.err; echo "probe rc=$?"; cut -c1-200 gpurun_out/psd_probe.jsonl; tail -3 gpurun_out/psd_probe.err
