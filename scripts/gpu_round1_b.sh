#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?"; tail -22 gpurun_out/probe.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "worst one-step|GPU-vs-C" gpurun_out/pytest_gpu.log | head -40; tail -60 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
