#!/bin/bash
# First GPU call of the next round: validate what was written after the GPU minutes ran out.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- bash scripts/gpu_next_round.sh
mkdir -p gpurun_out
{
  echo "== hybrid_check =="; timeout 600 python scripts/hybrid_check.py
  echo "== pytest -m gpu (golden fixtures first) =="; timeout 300 python -m pytest tests/test_golden.py -q -m gpu -x
  timeout 900 python -m pytest tests -q -m gpu -x
  echo "== C3 dense vs hybrid =="
  timeout 400 python scripts/config_runs.py c3 --iters 30
  timeout 400 python scripts/config_runs.py c3 --iters 30 --hybrid
} > gpurun_out/next_round.log 2>&1
tail -40 gpurun_out/next_round.log
