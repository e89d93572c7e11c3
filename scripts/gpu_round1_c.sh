#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "GPU-vs-C" gpurun_out/pytest_gpu.log | head; tail -40 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"; cat gpurun_out/bench_c.json; tail -5 gpurun_out/bench_c.err
# ncu: launch list (shares of the step) and one full capture of the dominant kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_dual_matvec_tma -s 3 -c 2 -o gpurun_out/prof_k1_r1 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
