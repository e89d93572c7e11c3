#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:k_cg_tail -c 2 -s 20 -o gpurun_out/prof_tail_r1 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tail.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tail.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1; echo "ncu list rc=$?"
