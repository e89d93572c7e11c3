#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:k_batch_solve -c 1 -s 1 -o gpurun_out/prof_batch_r1 -f python scripts/config_runs.py c5 --nprob 148 --iters 6 --warmup 2 > gpurun_out/ncu_batch.log 2>&1; echo "ncu batch rc=$?"; tail -5 gpurun_out/ncu_batch.log
timeout 900 $NCU -k regex:k5_psd_hestenes -c 1 -s 1 -o gpurun_out/prof_psd_r1 -f python scripts/psd_probe.py > gpurun_out/ncu_psd.log 2>&1; echo "ncu psd rc=$?"; tail -5 gpurun_out/ncu_psd.log
ls -la gpurun_out/*.ncu-rep
