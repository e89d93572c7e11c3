#!/bin/bash
# round 2, third pass.  (a) graph path under synccheck with a larger tracked-barrier table (the first pass printed
# "Detected overflow of tracked cuda::barrier structures" right before the launch failure); (b) graph path under racecheck, alone;
# (c) kernel-per-launch path (FOS_TEST_USE_GRAPHS=0) under synccheck and racecheck on tests that do not choose the path themselves.
mkdir -p gpurun_out
G='test_graph_path_is_bitwise_the_kernel_per_launch_path[lasso-DR]'
for nb in 4096 65536; do
timeout 600 compute-sanitizer --tool synccheck --num-cuda-barriers $nb --error-exitcode 7 python -m pytest tests/test_gpu_solvers.py -m gpu -q -k "$G" --timeout 500 -p no:cacheprovider > gpurun_out/r2_sanitizer_graph_synccheck_nb$nb.log 2>&1; echo "(a) synccheck graph path, --num-cuda-barriers $nb: rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Warning|launch failure" gpurun_out/r2_sanitizer_graph_synccheck_nb$nb.log | cut -c1-220 | sort | uniq -c | head -6
done
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_solvers.py -m gpu -q -k "$G" --timeout 800 -p no:cacheprovider > gpurun_out/r2_sanitizer_graph_racecheck.log 2>&1; echo "(b) racecheck graph path: rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|Segmentation|Race reported" gpurun_out/r2_sanitizer_graph_racecheck.log | cut -c1-220 | sort | uniq -c | head -6
SEL='test_lockstep_strict_1e10[lasso-DR] or test_lockstep_strict_1e10[socls-GAPA] or test_lockstep_strict_1e10[nnls-FISTA] or test_lockstep_strict_1e10[socls-Dykstra] or test_hybrid_lockstep_1e10[GAPA-socls_wide] or test_fused_cg_tail_matches_kernel_per_step_path[socls-FISTA] or test_psd_large_batched_cones[129-5]'
for tool in synccheck racecheck; do
FOS_TEST_USE_GRAPHS=0 timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_units.py tests/test_gpu_solvers.py tests/test_gpu_hybrid.py -m gpu -q -k "$SEL" --timeout 1100 -p no:cacheprovider > gpurun_out/r2_sanitizer_nograph_$tool.log 2>&1; echo "(c) $tool, kernel-per-launch path: rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_nograph_$tool.log
grep -E "Race reported|Error:" gpurun_out/r2_sanitizer_nograph_$tool.log | sed -E 's/\+0x[0-9a-f]+//' | cut -c1-220 | sort | uniq -c | sort -rn | head -8
done
