#!/bin/bash
# round 2, second pass: (1) is the synccheck / racecheck failure on the graph path the tool or the code?  Run the 40-line
# stand-alone WHILE-node probe (scripts/probes/cond_graph_probe.cu: no product code) under each tool.
# (2) the same tests with the graph path switched off (FOS_TEST_USE_GRAPHS=0): the same kernels, launched one by one.
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/cond_graph_probe scripts/probes/cond_graph_probe.cu 2>&1 | tail -3
echo "== probe plain"; timeout 120 /tmp/cond_graph_probe 2>&1 | tail -4
for tool in memcheck synccheck racecheck; do
echo "== probe under $tool"; timeout 300 compute-sanitizer --tool $tool /tmp/cond_graph_probe 2>&1 | tail -6
done
SEL='test_psd_large_batched_cones[129-5] or test_psd_projection_vs_lapack[True-130] or test_graph_path_is_bitwise_the_kernel_per_launch_path[socls-GAPA] or test_graph_path_is_bitwise_the_kernel_per_launch_path[nnls-FISTA] or test_graph_path_is_bitwise_the_kernel_per_launch_path[lasso-Dykstra] or test_hybrid_lockstep_1e10[GAPA-socls_wide] or test_lockstep_strict_1e10 and nnls and DR'
for tool in synccheck racecheck; do
FOS_TEST_USE_GRAPHS=0 timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_units.py tests/test_gpu_solvers.py tests/test_gpu_hybrid.py -m gpu -q -k "$SEL" --timeout 1100 -p no:cacheprovider > gpurun_out/r2_sanitizer_nograph_$tool.log 2>&1; echo "$tool (kernel-per-launch path) rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_nograph_$tool.log
grep -E "Race reported|Error:" gpurun_out/r2_sanitizer_nograph_$tool.log | sed -E 's/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -8
done
