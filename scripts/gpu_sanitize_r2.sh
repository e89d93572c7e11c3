#!/bin/bash
# round 2: compute-sanitizer over the kernels that are new or rewritten this round (small cases):
#   PSD group kernel (three lane-group widths + mixed orders), the rewritten batch pass, the per-iteration
#   CUDA graph with its conditional WHILE node, the relaxed-input cone kernels, hybrid row storage.
mkdir -p gpurun_out
SEL='test_batch_lockstep_strict_1e10[shape0-FISTA] or test_batch_lockstep_strict_1e10[shape1-Dykstra] or test_batch_hybrid_storage_equals_dense_storage or test_psd_projection_vs_lapack[True-8] or test_psd_projection_vs_lapack[True-33] or test_psd_projection_vs_lapack[True-100] or test_psd_projection_vs_lapack[True-130] or test_psd_mixed_orders_in_one_cone_set or test_graph_path_is_bitwise_the_kernel_per_launch_path[socls-GAPA] or test_graph_path_is_bitwise_the_kernel_per_launch_path[nnls-FISTA] or test_graph_path_is_bitwise_the_kernel_per_launch_path[lasso-Dykstra] or test_hybrid_lockstep_1e10[GAPA-socls_wide]'
for tool in memcheck synccheck racecheck; do
timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_batch.py tests/test_gpu_units.py tests/test_gpu_solvers.py tests/test_gpu_hybrid.py -m gpu -q -k "$SEL" --timeout 1100 -p no:cacheprovider > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Race reported|hazard|Invalid|Barrier" gpurun_out/r2_sanitizer_$tool.log | sed -E 's/0x[0-9a-f]+/0x…/g; s/thread \([0-9,]+\)/thread (…)/g; s/block \([0-9,]+\)/block (…)/g' | sort | uniq -c | sort -rn | head -12
grep -E "at .*\+0x|in .*\.cu" gpurun_out/r2_sanitizer_$tool.log | sed -E 's/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -8
done
