#!/bin/bash
# compute-sanitizer over the new kernels (small cases): memcheck + racecheck + synccheck
mkdir -p gpurun_out
SEL='test_batch_lockstep_strict_1e10[shape0-FISTA] or test_batch_results_independent or test_psd_large_batched_cones[129-5] or test_fused_cg_tail_matches_kernel_per_step_path[lasso-DR] or test_direct_affine_projection_is_exact[nnls] or test_linesearch_wrapper_lockstep[nnls-GAPA]'
for tool in memcheck racecheck synccheck; do
timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_batch.py tests/test_gpu_units.py tests/test_gpu_solvers.py -m gpu -q -x -k "$SEL" --timeout 1400 -p no:cacheprovider > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Race|hazard|Invalid|Barrier" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -12
done
