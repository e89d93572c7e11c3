#!/bin/bash
# round 2, last pass: the batch kernel after the grouped consumer barrier (8 partial-sum buffers) and the evict-first tile loads
mkdir -p gpurun_out
SEL='test_batch_lockstep_strict_1e10[shape0-FISTA] or test_batch_lockstep_strict_1e10[shape1-Dykstra] or test_batch_hybrid_storage_equals_dense_storage or test_batch_results_independent'
for tool in memcheck synccheck racecheck; do
timeout 500 compute-sanitizer --tool $tool --num-cuda-barriers 4096 --error-exitcode 7 python -m pytest tests/test_gpu_batch.py -m gpu -q -k "$SEL" --timeout 450 -p no:cacheprovider > gpurun_out/r2_sanitizer_batch_$tool.log 2>&1; echo "$tool rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_batch_$tool.log
grep -E "Race reported|Error:|and (Write|Read) access" gpurun_out/r2_sanitizer_batch_$tool.log | sed -E 's/\+0x[0-9a-f]+//; s/\[[0-9]+ hazards\]//' | cut -c1-220 | sort | uniq -c | sort -rn | head -10
done
