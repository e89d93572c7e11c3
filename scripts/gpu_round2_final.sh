#!/bin/bash
# final 1-GPU validation of round 2: full GPU test suite, smoke, PSD timing table, sanitizer on the rewritten large-PSD kernel, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_gputest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_gputest_final.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python scripts/psd_probe.py > gpurun_out/r2_psd_probe.jsonl 2>/dev/null; timeout 200 python scripts/psd_probe.py --large > gpurun_out/r2_psd_probe_large.jsonl 2>/dev/null
timeout 200 python scripts/config_runs.py c4 --iters 200 --warmup 30 2>/dev/null | tail -1 > gpurun_out/r2_c4_sdp512.jsonl
timeout 200 python scripts/config_runs.py c4 --iters 30 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/r2_c4_sdp512.jsonl
FOS_PSD_PROF=1 timeout 100 python scripts/probes/psd_large_debug.py 256:1 512:1 768:1 1024:1 2>&1 | grep -E "psd_large" | awk 'NR%8==1' > gpurun_out/r2_psd_large_cycles.txt
SEL='test_psd_large_batched_cones[129-5] or test_psd_projection_vs_lapack[True-130] or test_psd_projection_vs_lapack[False-200]'
for tool in memcheck synccheck racecheck; do
timeout 600 compute-sanitizer --tool $tool --num-cuda-barriers 4096 --error-exitcode 7 python -m pytest tests/test_gpu_units.py -m gpu -q -k "$SEL" --timeout 550 -p no:cacheprovider > gpurun_out/r2_sanitizer_psdlarge_$tool.log 2>&1; echo "$tool rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_psdlarge_$tool.log
grep -E "Race reported|Error:" gpurun_out/r2_sanitizer_psdlarge_$tool.log | sed -E 's/\+0x[0-9a-f]+//' | cut -c1-200 | sort | uniq -c | sort -rn | head -6
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 1500 gpurun_out/r2_bench_final.json
