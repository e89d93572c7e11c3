#!/bin/bash
# large-PSD kernel: the unit-test sequence replayed several times, unit tests, timing against cuSOLVER, config 4
mkdir -p gpurun_out
timeout 150 python scripts/probes/psd_large_debug2.py 5 2>&1 | grep -c "BAD" | sed 's/^/replay: BAD lines = /'
timeout 150 python scripts/probes/psd_large_debug2.py 1 2>&1 | tail -4
timeout 200 python -m pytest tests/test_gpu_units.py tests/test_golden.py -m gpu -q -x -k "psd" 2>&1 | tail -3
timeout 200 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_scale.py -m gpu -q -x -k "sdp or SDP" 2>&1 | tail -3
timeout 200 python scripts/psd_probe.py --large > gpurun_out/psd_probe_large.jsonl 2> gpurun_out/psd_probe_large.err; tail -3 gpurun_out/psd_probe_large.err
python - <<'PY'
import json
for l in open("gpurun_out/psd_probe_large.jsonl"):
    d = json.loads(l)
    print(d["d"], d["ncones"], "fos %.3f ms" % d["fos_cold_ms"], "sweeps", d["sweeps"], "lib %.3f" % d["cusolver_eigh_ms"], "speedup %.2f" % d["speedup_vs_best_library"], "diff %.1e" % d["max_rel_diff_vs_lib"])
PY
timeout 200 python scripts/config_runs.py c4 --iters 200 --warmup 30 2>&1 | tail -1 | tee gpurun_out/c4_new.jsonl | cut -c1-500
FOS_PSD_PROF=1 timeout 100 python scripts/probes/psd_large_debug.py 512:1 1024:1 256:1 2>&1 | grep -E "psd_large" | awk 'NR%8==1'
