#!/bin/bash
mkdir -p gpurun_out
for it in 60 100 60 30; do
timeout 300 python scripts/config_runs.py c5 --nprob 1024 --iters $it 2> gpurun_out/c5s.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print($it, d['algorithm'], round(d['problem_iterations_per_s']), 'cg/step', round(d['cg_iterations_per_step'],2), 'p', d['check_p_median'], 'ms', round(d['ms_total'],1))
"
done
timeout 300 python scripts/config_runs.py c5 --nprob 1024 --iters 60 --dense-batch 2> gpurun_out/c5s.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('dense', d['algorithm'], round(d['problem_iterations_per_s']), 'cg/step', round(d['cg_iterations_per_step'],2), 'p', d['check_p_median'], 'ms', round(d['ms_total'],1))
"
