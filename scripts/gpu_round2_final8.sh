#!/bin/bash
# final 8-GPU check of round 2 (the driver's own launch line): config 2 through bench.py at N = 8, config 5 (8192 problems) split over 8 GPUs
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err; echo "bench8 rc=$?"
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2_bench_n8_final.json") if x.startswith("{")]
d=json.loads(l[-1]); print(d["n_gpus"], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "k1 frac", round(d["roofline"]["frac"],3), "tail us", round(d.get("cg_tail_avg_launch_us",0),1), "tte", d.get("time_to_eps",{}).get("seconds"), d.get("time_to_eps",{}).get("status"), "parity", json.dumps(d.get("multi_gpu_parity",{}))[:400])
PY
grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_n8_final.err | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 scripts/config_runs.py c5 --nprob 8192 --iters 100 > gpurun_out/r2_c5_8gpu_final.jsonl 2> gpurun_out/r2_c5_8gpu_final.err; echo "c5x8 rc=$?"
python - <<'PY'
import json
for x in open("gpurun_out/r2_c5_8gpu_final.jsonl"):
    if x.startswith("{"):
        d=json.loads(x); print(d["algorithm"], d["n_gpus"], "problem-iterations/s", round(d["problem_iterations_per_s"]), "frac/gpu", round(d["frac_per_gpu"],3))
PY
grep -v "^\*\|OMP_NUM" gpurun_out/r2_c5_8gpu_final.err | tail -3
