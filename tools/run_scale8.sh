#!/bin/bash
# 8-GPU bench session (one gpurun --gpus 8 call): metric config, BASELINE config 3 and config 5; lines land in gpurun_out/r2_bench_*_n8.json
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2_bench_metric_n8.json 2> gpurun_out/n8_metric.err
timeout 300 $TR --master-port 29512 bench.py --gpus 8 --config c3 --no-cpu-baseline > gpurun_out/r2_bench_c3_n8.json 2> gpurun_out/n8_c3.err
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --config c5 --steps 3 --no-cpu-baseline > gpurun_out/r2_bench_c5_n8.json 2> gpurun_out/n8_c5.err
for f in metric_n8 c3_n8 c5_n8; do python - <<PY
import json
try:
    lines=[l for l in open("gpurun_out/r2_bench_$f.json") if l.startswith("{")]
    d=json.loads(lines[-1]); print("$f", d["value"], d["ms_per_step"], d.get("ba_ms_per_step"), d["e2e"]["value"], d.get("pose_system_reduction",{}).get("how"))
except Exception as e:
    print("$f FAILED", e)
PY
done
tail -3 gpurun_out/n8_metric.err gpurun_out/n8_c3.err gpurun_out/n8_c5.err | cut -c1-300
