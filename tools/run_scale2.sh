#!/bin/bash
# 2-GPU bench session (one gpurun --gpus 2 call): metric config (fused peer-to-peer reduction) and BASELINE config 3 (NCCL all-reduce)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2_bench_metric_n2.json 2> gpurun_out/n2_metric.err
timeout 300 $TR --master-port 29522 bench.py --gpus 2 --config c3 --no-cpu-baseline > gpurun_out/r2_bench_c3_n2.json 2> gpurun_out/n2_c3.err
for f in metric_n2 c3_n2; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_$f.json") if l.startswith("{")][-1]); print("$f", d["value"], d["ms_per_step"], d.get("ba_ms_per_step"), d["e2e"]["value"], d.get("pose_system_reduction",{}).get("how"))
except Exception as e:
    print("$f FAILED", e)
PY
done
tail -n 3 gpurun_out/n2_metric.err | cut -c1-300
