#!/bin/bash
# single-GPU closing session: full GPU test suite, the default bench line, the reference arm, and the other BASELINE configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench_metric.json 2> gpurun_out/r2_bench_metric.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference_metric.json 2> gpurun_out/r2_bench_reference_metric.err
for c in c2 c3 c4; do timeout 600 python bench.py --config $c --no-cpu-baseline > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; done
timeout 600 python bench.py --config c5_rank --steps 3 --no-cpu-baseline > gpurun_out/r2_bench_c5_rank.json 2> gpurun_out/r2_bench_c5_rank.err
for f in metric reference_metric c2 c3 c4 c5_rank; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_$f.json") if l.startswith("{")][-1])
    print("$f", d.get("impl"), round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "ba", d.get("ba_ms_per_step"), "frac", d.get("roofline",{}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$f FAILED", e)
PY
done
