#!/bin/bash
# usage: scratch/run_tests_bench.sh  -> GPU tests summary + short bench summary
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/b.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b.json").read())
print("it/s", round(d["value"],1), "ms", round(d["ms_per_step"],4), "omega iters", d.get("omega_iters_last_sweep"))
r=d["roofline"]; print("frac", round(r["frac"],3), "fp64", r["fp64"])
for k,v in r["kernel_ms"].items(): print(k, [round(x*1000,1) for x in v])
PY
