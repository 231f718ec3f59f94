#!/bin/bash
# usage: scratch/variants.sh tag "ENV1=.. ENV2=.." ...   -> one short bench line per environment variant
tag=$1; shift
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --no-calibration"
i=0
for v in "$@"; do
  env $v $B 2>/dev/null | tail -1 > gpurun_out/${tag}_$i.json
  python - "$v" gpurun_out/${tag}_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print('%-60s it/s %.0f ms %.4f median %.4f settled %.4f warm %s' % (sys.argv[1], d['value'], d['ms_per_step'], d.get('ms_per_step_median',0), d.get('ms_per_step_settled',0), d['roofline'].get('sweep_kernel',{}).get('ms_warm_l2')))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
  i=$((i+1))
done
