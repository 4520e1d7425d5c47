#!/bin/bash
# A/B on ONE box: bench.py under two settings of an environment variable, alternating, 2 rounds.
# usage: bash scripts/cycle_ab.sh VAR A B
mkdir -p gpurun_out
for round in 1 2; do
for v in $2 $3; do
env $1=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
    print('$1=$v', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], {k:(round(x['ms_per_step'],2)) for k,x in d['roofline'].get('hbm_families',{}).items()}, round(d['roofline']['ms_per_step'],2))
except Exception as e:
    print('$1=$v failed', e); print(open('gpurun_out/ab_$v.err').read()[-1500:])
PY
done
done
