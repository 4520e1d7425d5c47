#!/bin/bash
# GPU cycle: folded BN reduce — tests, then A/B bench (IRFD_BN_FOLD=1 / 0).
mkdir -p gpurun_out
python -m pytest tests/test_gpu_encoder_group.py -m gpu -q -s -p no:cacheprovider > gpurun_out/cycle_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/cycle_tests.log
grep -a "parity\|passed\|failed\|Error\|rc=" gpurun_out/cycle_tests.log | tail -12
for f in 1 0; do
IRFD_BN3_FOLD=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/cycle_bench_fold$f.json 2> gpurun_out/cycle_bench_fold$f.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/cycle_bench_fold$f.json').read().strip().splitlines()[-1])
    print('fold=$f', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline'].get('frac'))
    print({k:(round(v['ms_per_step'],2)) for k,v in d['roofline'].get('hbm_families',{}).items()}, {k: round(v,2) if isinstance(v,float) else v for k,v in d['roofline'].items() if k in ('ms_per_step',)})
except Exception as e:
    print('fold=$f failed', e); print(open('gpurun_out/cycle_bench_fold$f.err').read()[-1500:])
PY
done
