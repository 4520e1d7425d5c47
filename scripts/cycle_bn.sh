#!/bin/bash
# Quick GPU cycle for BatchNorm changes: the BN / encoder tests, then the default bench line.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_encoder_group.py tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_train_parity.py tests/test_gpu_trainer.py -m gpu -q -x -p no:cacheprovider > gpurun_out/cycle_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/cycle_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/cycle_bench.json 2> gpurun_out/cycle_bench.err
tail -4 gpurun_out/cycle_tests.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/cycle_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline'].get('frac'))
print({k:(round(v['ms_per_step'],2) if isinstance(v,dict) and 'ms_per_step' in v else v) for k,v in d['roofline'].get('hbm_families',{}).items()})
PY
