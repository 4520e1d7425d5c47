#!/bin/bash
# GPU cycle: generator/trainer tests, the default bench line, then launch list + graph timeline of the current build.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_trainer.py tests/test_gpu_train_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/cycle_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/cycle_tests.log
tail -3 gpurun_out/cycle_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/cycle_bench.json 2> gpurun_out/cycle_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/cycle_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline'].get('frac'), d.get('gpu_launches'))
PY
bash scripts/profile_final_r2.sh
