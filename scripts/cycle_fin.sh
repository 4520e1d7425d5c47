#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_encoder_group.py tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_trainer.py -m gpu -q -x -p no:cacheprovider > gpurun_out/cycle_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/cycle_tests.log
tail -3 gpurun_out/cycle_tests.log
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/cycle_bench.json 2> gpurun_out/cycle_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/cycle_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],2) for k,v in d['roofline']['hbm_families'].items()})
PY
done
timeout 300 python scripts/graph_timeline.py 32 gpurun_out/fin_graph_timeline.md > gpurun_out/fin_timeline.log 2>&1
grep "wall\|finalize\|bn_" gpurun_out/fin_graph_timeline.md
