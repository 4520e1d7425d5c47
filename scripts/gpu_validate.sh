#!/bin/bash
# Full GPU validation of the tree (what the driver runs at round end, plus every bench config): parity suite, smoke,
# the default bench line, the other BASELINE configs.  usage: gpurun --timeout 2400 -- 'bash scripts/gpu_validate.sh'
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s -p no:cacheprovider > gpurun_out/validate_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/validate_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/validate_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/validate_smoke.log
for cfg in train gen_infer infer512 d_step; do
  timeout 900 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/validate_bench_$cfg.json 2> gpurun_out/validate_bench_$cfg.err
done
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/validate_bench_reference.json 2> gpurun_out/validate_bench_reference.err
tail -3 gpurun_out/validate_tests.log; tail -2 gpurun_out/validate_smoke.log
