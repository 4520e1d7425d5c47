#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
timeout 300 python scripts/faithful_probe.py > gpurun_out/r2c_faithful_probe.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_train.json 2> gpurun_out/r2c_bench_train.err
tail -3 gpurun_out/r2c_tests.log
