#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2o_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_bench2.json 2> gpurun_out/r2o_bench2.err
tail -3 gpurun_out/r2o_tests.log
