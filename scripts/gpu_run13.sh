#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2l_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
timeout 600 python bench.py --config gen_infer --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench_gen.json 2> gpurun_out/r2l_bench_gen.err
tail -3 gpurun_out/r2l_tests.log
