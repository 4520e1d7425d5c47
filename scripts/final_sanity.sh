#!/bin/bash
# Last check of a build: the whole GPU suite, the smoke entry point, the default bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -3 gpurun_out/final_tests.log; tail -2 gpurun_out/final_smoke.log; tail -c 400 gpurun_out/final_bench.json
