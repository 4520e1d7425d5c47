#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -x -p no:cacheprovider > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_train.json 2> gpurun_out/r2f_bench_train.err
timeout 600 python bench.py --config gen_infer --steps 50 --warmup 5 > gpurun_out/r2f_bench_gen_infer.json 2> gpurun_out/r2f_bench_gen_infer.err
timeout 600 python bench.py --batch 64 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_b64.json 2> gpurun_out/r2f_bench_b64.err
timeout 300 python scripts/graph_timeline.py 32 gpurun_out/r2f_graph_timeline.md > gpurun_out/r2f_timeline.log 2>&1
bash scripts/ncu_full_r2.sh > gpurun_out/r2f_ncu_full.log 2>&1
tail -3 gpurun_out/r2f_tests.log
