#!/bin/bash
# round-2 GPU pass 2: lockstep encoders + stacked generator call; parity suite, the three bench configs
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_train.json 2> gpurun_out/r2b_bench_train.err
timeout 600 python bench.py --config gen_infer --steps 50 --warmup 5 > gpurun_out/r2b_bench_gen_infer.json 2> gpurun_out/r2b_bench_gen_infer.err
timeout 600 python bench.py --config infer512 --steps 20 --warmup 5 > gpurun_out/r2b_bench_infer512.json 2> gpurun_out/r2b_bench_infer512.err
tail -3 gpurun_out/r2b_tests.log
