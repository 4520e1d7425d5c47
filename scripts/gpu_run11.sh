#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_encoder_group.py tests/test_gpu_kernels.py tests/test_gpu_trainer.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2k_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_tests.log
IRFD_ZIGZAG=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench_z0.json 2> gpurun_out/r2k_bench_z0.err
IRFD_ZIGZAG=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench_z1.json 2> gpurun_out/r2k_bench_z1.err
tail -3 gpurun_out/r2k_tests.log
