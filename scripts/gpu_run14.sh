#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_encoder_group.py tests/test_gpu_fullsize.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2m_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
timeout 600 python bench.py --config infer512 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench_infer512.json 2> gpurun_out/r2m_bench_infer512.err
timeout 600 python scripts/bench_conv_shapes.py --only "1x1" --md gpurun_out/r2m_conv_shapes_1x1.md > /dev/null 2>&1
tail -3 gpurun_out/r2m_tests.log
