#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_encoder_group.py tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_inference.py -m gpu -q -x -p no:cacheprovider > gpurun_out/cycle_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/cycle_tests.log
tail -5 gpurun_out/cycle_tests.log
grep -q "rc=0" gpurun_out/cycle_tests.log || exit 1
bash scripts/cycle_ab.sh IRFD_GEMM_CLUSTER 0 1
