#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_trainer.py tests/test_gpu_train_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/cycle_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/cycle_tests.log
tail -3 gpurun_out/cycle_tests.log
bash scripts/cycle_ab.sh IRFD_PREPACK 0 1
