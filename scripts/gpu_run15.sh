#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/exp_epilogue.py > gpurun_out/r2n_exp_epilogue.log 2>&1
cat gpurun_out/r2n_exp_epilogue.log
