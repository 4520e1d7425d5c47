#!/bin/bash
# N=4 data-parallel weak-scaling line (32 pairs per GPU) with dp_check
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555"
IRFD_BENCH_WATCHDOG_S=400 timeout 500 $TR bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2k_bench_n4.json 2> gpurun_out/r2k_bench_n4.err
tail -c 300 gpurun_out/r2k_bench_n4.json
