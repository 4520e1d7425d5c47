#!/bin/bash
# N=2 data-parallel: weak-scaling line with dp_check, and config 4's 2 x 128 shard
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --global-batch 256 > gpurun_out/r2g_bench_n2_g256.json 2> gpurun_out/r2g_bench_n2_g256.err
tail -c 600 gpurun_out/r2g_bench_n2.json
