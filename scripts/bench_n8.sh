#!/bin/bash
# N=8 data-parallel weak-scaling line (config 4: global 256 = 8 x 32) with dp_check
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
IRFD_BENCH_WATCHDOG_S=500 timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err
tail -c 700 gpurun_out/r2j_bench_n8.json; tail -c 500 gpurun_out/r2j_bench_n8.err
# N=1 on the same box (same power/clock conditions) for the scaling ratio
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench_n1_samebox.json 2> gpurun_out/r2j_bench_n1_samebox.err
tail -c 300 gpurun_out/r2j_bench_n1_samebox.json
