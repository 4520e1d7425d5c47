#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --config d_step --steps 10 --warmup 3 > gpurun_out/r2h_bench_d_step.json 2> gpurun_out/r2h_bench_d_step.err
timeout 600 python bench.py --config infer512 --steps 20 --warmup 5 > gpurun_out/r2h_bench_infer512.json 2> gpurun_out/r2h_bench_infer512.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err
tail -c 400 gpurun_out/r2h_bench_d_step.json; tail -c 300 gpurun_out/r2h_bench_d_step.err
