#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2i_smoke.log
timeout 900 python bench.py --config d_step --steps 10 --warmup 3 > gpurun_out/r2i_bench_d_step.json 2> gpurun_out/r2i_bench_d_step.err
tail -3 gpurun_out/r2i_smoke.log; tail -c 300 gpurun_out/r2i_bench_d_step.err
