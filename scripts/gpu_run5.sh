#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/graph_timeline.py 32 gpurun_out/r2e_graph_timeline.md > gpurun_out/r2e_timeline.log 2>&1
IRFD_SIDE_STREAM=0 timeout 300 python scripts/graph_timeline.py 32 gpurun_out/r2e_graph_timeline_noside.md > gpurun_out/r2e_timeline_noside.log 2>&1
IRFD_SIDE_STREAM=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench_noside.json 2> gpurun_out/r2e_bench_noside.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench_side.json 2> gpurun_out/r2e_bench_side.err
timeout 900 python scripts/bench_conv_shapes.py --md gpurun_out/r2e_conv_shapes.md > gpurun_out/r2e_conv_shapes.log 2>&1
tail -2 gpurun_out/r2e_timeline.log
