#!/bin/bash
# final round-2 profile: launch list of one static-eager step (time + DRAM bytes), graph timeline
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/r2z_plain.log 2>&1 && \
timeout 1100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1600 -c 1800 --csv \
    --log-file gpurun_out/r2z_launches.csv $CMD > gpurun_out/r2z_ncu.log 2>&1
timeout 300 python scripts/graph_timeline.py 32 gpurun_out/r2z_graph_timeline.md > gpurun_out/r2z_timeline.log 2>&1
wc -l gpurun_out/r2z_launches.csv; head -8 gpurun_out/r2z_graph_timeline.md
