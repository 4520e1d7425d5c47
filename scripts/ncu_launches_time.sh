#!/bin/bash
# Time-only variant of scripts/ncu_launches.sh (gpu__time_duration.sum): one eager step = 1920 launches.
mkdir -p gpurun_out
SKIP=${1:-7680}; COUNT=${2:-1920}
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $COUNT --csv \
    --log-file gpurun_out/launches_time.csv $CMD > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-120
wc -l gpurun_out/launches_time.csv
