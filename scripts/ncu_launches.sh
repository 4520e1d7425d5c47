#!/bin/bash
# Per-launch device times of one IRFD train step (B200_PROFILING.md recipe: plain run first, then ncu on the same cmd).
# usage (under gpurun): bash scripts/ncu_launches.sh [skip] [count]
set -x
mkdir -p gpurun_out
SKIP=${1:-13500}; COUNT=${2:-5200}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $COUNT --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-300
wc -l gpurun_out/launches.csv
