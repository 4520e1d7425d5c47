#!/bin/bash
# Per-launch device time AND DRAM bytes of one IRFD train step (B200_PROFILING.md recipe: plain run first, then ncu on
# the same command).  Eager launch mode (--no-graph) so every kernel is an ordinary launch; the kernels are the same
# ones the graph replays.  usage (under gpurun): bash scripts/ncu_launches.sh [skip] [count]
mkdir -p gpurun_out
SKIP=${1:-7000}; COUNT=${2:-2400}
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $SKIP -c $COUNT --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-200
wc -l gpurun_out/launches.csv
