#!/bin/bash
# round-2 GPU pass 4: parity (faithful oracle on conditioned weights), bench, ncu launch list of one static-eager step
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_train.json 2> gpurun_out/r2d_bench_train.err
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/r2d_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 9000 --csv \
    --log-file gpurun_out/r2d_launches.csv $CMD > gpurun_out/r2d_ncu.log 2>&1
tail -1 gpurun_out/r2d_plain.log | cut -c1-200
wc -l gpurun_out/r2d_launches.csv
tail -3 gpurun_out/r2d_tests.log
