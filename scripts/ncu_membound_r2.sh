#!/bin/bash
# ncu --set full of the generator's memory-bound kernels (the ones furthest below the HBM roofline)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain_mb.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_mb.log; exit 1; }
cap() {
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -f -o gpurun_out/$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap r2_upsample_fwd "upsample2x_fwd_kernel<.bool.0>" 6 2
cap r2_upsample_bwd "upsample2x_bwd_kernel" 18 2
cap r2_style_bwd "^irfd::style_bwd_kernel" 36 2
cap r2_maxpool_bwd "maxpool_bwd_kernel" 3 1
ls -la gpurun_out/r2_*.ncu-rep
