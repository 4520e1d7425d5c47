#!/bin/bash
# ncu --set full of the GEMM kernels on selected single shapes (scripts/bench_conv_shapes.py --only ...)
mkdir -p gpurun_out
cap() {  # name, shape filter
  python scripts/bench_conv_shapes.py --only "$2" --reps 1 > gpurun_out/shape_$1.log 2>&1 || { tail -3 gpurun_out/shape_$1.log; return; }
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_kernel|halo_kernel" \
      -f -o gpurun_out/shape_$1 python scripts/bench_conv_shapes.py --only "$2" --reps 1 > gpurun_out/ncu_shape_$1.log 2>&1
  tail -1 gpurun_out/ncu_shape_$1.log
}
cap e64_256 "enc 64->256 1x1"
cap e256_1024 "enc 256->1024 1x1"
cap e64_64 "enc 64->64 3x3"
cap g128_64 "gen 128->64"
ls -la gpurun_out/shape_*.ncu-rep
