#!/bin/bash
# ncu --set full of the GEMM kernels on selected single shapes (scripts/bench_conv_shapes.py --only ...)
# usage: bash scripts/ncu_shapes.sh [tag:filter ...]   default: the 256^2 generator layer and a small encoder layer
mkdir -p gpurun_out
cap() {  # name, shape filter
  python scripts/bench_conv_shapes.py --only "$2" --reps 1 > gpurun_out/shape_$1.log 2>&1 || { tail -3 gpurun_out/shape_$1.log; return; }
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_kernel|halo_kernel" \
      -f -o gpurun_out/shape_$1 python scripts/bench_conv_shapes.py --only "$2" --reps 1 > gpurun_out/ncu_shape_$1.log 2>&1
  tail -1 gpurun_out/ncu_shape_$1.log
}
if [ $# -eq 0 ]; then set -- "g128_64:gen 128->64" "e256_1024:enc 256->1024 1x1"; fi
for spec in "$@"; do cap "${spec%%:*}" "${spec#*:}"; done
ls -la gpurun_out/shape_*.ncu-rep
