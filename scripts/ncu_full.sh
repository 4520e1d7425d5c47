#!/bin/bash
# ncu --set full captures of the tensor-core GEMM kernels (B200_PROFILING.md recipe: plain run first, same command).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_full.log; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -f -o gpurun_out/$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -2 gpurun_out/ncu_$1.log
}
cap conv_style64 "conv_gemm_kernel<.int.64, .int.2>" 2 2
cap conv_plain256 "conv_gemm_kernel<.int.256, .int.0>" 10 2
cap conv_style256 "conv_gemm_kernel<.int.256, .int.2>" 4 2
cap wgrad64 "wgrad_gemm_kernel<.int.64>" 3 2
cap wgrad256 "wgrad_gemm_kernel<.int.256>" 40 2
ls -la gpurun_out/*.ncu-rep
