#!/bin/bash
# ncu --set full captures of the round-2 top kernels (B200_PROFILING.md recipe: plain run first, same command).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_full.log; exit 1; }
cap() {  # name regex skip count
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -f -o gpurun_out/$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap r2_conv_stats256 "conv_gemm_kernel<.int.256, .int.1>" 40 2
cap r2_conv_dgrad256 "conv_gemm_kernel<.int.256, .int.0>" 45 2
cap r2_wgrad256 "wgrad_gemm_kernel<.int.256>" 50 2
cap r2_bn_bwd_apply "bn_bwd_apply_kernel<.bool.1, .int.1, .bool.1>" 14 2
cap r2_bn_bwd_reduce "bn_bwd_reduce_kernel<.bool.1, .int.1>" 14 2
ls -la gpurun_out/*.ncu-rep
