#!/bin/bash
# ncu --set full captures of the kernels changed late in round 2 (per-warp epilogue, BN reduce folded into the dgrad
# epilogue, BN backward with the mask plane).  Plain run first, same command (B200_PROFILING.md recipe).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_full.log; exit 1; }
cap() {  # name regex skip count
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -f -o gpurun_out/$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap r2b_conv_stats256 "conv_gemm_kernel<.int.256, .int.1" 40 2
cap r2b_conv_bnbwd256 "conv_gemm_kernel<.int.256, .int.4" 12 2
cap r2b_conv_plain256 "conv_gemm_kernel<.int.256, .int.0" 45 2
cap r2b_bn_bwd_reduce "bn_bwd_reduce_kernel<.bool.1, .int.3, .bool.1>" 6 2
cap r2b_bn_bwd_apply "bn_bwd_apply_kernel<.bool.0, .int.0, .bool.0>" 20 2
for n in r2b_conv_stats256 r2b_conv_bnbwd256 r2b_conv_plain256 r2b_bn_bwd_reduce r2b_bn_bwd_apply; do
  echo "## $n"; ncu -i gpurun_out/$n.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_raw_summary.py
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/$n.raw.csv 2>/dev/null
done > gpurun_out/r2b_ncu_full_summary.md
cat gpurun_out/r2b_ncu_full_summary.md
bash scripts/profile_final_r2.sh
python scripts/bench_conv_shapes.py --md gpurun_out/r2b_conv_shapes.md > /dev/null 2>&1; head -6 gpurun_out/r2b_conv_shapes.md
