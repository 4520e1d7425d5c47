#!/bin/bash
# Experiment: row-streaming kernels sized for 1/2/4/8 blocks per SM (IRFD_ROW_WAVES), plus ncu --set full of the BN
# backward / style backward kernels at the default setting.
mkdir -p gpurun_out
for w in 1 2 4 8; do
  IRFD_ROW_WAVES=$w python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/waves_$w.json 2> gpurun_out/waves_$w.err
  python - <<P
import json
d=json.loads(open('gpurun_out/waves_$w.json').read().strip().splitlines()[-1])
print('waves $w', round(d['value'],1), 'pairs/s', d['roofline']['hbm_families'])
P
done
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
cap() {
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -f -o gpurun_out/$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap bn_bwd_reduce "bn_bwd_reduce_kernel" 150 3
cap bn_bwd_apply "bn_bwd_apply_kernel" 150 3
cap style_bwd "style_bwd_kernel" 0 2
ls -la gpurun_out/*.ncu-rep
