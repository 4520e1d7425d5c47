#!/bin/bash
# GPU cycle for the small memory-bound kernels: kernel tests, then cold per-launch times (ncu) of the named kernels
# inside one static-eager step.   usage: bash scripts/cycle_small.sh "<kernel-name regex>"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider > gpurun_out/cycle_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/cycle_tests.log
tail -3 gpurun_out/cycle_tests.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --kernel-name "regex:$1" -c 40 --csv --log-file gpurun_out/small_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/small_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/small_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d=collections.defaultdict(dict)
for r in rows[1:]:
    d[(r[ii], r[ki][:40])][r[mi]]=float(r[vi].replace(',',''))
for (i,k),m in d.items():
    t=m.get('gpu__time_duration.sum',0)/1e3
    by=(m.get('dram__bytes_read.sum',0)+m.get('dram__bytes_write.sum',0))
    print(f"{k:42s} {t:8.1f} us  {by/1e6:8.1f} MB")
PY
