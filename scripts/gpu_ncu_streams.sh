#!/bin/bash
# full ncu captures of the two HBM-streaming kernels at the full batch size (1024 tiles)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in k_final_vote_v4 k_prep_flow_v4; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:^${k} -s 3 -c 1 -f -o gpurun_out/prof_${k} \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_${k}.log 2>&1; echo "ncu full ${k} rc=$?"
  ncu -i gpurun_out/prof_${k}.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for k,v in zip(h,r):
    if any(s in k for s in ['gpu__time_duration.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum ','dram__bytes_write.sum ','dram__bytes_read.sum.per_second','dram__bytes_write.sum.per_second','sm__warps_active.avg.pct','launch__grid_size']): print('  ',k,'=',v)"
done
