#!/bin/bash
# Final single-GPU evidence for a round: tests, bench, ncu launch list, full captures of the two dominant kernels,
# DRAM traffic per kernel at the full batch size.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STEPS=${STEPS:-30} NCU=1 bash scripts/gpu_round.sh
for k in k_follow_pool k_diffuse_warp_q; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k} -s 2 -c 1 -f -o gpurun_out/prof_${k} \
      python bench.py --tiles 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_${k}.log 2>&1; echo "ncu full ${k} rc=$?"
done
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"^k_(follow_pool|diffuse_warp_q|prep_flow_v4|lookup_list|final_vote_v4|seeds|flow_err|fill_holes_warp|recount)" -s 30 -c 30 \
    --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/traffic_bench.log 2>&1
echo "traffic rc=$?"
