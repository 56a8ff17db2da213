#!/bin/bash
# A/B session 2: fused flow error / fused vote / queue switches, merge schedules, full ncu captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab_r01c.txt
: > $OUT
run() { timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('$1', '| tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| follow', round(s['follow_flows'],3), '| diffuse', round(s['diffuse'],3), '| flow_err', round(s['flow_err'],3), '| final', round(s['final_map'],3), '| vote', round(s['vote'],3), '| e2e', round(d['e2e']['value']))" | tee -a $OUT; }
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT; tail -3 gpurun_out/pytest_gpu.log
run "all on (default)      "
CPB_QC_FUSED=0 run "qc unfused            "
CPB_VOTE_FUSED=0 run "vote unfused          "
CPB_DIFFUSE_QUEUE=0 run "static diffuse map    "
CPB_FOLLOW_SCHEDULE=32,48,72,112 run "sched B 32,48,72,112"
CPB_FOLLOW_SCHEDULE=28,44,64,96,144 run "sched C 28,44,64,96,144"
CPB_FOLLOW_SCHEDULE=30,44,60,80,110 run "sched H 30,44,60,80,110"
CPB_FOLLOW_SCHEDULE=32,56,96 run "sched E 32,56,96"
for k in k_follow_pool k_diffuse_warp_q k_final_vote_v4 k_flow_err; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k} -s 2 -c 2 -f -o gpurun_out/prof_${k} \
      python bench.py --tiles 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_${k}.log 2>&1; echo "ncu full ${k} rc=$?"
done
