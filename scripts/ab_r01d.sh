#!/bin/bash
# A/B session 3: packed-FP32 Euler step, 4-row diffusion window, schedules, occupancy targets.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab_r01d.txt
: > $OUT
build() { nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -I include -I classpose_b200/csrc -o classpose_b200/libclasspose_b200.so classpose_b200/csrc/cpb_api.cu; }
run() { timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('$1', '| tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| follow', round(s['follow_flows'],3), '| diffuse', round(s['diffuse'],3), '| flow_err', round(s['flow_err'],3), '| final', round(s['final_map'],3), '| e2e', round(d['e2e']['value']))" | tee -a $OUT; }
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT; tail -3 gpurun_out/pytest_gpu.log
run "default (packed, rows4, sched 32,48,72,112)"
CPB_DIFFUSE_ROWS=2 run "diffuse rows=2"
CPB_FOLLOW_SCHEDULE=32,56,96 run "sched 32,56,96"
CPB_FOLLOW_SCHEDULE=40,72 run "sched 40,72"
CPB_FOLLOW_SCHEDULE=36,56,88,136 run "sched 36,56,88,136"
CPB_FOLLOW_SCHEDULE=28,40,56,72,96,128 run "sched 28,40,56,72,96,128"
CPB_FOLLOW_MERGE=0 run "plain scalar follow"
build -DCPB_DQ4_MINBLOCKS=5; run "rows4 minblocks=5"
build -DCPB_DQ4_MINBLOCKS=7; run "rows4 minblocks=7"
build -DCPB_FP_MINBLOCKS=8; run "pool minblocks=8"
build -DCPB_FP_MINBLOCKS=5; run "pool minblocks=5"
build
for k in k_follow_pool k_diffuse_warp_q; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k} -s 2 -c 1 -f -o gpurun_out/prof_${k} \
      python bench.py --tiles 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_${k}.log 2>&1; echo "ncu full ${k} rc=$?"
done
