#!/bin/bash
# A/B session: trajectory pool vs two-point merge, diffusion job queue vs static map, merge schedules, occupancy.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab_r01b.txt
: > $OUT
build() { nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -I include -I classpose_b200/csrc -o classpose_b200/libclasspose_b200.so classpose_b200/csrc/cpb_api.cu; }
run() { timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('$1', '| tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| follow', round(s['follow_flows'],3), '| diffuse', round(s['diffuse'],3), '| e2e', round(d['e2e']['value']))" | tee -a $OUT; }
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT; tail -3 gpurun_out/pytest_gpu.log
CPB_FOLLOW_MERGE=1 CPB_DIFFUSE_QUEUE=0 run "old merge, static diffuse"
CPB_FOLLOW_MERGE=2 CPB_DIFFUSE_QUEUE=0 run "pool,      static diffuse"
CPB_FOLLOW_MERGE=1 CPB_DIFFUSE_QUEUE=1 run "old merge, queue diffuse "
CPB_FOLLOW_MERGE=2 CPB_DIFFUSE_QUEUE=1 run "pool,      queue diffuse "
CPB_FOLLOW_SCHEDULE=24,36,48,64,80,96,128,160 run "pool S2 (8 merges)"
CPB_FOLLOW_SCHEDULE=28,40,56,72,96,128 run "pool S3 (6 merges)"
CPB_FOLLOW_SCHEDULE=20,28,36,44,52,60,68,76,84,92,100,116,132,148,164,180 run "pool S4 (16 merges)"
CPB_FOLLOW_SCHEDULE=16,24,32,40,48,56,64,72,80,96,112,128,144,160,176 run "pool S5 (15 merges)"
for mb in 5 7 8; do build -DCPB_FP_MINBLOCKS=$mb; run "pool minblocks=$mb"; done
build
for k in k_follow_pool k_diffuse_warp_q; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k} -s 1 -c 1 -f -o gpurun_out/prof_${k} \
      python bench.py --tiles 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_${k}.log 2>&1; echo "ncu full ${k} rc=$?"
done
