#!/bin/bash
# A/B session 4: register-resident diffusion columns, tap reuse in the Euler step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab_r01e.txt
: > $OUT
build() { nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -I include -I classpose_b200/csrc -o classpose_b200/libclasspose_b200.so classpose_b200/csrc/cpb_api.cu; }
run() { timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('$1', '| tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| follow', round(s['follow_flows'],3), '| diffuse', round(s['diffuse'],3), '| flow_err', round(s['flow_err'],3), '| final', round(s['final_map'],3), '| e2e', round(d['e2e']['value']))" | tee -a $OUT; }
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT; tail -3 gpurun_out/pytest_gpu.log
run "default (reg diffusion 16+24, tap reuse)"
CPB_DIFFUSE_REG=0 run "smem diffusion"
CPB_DIFFUSE_REG_CLASSES=1 run "reg diffusion, one class (24)"
CPB_QC_FUSED=0 run "reg diffusion, qc unfused"
build -DCPB_DR24_MINBLOCKS=3 -DCPB_DR16_MINBLOCKS=4; run "reg minblocks 16:4 24:3"
build -DCPB_DR24_MINBLOCKS=5 -DCPB_DR16_MINBLOCKS=6; run "reg minblocks 16:6 24:5"
build -DCPB_FP_MINBLOCKS=5; run "pool minblocks=5"
build
for k in k_diffuse_reg k_follow_pool; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k} -s 2 -c 2 -f -o gpurun_out/prof_${k} \
      python bench.py --tiles 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_${k}.log 2>&1; echo "ncu full ${k} rc=$?"
done
