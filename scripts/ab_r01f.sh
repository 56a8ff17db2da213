#!/bin/bash
# A/B session 5 (small kernels), then the final single-GPU evidence of the round.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab_r01f.txt
: > $OUT
build() { nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -I include -I classpose_b200/csrc -o classpose_b200/libclasspose_b200.so classpose_b200/csrc/cpb_api.cu; }
run() { timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('$1', '| tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| prep', round(s['prep_flow'],3), '| follow', round(s['follow_flows'],3), '| diffuse', round(s['diffuse'],3), '| final', round(s['final_map'],3), '| e2e', round(d['e2e']['value']))" | tee -a $OUT; }
run "default"
build -DCPB_PREP_MINBLOCKS=6; run "prep minblocks=6"
build -DCPB_PREP_MINBLOCKS=8; run "prep minblocks=8"
build
bash scripts/gpu_final.sh
