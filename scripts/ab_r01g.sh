#!/bin/bash
# A/B session 6 (lookup block statistics, diffusion occupancy / front restriction), then the final evidence.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab_r01g.txt
: > $OUT
build() { nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -I include -I classpose_b200/csrc -o classpose_b200/libclasspose_b200.so classpose_b200/csrc/cpb_api.cu; }
run() { timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('$1', '| tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| seeds', round(s['seeds'],3), '| lookup', round(s['lookup'],3), '| diffuse', round(s['diffuse'],3), '| fill', round(s['fill_holes'],3), '| recount', round(s['recount_hole_tiles'],3))" | tee -a $OUT; }
run "default (block stats, seed scan, lazy hole plane)"
build -DCPB_LOOKUP_BLOCK_STATS=0; run "lookup: warp-level stats only"
build -DCPB_DQ_MINBLOCKS=7; run "diffuse minblocks=7"
build -DCPB_DQ_MINBLOCKS=6; run "diffuse minblocks=6"
build -DCPB_DIFFUSE_FRONT=1; run "diffuse front restriction"
build
bash scripts/gpu_final.sh
