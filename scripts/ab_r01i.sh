#!/bin/bash
# A/B session 8: batched label loads + L2 prefetch of the flows in the diffusion warp job.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab_r01i.txt
: > $OUT
build() { nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -I include -I classpose_b200/csrc -o classpose_b200/libclasspose_b200.so classpose_b200/csrc/cpb_api.cu; }
run() { timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('$1', '| tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| diffuse', round(s['diffuse'],3), '| follow', round(s['follow_flows'],3))" | tee -a $OUT; }
build -DCPB_DIFFUSE_PREFETCH=0; run "diffuse: row-by-row loads, no prefetch"
build; run "diffuse: 8 rows per trip + L2 prefetch of the flows (default)"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee -a $OUT
