#!/bin/bash
# A/B of follow_flows occupancy targets on the GPU box (rebuilds the library with different launch bounds).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
build() { nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -I include -I classpose_b200/csrc -o classpose_b200/libclasspose_b200.so classpose_b200/csrc/cpb_api.cu; }
run() { python bench.py --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['stages_ms']['follow_flows'],3), round(d['ms_per_step'],3))"; }
for mb in 4 5 6 8; do build -DCPB_FM_MINBLOCKS=$mb; run "merge minblocks=$mb"; done
for mb in 4 6 8; do build -DCPB_F_MINBLOCKS=$mb; CPB_FOLLOW_MERGE=0 run "plain minblocks=$mb"; done
build
