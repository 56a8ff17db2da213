#!/bin/bash
# A/B session 9: exact three-operation division by 5 in the prep kernel (exhaustive check on the GPU first).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab_r01j.txt
: > $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/div5_exhaustive tests/studies/div5_exhaustive.cu && ./gpurun_out/div5_exhaustive | tee -a $OUT
rm -f gpurun_out/div5_exhaustive
build() { nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -I include -I classpose_b200/csrc -o classpose_b200/libclasspose_b200.so classpose_b200/csrc/cpb_api.cu; }
run() { timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('$1', '| tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| prep', round(s['prep_flow'],3))" | tee -a $OUT; }
build -DCPB_DIV5_FAST=0; run "prep: IEEE division"
build
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee -a $OUT
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); s=d['stages_ms']
print('prep: 3-op division (default) | tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| prep', round(s['prep_flow'],3), '| e2e', round(d['e2e']['value']))" | tee -a $OUT
