#!/bin/bash
# round 2, session U: what a packed FP32 instruction costs (microbenchmark) and the A/B of four bit-identical forms of
# the Euler step (build_ab/lib_v{0,1,2,3}.so = -DCPB_EULER_VARIANT=n)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02u
O=gpurun_out/r02u
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/f32x2_issue tests/studies/f32x2_issue.cu && /tmp/f32x2_issue | tee $O/f32x2_issue.txt
cp classpose_b200/libclasspose_b200.so /tmp/lib_keep.so
for v in 0 1 2 3 1 0; do
  cp build_ab/lib_v$v.so classpose_b200/libclasspose_b200.so
  timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-extras 2>$O/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('CPB_EULER_VARIANT=$v tiles/s', round(d['value']), '| ms/step', round(d['ms_per_step'],3), '| follow_flows ms', round(s['follow_flows'],4))" | tee -a $O/ab_euler_variants.txt
done
cp build_ab/lib_v1.so classpose_b200/libclasspose_b200.so
timeout 600 python -m pytest tests -m gpu -x -q -k "follow or fused or golden" 2>&1 | tail -3
