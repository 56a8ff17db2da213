#!/bin/bash
# round 2, session V: occupancy / chunk-size A/B of the trajectory pool (build_ab/*.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02v
O=gpurun_out/r02v
for v in v1 mb5 mb8 pool512 v1; do
  cp build_ab/lib_$v.so classpose_b200/libclasspose_b200.so
  timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-extras 2>$O/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('$v tiles/s', round(d['value']), '| ms/step', round(d['ms_per_step'],3), '| follow_flows ms', round(s['follow_flows'],4))" | tee -a $O/ab_pool_shape.txt
done
