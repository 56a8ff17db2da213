#!/bin/bash
# round 2, session M: A/B of the TMA-staged follow_flows variant against the plain kernel and the pool kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02m
O=gpurun_out/r02m
timeout 900 python -m pytest tests -m gpu -x -q -k "follow" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
for mode in 0 3 2; do
  echo "== CPB_FOLLOW_MERGE=$mode"
  CPB_FOLLOW_MERGE=$mode timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-extras 2>$O/m$mode.err | tee $O/bench_mode$mode.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| follow_flows', round(s['follow_flows'],3), '| cells', d['cells_per_step'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_follow_staged -s 2 -c 1 -f -o $O/prof_k_follow_staged \
    env CPB_FOLLOW_MERGE=3 python bench.py --tiles 1024 --steps 1 --warmup 1 --profile-only > $O/ncu.log 2>&1; echo "ncu rc=$?"
