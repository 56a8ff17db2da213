#!/bin/bash
# round 2, session Y: seed candidates listed by the trajectory kernel while it counts end points (CPB_SEED_CANDS A/B)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02y
O=gpurun_out/r02y
echo skip-pytest
for p in 0 1 0 1; do
  CPB_SEED_CANDS=$p timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extras 2>$O/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('CPB_SEED_CANDS=$p tiles/s', round(d['value']), '| ms/step', round(d['ms_per_step'],3), '| follow', round(s['follow_flows'],3), '| seeds', round(s['seeds'],3), '| stage sum', round(sum(s.values()),3))" | tee -a $O/ab_seed_cands.txt
done
