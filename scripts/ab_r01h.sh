#!/bin/bash
# A/B session 7 (peek before min/max atomics), then the final evidence of the round.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ab_r01h.txt
: > $OUT
build() { nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -I include -I classpose_b200/csrc -o classpose_b200/libclasspose_b200.so classpose_b200/csrc/cpb_api.cu; }
run() { timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('$1', '| tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| seeds', round(s['seeds'],3), '| lookup', round(s['lookup'],3), '| diffuse', round(s['diffuse'],3), '| fill', round(s['fill_holes'],3), '| recount', round(s['recount_hole_tiles'],3))" | tee -a $OUT; }
build -DCPB_STATS_PEEK=0; run "stats: atomics only"
build; run "stats: peek before min/max atomics (default)"
if [ "$(python - <<'PY'
import re
l=open('gpurun_out/ab_r01h.txt').read().strip().splitlines()
ms=[float(re.search(r'\| ms ([0-9.]+)', x).group(1)) for x in l]
print('peek' if ms[1] <= ms[0] else 'plain')
PY
)" = "plain" ]; then echo "peek is slower: final evidence with -DCPB_STATS_PEEK=0" | tee -a $OUT; build -DCPB_STATS_PEEK=0; fi
bash scripts/gpu_final.sh
