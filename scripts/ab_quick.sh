#!/bin/bash
# quick check: GPU tests + one bench line with the stage table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| e2e', round(d['e2e']['value']))
print({k: round(v,3) for k,v in s.items() if v>0})
print(d['roofline'])"
done
