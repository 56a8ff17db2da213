#!/bin/bash
# round 2, final build: the bench at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02n2
O=gpurun_out/r02n2
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>$O/bench_n2.err > $O/bench_n2.json; echo "bench n2 rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02n2/bench_n2.json'))
print('N=2 tiles/s', round(d['value']), '| ms', round(d['ms_per_step'],3), '| e2e', round(d['e2e']['value']))"
