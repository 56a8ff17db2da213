#!/bin/bash
# scaling points N = 2 and 4 of the main bench (run under `gpurun --gpus 4`)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for n in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/multi_conic1024_n$n.json 2> gpurun_out/multi_conic1024_n$n.err
  echo "conic n=$n rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/multi_conic1024_n$n.json').read().strip().splitlines()[-1])
print('   tiles/s %.0f  ms/step %.2f  cells/s %.0f  e2e %.0f' % (d['value'], d['ms_per_step'], d['cells_per_sec'], d['e2e']['value']))"
done
