#!/bin/bash
# 8-GPU session (run under `gpurun --gpus 8`): conic batch, synthetic WSI, dense nuclei.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { n=$1; w=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --workload $w "$@" > gpurun_out/multi_${w}_n$n.json 2> gpurun_out/multi_${w}_n$n.err
  echo "$w n=$n rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/multi_${w}_n$n.json").read().strip().splitlines()[-1])
    print("   tiles/s %.0f  ms/step %.2f  cells/s %.0f  e2e %s  extra %s" % (d["value"], d["ms_per_step"], d.get("cells_per_sec", 0),
          d.get("e2e", {}).get("value"), {k: d[k] for k in ("slide_seconds",) if k in d}))
except Exception as e:
    print("   parse error", e)
PY
}
run 8 conic1024 --steps 20 --warmup 3 --no-cpu-baseline
run 8 wsi --warmup 3
run 8 dense --steps 10 --warmup 3
run 4 tta --steps 10 --warmup 3
