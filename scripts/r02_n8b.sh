#!/bin/bash
# round 2, final build: the bench at N = 8 (weak scaling) and the WSI config sharded over 8 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02n8
O=gpurun_out/r02n8
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>$O/bench_n8.err > $O/bench_n8.json; echo "bench n8 rc=$?"; tail -1 $O/bench_n8.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload wsi --steps 1 --warmup 3 2>$O/wsi_n8.err > $O/wsi_n8.json; echo "wsi rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02n8/bench_n8.json"))
print("N=8 tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2), "| upload-all", round(d["e2e"]["variants"]["upload_everything"]["value"]))
d=json.load(open("gpurun_out/r02n8/wsi_n8.json")); print("wsi", round(d["value"]), d.get("slide_seconds"))
PY
