#!/bin/bash
# round 2: eight GPUs of one box -- the bench at N = 8 and N = 4 (weak scaling), WSI config sharded over 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02n8
O=gpurun_out/r02n8
nvidia-smi topo -m > $O/topo.txt 2>&1
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline 2>$O/bench_n$n.err > $O/bench_n$n.json; echo "bench n$n rc=$?"; tail -1 $O/bench_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload wsi --steps 1 --warmup 3 2>$O/wsi_n8.err > $O/wsi_n8.json; echo "wsi rc=$?"
python - <<'PY'
import json
for n in (8, 4):
    d=json.load(open(f"gpurun_out/r02n8/bench_n{n}.json"))
    print(f"N={n} tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2), "| upload-all", round(d["e2e"]["variants"]["upload_everything"]["value"]))
d=json.load(open("gpurun_out/r02n8/wsi_n8.json")); print("wsi", round(d["value"]), d.get("slide_seconds"))
PY
