#!/bin/bash
# round 2: two GPUs of one box -- GPU tests (incl. two devices in one process) and the bench at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02n2
O=gpurun_out/r02n2
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "two_devices or native_library or host_buffer" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>$O/bench_n2.err > $O/bench_n2.json; echo "bench n2 rc=$?"; tail -2 $O/bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02n2/bench_n2.json"))
print("N=2 tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2), d["e2e"]["variants"]["upload_everything"]["value"])
PY
