#!/bin/bash
# round 2, session F: GPU tests, full default bench (hooks_e2e + extras), ncu of the blend kernel on the TTA workload
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02g
O=gpurun_out/r02g
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
SECONDS=0; timeout 900 python bench.py 2>$O/bench.err > $O/bench.json; echo "bench rc=$? wall=${SECONDS}s"; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02g/bench.json"))
print("tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"])
print({k: round(v,3) for k,v in d["stages_ms"].items() if v>0})
print(d["flow_check"], d["cpu_baseline"])
print(d.get("hooks_e2e"))
for k,v in d.get("extra_configs",{}).items():
    print(k, {a:b for a,b in v.items() if a not in ("workload",)})
PY
