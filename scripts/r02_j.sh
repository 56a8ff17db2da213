#!/bin/bash
# round 2, session I: GPU tests, bench, hooks breakdown, ncu of the new blend kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02j
O=gpurun_out/r02j
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
SECONDS=0; timeout 600 python bench.py 2>$O/bench.err > $O/bench.json; echo "bench rc=$? wall=${SECONDS}s"; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02j/bench.json"))
print("tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"])
print({k: round(v,3) for k,v in d["stages_ms"].items() if v>0})
print(d["flow_check"], d["cpu_baseline"])
print(d.get("hooks_e2e"))
for k,v in d.get("extra_configs",{}).items():
    print(k, {a:b for a,b in v.items() if a not in ("workload",)})
PY
timeout 300 python scripts/hooks_breakdown.py 2>&1 | tee $O/hooks_breakdown.txt
echo "== ncu blend"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_average_tiles -s 4 -c 2 -f -o $O/prof_k_average_tiles_eft \
    python bench.py --workload tta --steps 1 --warmup 1 > $O/ncu_blend.log 2>&1; echo "ncu rc=$?"
