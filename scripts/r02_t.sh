#!/bin/bash
# round 2, session T: follow_flows instruction diet (FADD.RM floor, XORSIGN clamp, paired tap distances) + 256-entry chunks
# for small batches: GPU tests, bench, single-tile breakdown, A/B of the small-chunk switch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02t
O=gpurun_out/r02t
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
SECONDS=0; timeout 600 python bench.py 2>$O/bench.err > $O/bench.json; echo "bench rc=$? wall=${SECONDS}s"; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02t/bench.json"))
print("tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"])
print({k: round(v,3) for k,v in d["stages_ms"].items() if v>0})
print(d.get("hooks_e2e"))
for k,v in d.get("extra_configs",{}).items():
    print(k, round(v.get("tiles_per_sec",0)), {a:b for a,b in v.items() if a in ("ms_per_step","blend_ms","blend_GBs","error")})
PY
echo "== single tile, small chunks on"; timeout 300 python scripts/hooks_breakdown.py 2>&1 | tail -14
echo "== single tile, small chunks off"; CPB_FOLLOW_SMALL=0 timeout 300 python scripts/hooks_breakdown.py 2>&1 | tail -14
