#!/bin/bash
# round 2, session L: GPU tests, default bench, clean launch list (device-resident steps only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02s
O=gpurun_out/r02s
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
SECONDS=0; timeout 600 python bench.py 2>$O/bench.err > $O/bench.json; echo "bench rc=$? wall=${SECONDS}s"; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02s/bench.json"))
print("tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"])
print({k: round(v,3) for k,v in d["stages_ms"].items() if v>0})
print(d.get("hooks_e2e"))
for k,v in d.get("extra_configs",{}).items():
    print(k, round(v.get("tiles_per_sec",0)), {a:b for a,b in v.items() if a in ("ms_per_step","blend_ms","blend_GBs","flow_check","error")})
PY
echo "== ncu launch list (device-resident steps only)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file $O/launches.csv \
    python bench.py --tiles 1024 --steps 2 --warmup 1 --profile-only > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/summarise_launches.py $O/launches.csv > $O/launches_summary.txt 2>&1; tail -40 $O/launches_summary.txt
