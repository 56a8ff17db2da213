#!/bin/bash
# round 2, session K: GPU tests, default bench, launch list of the default bench, full captures of the top kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02k
O=gpurun_out/r02k
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
SECONDS=0; timeout 600 python bench.py 2>$O/bench.err > $O/bench.json; echo "bench rc=$? wall=${SECONDS}s"; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02k/bench.json"))
print("tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"])
print({k: round(v,3) for k,v in d["stages_ms"].items() if v>0})
for k,v in d.get("extra_configs",{}).items():
    print(k, round(v.get("tiles_per_sec",0)), {a:b for a,b in v.items() if a in ("ms_per_step","blend_ms","blend_GBs","flow_check")})
PY
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file $O/launches.csv \
    python bench.py --tiles 1024 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/summarise_launches.py $O/launches.csv > $O/launches_summary.txt 2>&1; tail -45 $O/launches_summary.txt
for k in k_follow_pool k_diffuse32 k_prep_flow_v4 k_final_vote_v4 k_qc_scan32 k_lookup_list; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $O/prof_$k \
      python bench.py --tiles 1024 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
