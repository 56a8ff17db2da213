#!/bin/bash
# round 2, final N=1 session: GPU tests, default bench, launch lists (four parts / one part), full captures of the kernels
# changed since the last captures (follow_flows step, prep division, lookup ILP), single-tile breakdown
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02final
O=gpurun_out/r02final
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
SECONDS=0; timeout 600 python bench.py 2>$O/bench.err > $O/bench.json; echo "bench rc=$? wall=${SECONDS}s"; tail -3 $O/bench.err
echo "(reference arm: see the earlier session)"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02final/bench.json"))
print("tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], "| link", d["e2e"].get("pcie_link_gbs"), d["e2e"].get("pcie_link_frac"))
print({k: round(v,3) for k,v in d["stages_ms"].items() if v>0})
print(d.get("hooks_e2e"))
print(d["roofline"]["frac"], d["roofline"]["streaming_stages"])
for k,v in d.get("extra_configs",{}).items():
    print(k, round(v.get("tiles_per_sec",0)), {a:b for a,b in v.items() if a in ("ms_per_step","blend_ms","blend_GBs","error")})
PY
echo "== ncu launch list, shipped four batch parts"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file $O/launches.csv \
    python bench.py --tiles 1024 --steps 2 --warmup 1 --profile-only > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/summarise_launches.py $O/launches.csv > $O/launches_summary.txt 2>&1; head -12 $O/launches_summary.txt
echo "== ncu launch list, one part"
CPB_BATCH_PARTS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file $O/launches_one_part.csv \
    python bench.py --tiles 1024 --steps 2 --warmup 1 --profile-only > $O/ncu_bench1.log 2>&1; echo "ncu rc=$?"
python scripts/summarise_launches.py $O/launches_one_part.csv > $O/launches_summary_one_part.txt 2>&1; head -12 $O/launches_summary_one_part.txt
for k in k_follow_pool; do
  CPB_BATCH_PARTS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $O/prof_$k \
      python bench.py --tiles 1024 --steps 2 --warmup 1 --profile-only > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
echo "== single tile"; timeout 300 python scripts/hooks_breakdown.py 2>&1 | tail -13 | tee $O/hooks_breakdown.txt
