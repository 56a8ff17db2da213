#!/bin/bash
# round 2, session A: GPU tests, bench with the float32 flow-check screen on / off, launch list, full capture of k_diffuse32
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
show='import json,sys
d=json.loads(sys.stdin.read()); s=d["stages_ms"]
print("tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]))
print({k: round(v,3) for k,v in s.items() if v>0})
print(d.get("flow_check"))'
echo "== screen on";  timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>$O/on.err | tee $O/bench_on.json | python -c "$show"
echo "== screen off"; CPB_QC_SCREEN=0 timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>$O/off.err | tee $O/bench_off.json | python -c "$show"
echo "== dense"; timeout 300 python bench.py --workload dense --steps 5 2>$O/dense.err | tee $O/bench_dense.json | cut -c1-400
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file $O/launches.csv \
    python bench.py --tiles 1024 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/summarise_launches.py $O/launches.csv > $O/launches_summary.txt 2>&1; tail -40 $O/launches_summary.txt
echo "== ncu full k_diffuse32"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_diffuse32 -s 2 -c 1 -f -o $O/prof_k_diffuse32 \
    python bench.py --tiles 1024 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qc_pack -s 2 -c 1 -f -o $O/prof_k_qc_pack \
    python bench.py --tiles 1024 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full2.log 2>&1; echo "ncu full rc=$?"
