#!/bin/bash
# round 2, session B: GPU tests, bench (e2e modes), ncu of the screen kernels after the instruction diet
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02d
O=gpurun_out/r02d
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
show='import json,sys
d=json.loads(sys.stdin.read()); s=d["stages_ms"]
print("tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]))
print({k: round(v,3) for k,v in s.items() if v>0})
print(d.get("flow_check")); print(d["e2e"])'
timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>$O/bench.err | tee $O/bench.json | python -c "$show"
tail -3 $O/bench.err
echo "== chunk 64";  timeout 600 python bench.py --steps 5 --chunk 64 --no-cpu-baseline 2>>$O/bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['e2e'])"
echo "== chunk 256"; timeout 600 python bench.py --steps 5 --chunk 256 --no-cpu-baseline 2>>$O/bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['e2e'])"
echo "== ncu full k_diffuse32"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_diffuse32 -s 2 -c 1 -f -o $O/prof_k_diffuse32 \
    python bench.py --tiles 1024 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qc_pack -s 2 -c 1 -f -o $O/prof_k_qc_pack \
    python bench.py --tiles 1024 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full2.log 2>&1; echo "ncu full rc=$?"
