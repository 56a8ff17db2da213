#!/bin/bash
# round 2, session Z: the library's own memory pool for stream-ordered scratch (blend tables); final default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02z
O=gpurun_out/r02z
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
for i in 1 2; do
timeout 300 python bench.py --workload tta --steps 20 --no-cpu-baseline 2>$O/tta.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('tta tiles/s', round(d['value']), '| ms/step', round(d['ms_per_step'],3), {k:v for k,v in d.items() if 'blend' in k}, {k:v for k,v in d.get('config',{}).items() if 'blend' in k})" ; tail -2 $O/tta.err
done
SECONDS=0; timeout 600 python bench.py 2>$O/bench.err > $O/bench.json; echo "bench rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02z/bench.json"))
print("tiles/s", round(d["value"]), "| ms", round(d["ms_per_step"],3), "| e2e", round(d["e2e"]["value"]), "| hooks", d["hooks_e2e"]["single_tile_ms_median"], d["hooks_e2e"]["tiles_per_sec"])
for k,v in d.get("extra_configs",{}).items():
    print(k, round(v.get("tiles_per_sec",0)), {a:b for a,b in v.items() if a in ("ms_per_step","blend_ms","blend_GBs","error")})
PY
