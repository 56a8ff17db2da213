#!/bin/bash
# Single-tile latency through the reference-facing hooks (BASELINE configs[0]) + compute-sanitizer passes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/latency.txt
import time, numpy as np, torch, sys
sys.path.insert(0, ".")
from oracle import synth, dynamics as odyn, classpose_ref
from classpose_b200 import dynamics, models
from classpose_b200.engine import get_engine
t = synth.make_tile(1)
eng = get_engine()
for _ in range(5):
    m = dynamics.resize_and_compute_masks(t["dP"], t["cellprob"]); cm, u = models.compute_class_masks(m, t["logits"][:, None])
torch.cuda.synchronize()
n = 50
t0 = time.perf_counter()
for _ in range(n):
    m = dynamics.resize_and_compute_masks(t["dP"], t["cellprob"])
    cm, u = models.compute_class_masks(m, t["logits"][:, None])
dt = (time.perf_counter() - t0) / n
print("hooks B+C, numpy in/out, one 256x256 conic tile: %.3f ms per tile" % (dt * 1e3))
d = {k: torch.from_numpy(t[k][None]).cuda() for k in ("dP", "cellprob", "logits")}
for _ in range(5): eng.compute_masks_batch(d["dP"], d["cellprob"], d["logits"])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n): out = eng.compute_masks_batch(d["dP"], d["cellprob"], d["logits"])
torch.cuda.synchronize(); dt2 = (time.perf_counter() - t0) / n
print("device-resident fused call, one tile: %.3f ms" % (dt2 * 1e3))
t0 = time.perf_counter()
for _ in range(5):
    mo = odyn.resize_and_compute_masks(t["dP"], t["cellprob"]); classpose_ref.compute_class_masks(mo, t["logits"][:, None])
print("oracle port on one core-ish (torch threads=%d): %.1f ms per tile" % (torch.get_num_threads(), (time.perf_counter() - t0) / 5 * 1e3))
PY
echo "== memcheck (smoke + selected parity cases)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or contours or dedup or fill_holes or merge or average or flows or get_masks or qc" > gpurun_out/memcheck2.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck2.log
echo "== racecheck (smoke)"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log
