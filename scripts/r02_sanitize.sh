#!/bin/bash
# round 2: compute-sanitizer passes over the kernels written this round (float32 screen, contact-label errors, EFT blend +
# fused threshold, batch parts, tile plans, small-chunk trajectory pool)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02san
O=gpurun_out/r02san
echo "== memcheck: parity cases of the new kernels"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q \
   -k "screen or eval_tail or follow_flows_merge or fill_holes_oversized or min_size_zero or baseline_config or average" > $O/memcheck_parity.log 2>&1; echo "memcheck parity rc=$?"; tail -3 $O/memcheck_parity.log
echo "== memcheck: api tests (host path with mapped buffers, tile plans, touching workload, eval_tail)"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_api.py -x -q \
   -k "mapped_logits or single_tile_fast_path or touching or eval_tail or host_buffer_path_matches" > $O/memcheck_api.log 2>&1; echo "memcheck api rc=$?"; tail -3 $O/memcheck_api.log
echo "== racecheck: smoke + the screen case"
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -3 $O/racecheck_smoke.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "screen_is_decision_exact or follow_flows_merge" > $O/racecheck_screen.log 2>&1; echo "racecheck screen rc=$?"; tail -3 $O/racecheck_screen.log
echo "== synccheck: smoke"
timeout 180 compute-sanitizer --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/synccheck_smoke.log 2>&1; echo "synccheck smoke rc=$?"; tail -3 $O/synccheck_smoke.log
