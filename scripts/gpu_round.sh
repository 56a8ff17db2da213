#!/bin/bash
# One GPU session: smoke, GPU tests, bench, ncu launch list (+ optional full capture).  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
if [ "${SANITIZE:-0}" = "1" ]; then
  echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/memcheck.log
fi
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/launches.csv \
      python bench.py --tiles 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
  python scripts/summarise_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; tail -40 gpurun_out/launches_summary.txt
fi
if [ -n "${NCU_FULL:-}" ]; then
  echo "== ncu full capture of ${NCU_FULL}"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_FULL} -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-2} -f -o gpurun_out/prof_${NCU_FULL} \
      python bench.py --tiles 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
