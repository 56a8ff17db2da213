#!/bin/bash
# single-GPU numbers of the other BASELINE configs (wsi shard loop, TTA blend, dense nuclei)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in wsi tta dense; do
  timeout 600 python bench.py --gpus 1 --workload $w --steps 10 --warmup 3 > gpurun_out/extra_$w.json 2> gpurun_out/extra_$w.err; echo "$w rc=$?"
  python -c "
import json
d=json.loads(open('gpurun_out/extra_$w.json').read().strip().splitlines()[-1])
print('  ', d['config']['workload'][:60], '| tiles/s %.0f | ms/step %.2f | cells/s %.0f' % (d['value'], d['ms_per_step'], d['cells_per_sec']), {k: d[k] for k in ('slide_seconds','cells_per_tile','blend_max_abs_err_vs_unblended','cells_vs_unblended') if k in d})"
done
