#!/bin/bash
cd "$(dirname "$0")/.."
for tiles in 256 512 1024; do for parts in 1 2 4; do
  CPB_BATCH_PARTS=$parts timeout 200 python bench.py --tiles $tiles --steps 20 --profile-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tiles $tiles parts $parts:', round(d['value']), 'tiles/s', round(d['ms_per_step'],3), 'ms')"
done; done
