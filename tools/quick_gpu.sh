#!/bin/bash
# quick GPU check used while tuning (experiment aid)
python -m pytest tests/test_gpu_canopy.py -q -x 2>&1 | grep -E "^E|FAILED|passed|failed" | cut -c1-400
for size in f09 f02; do
python bench.py --size $size --routines canopyfluxes --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$size canopy ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value']), 'launches', d['gpu_launches'])
    else: print(line.strip()[:300])
"
done
