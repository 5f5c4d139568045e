#!/bin/bash
# quick GPU check used while tuning (experiment aid)
python -m pytest tests/test_gpu_canopy.py -q -x 2>&1 | grep -E "^E|FAILED|passed|failed" | cut -c1-400
for tf in 0 0.1 0.5 1.0; do
for size in f09; do
CTSM_B200_TAIL_FRAC=$tf python bench.py --size $size --routines canopyfluxes --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('tail_frac $tf $size canopy ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value']), 'launches', d['gpu_launches'])
    else: print(line.strip()[:300])
"
done; done
