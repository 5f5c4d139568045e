#!/bin/bash
# runs the canopy-only bench for each library variant (experiment aid)
size=${1:-f02}
for v in "" _nobins _mb6 _mb8; do
  lib=ctsm_b200/lib/libctsm_b200$v.so
  [ -f $lib ] || continue
  CTSM_B200_TAIL_FRAC=0 CTSM_B200_LIB=$PWD/$lib python bench.py --size $size --routines canopyfluxes --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$v $size ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value']), 'launches', d['gpu_launches'])
    else: print(line.strip()[:200])
"
done
