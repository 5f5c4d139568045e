#!/bin/bash
# tools/gpu_variants.sh SUFFIX...: canopy-only f02 bench (3 steps) for the default library and each variant.  Experiment aid.
for v in "" "$@"; do
  lib=$PWD/ctsm_b200/lib/libctsm_b200${v:+_$v}.so
  [ -f $lib ] || continue
  CTSM_B200_LIB=$lib python bench.py --size ${SIZE:-f02} --routines canopyfluxes --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('variant [$v] ms_per_step', round(d['ms_per_step'],3))
"
done
