#!/bin/bash
# tools/gpu_check.sh TAG [LIBSUFFIX...]: gpu parity tests, then canopy-only and full-step bench + per-kernel launch totals
# for the default library and each variant library ctsm_b200/lib/libctsm_b200_SUFFIX.so.  Experiment aid.
tag=${1:-chk}; shift
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
for v in "" "$@"; do
  lib=$PWD/ctsm_b200/lib/libctsm_b200${v:+_$v}.so
  [ -f $lib ] || continue
  for size in ${SIZES:-f02}; do
    CTSM_B200_LIB=$lib python bench.py --size $size --routines canopyfluxes --steps 3 --warmup 3 --no-e2e --no-cpu > $out/${tag}_bench_${size}_canopy${v:+_$v}.json 2>> $out/${tag}_err.log
  done
  CTSM_B200_LIB=$lib ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $out/${tag}_launches${v:+_$v}.csv \
      python bench.py --routines canopyfluxes --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>> $out/${tag}_err.log
done
python - <<PY
import json,glob,csv,collections
for fn in sorted(glob.glob('$out/${tag}_bench_*.json')):
    try:
        d=json.loads(open(fn).read().strip().splitlines()[-1]); print(fn, 'ms_per_step', round(d['ms_per_step'],3), 'launches', d['gpu_launches'])
    except Exception as e: print(fn, 'FAILED', e)
for fn in sorted(glob.glob('$out/${tag}_launches*.csv')):
    rows=list(csv.reader(open(fn)))
    hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r]
    if not hi: print(fn,'no data'); continue
    h=rows[hi[0]]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
    tot=collections.defaultdict(float); cnt=collections.Counter()
    for r in rows[hi[0]+1:]:
        if len(r)<=mv: continue
        n=r[kn].split('(')[0].split('::')[-1]; tot[n]+=float(r[mv].replace(',','')); cnt[n]+=1
    print(fn)
    for n,v in sorted(tot.items(), key=lambda x:-x[1]): print('   %-28s %6d launches %10.3f ms/step'%(n,cnt[n]//4,v/1e6/4))
PY
