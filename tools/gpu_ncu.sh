#!/bin/bash
# tools/gpu_ncu.sh TAG "name:regex:skip:count" ...: ncu --set full captures of canopy-only bench kernels, summarised on the box
tag=$1; shift
out=gpurun_out; mkdir -p $out
for spec in "$@"; do
  IFS=: read name regex skip count <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $count -f -o /tmp/$name \
      python bench.py --routines ${ROUTINES:-canopyfluxes} --steps 1 --warmup 3 --no-e2e --no-cpu > $out/${tag}_$name.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep > $out/${tag}_ncu_$name.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $out/${tag}_src_$name.csv.gz
  rm -f /tmp/$name.ncu-rep
done
du -sh $out
