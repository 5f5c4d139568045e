#!/bin/bash
# tools/gpu_profile.sh TAG: gpu tests + bench + ncu launch list + ncu --set full of the top kernels, summarised to text
# on the box (the .ncu-rep files are too large to travel back).  Experiment aid.
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest.log
python bench.py --steps 5 --no-cpu > $out/${tag}_bench_f02.json 2> $out/${tag}_bench_f02.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches_f02.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $out/${tag}_b.log 2>&1
cap() {  # name regex count routines
  ncu --set full --clock-control none --import-source on -k regex:"$2" -c $3 -f -o /tmp/$1 \
      python bench.py --routines $4 --steps 1 --warmup 3 --no-e2e --no-cpu > $out/${tag}_$1.log 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep > $out/${tag}_ncu_$1.txt 2>&1
  ncu -i /tmp/$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $out/${tag}_src_$1.csv.gz
  rm -f /tmp/$1.ncu-rep
}
cap phs canopy_phs_kernel 3 canopyfluxes
cap step canopy_step_kernel 3 canopyfluxes
cap soil "soiltemp_kernel|soilwater_kernel" 2 soiltemperature,soilwater
du -sh $out
