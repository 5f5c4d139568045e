#!/bin/bash
# tools/gpu_profile.sh TAG: the round's measurement pass on one B200 - gpu tests, default bench, reference arm, ncu launch
# list (time + DRAM bytes per launch) of one full step, and ncu --set full captures of the top kernels, summarised to
# text on the box (the .ncu-rep files are too large to travel back).
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench_f02.json 2> $out/${tag}_bench_f02.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
python bench.py --size f09 --steps 5 --no-cpu > $out/${tag}_bench_f09.json 2>> $out/${tag}_bench_f02.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12000 --csv \
    --log-file $out/${tag}_launches_f02.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $out/${tag}_b.log 2>&1
python tools/launch_summary.py $out/${tag}_launches_f02.csv $out/${tag}_traffic.json f02 > $out/${tag}_launch_summary_f02.txt 2>&1
gzip -f $out/${tag}_launches_f02.csv
cap() {  # name regex skip count routines
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/$1 \
      python bench.py --routines $5 --steps 1 --warmup 3 --no-e2e --no-cpu > $out/${tag}_$1.log 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep > $out/${tag}_ncu_$1.txt 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap newton phs_newton_kernel 4 1 canopyfluxes
cap ci phs_ci_kernel 4 1 canopyfluxes
cap close canopy_close_kernel 2 1 canopyfluxes
cap fric canopy_fric_kernel 1 1 canopyfluxes
cap leaf canopy_leaf_kernel 1 1 canopyfluxes
cap soil "soiltemp_kernel|soilwater_kernel" 0 2 soiltemperature,soilwater
du -sh $out
