#!/bin/bash
# tools/gpu_refresh.sh TAG: gpu tests + default bench + f09 bench + launch list (time, DRAM bytes) of one step.  Experiment aid.
tag=${1:-r01}; out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench_f02.json 2> $out/${tag}_bench_f02.err
python bench.py --size f09 --steps 5 --no-cpu > $out/${tag}_bench_f09.json 2>> $out/${tag}_bench_f02.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12000 --csv \
    --log-file $out/${tag}_launches_f02.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $out/${tag}_b.log 2>&1
python tools/launch_summary.py $out/${tag}_launches_f02.csv $out/${tag}_traffic.json f02 > $out/${tag}_launch_summary_f02.txt 2>&1
gzip -f $out/${tag}_launches_f02.csv
tail -2 $out/${tag}_pytest.log; head -12 $out/${tag}_launch_summary_f02.txt
