#!/bin/bash
# tools/gpu_refresh.sh TAG [notest]: gpu tests + default bench + f09 bench + config-5 bench + ncu launch list (time, DRAM bytes, FP64
# pipe counters) of one f02 step, summarised by tools/launch_summary.py.  Everything lands in gpurun_out/; copy what is to be judged to profiles/.
tag=${1:-r02}; out=gpurun_out; mkdir -p $out
if [ "$2" != "notest" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest.log
fi
timeout 600 python bench.py > $out/${tag}_bench_f02.json 2> $out/${tag}_bench_f02.err
timeout 300 python bench.py --size f09 --steps 5 --no-cpu > $out/${tag}_bench_f09.json 2>> $out/${tag}_bench_f02.err
timeout 300 python bench.py --size f19 --members 32 --steps 5 --no-cpu > $out/${tag}_bench_f19x32.json 2>> $out/${tag}_bench_f02.err
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,sm__thread_inst_executed_pipe_fp64_pred_on.sum,sm__pipe_fp64_cycles_active.sum \
    --clock-control none -c 20000 --csv --log-file $out/${tag}_launches_f02.csv python bench.py --steps 1 --warmup 1 --under-profiler --no-e2e --no-cpu > $out/${tag}_b.log 2>&1
python tools/launch_summary.py $out/${tag}_launches_f02.csv $out/${tag}_traffic.json f02 2 > $out/${tag}_launch_summary_f02.txt 2>&1
gzip -f $out/${tag}_launches_f02.csv
tail -2 $out/${tag}_pytest.log; head -30 $out/${tag}_launch_summary_f02.txt; tail -3 $out/${tag}_bench_f02.err
