#!/bin/bash
# tools/gpu_final.sh TAG: the round-end measurement pass on one B200: full GPU test suite, smoke, targeted ncu re-capture of the
# SoilTemperature kernels (merged into profiles/<TAG>_traffic.json), then the bench lines.  Everything lands in gpurun_out/.
tag=${1:-r02}; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,sm__thread_inst_executed_pipe_fp64_pred_on.sum,sm__pipe_fp64_cycles_active.sum \
    --clock-control none -k regex:"soiltemp|patchmask" -c 8 --csv --log-file $out/${tag}_launches_soiltemp.csv \
    python bench.py --routines soiltemperature --steps 1 --warmup 1 --under-profiler --no-e2e --no-cpu > $out/${tag}_b2.log 2>&1
python tools/launch_summary.py $out/${tag}_launches_soiltemp.csv $out/${tag}_traffic_soiltemp.json f02 2 > $out/${tag}_launch_summary_soiltemp.txt 2>&1
python - <<PY
import json
a = json.load(open("profiles/${tag}_traffic.json")); b = json.load(open("$out/${tag}_traffic_soiltemp.json"))
for k, v in b.items():
    if isinstance(v, dict) and "SoilTemperature" in v:
        a[k]["SoilTemperature"] = v["SoilTemperature"]
a["source"] += "; SoilTemperature re-captured after the level-streaming kernel (${tag}_launches_soiltemp.csv)"
json.dump(a, open("profiles/${tag}_traffic.json", "w"), indent=1)
json.dump(a, open("$out/${tag}_traffic.json", "w"), indent=1)
PY
timeout 600 python bench.py > $out/${tag}_bench_f02.json 2> $out/${tag}_bench_f02.err
timeout 400 python bench.py --routines pre,canopyfluxes,soiltemperature,soilfluxes,patch2col,plantsink,soilwater,balancecheck --no-cpu > $out/${tag}_bench_f02_pre.json 2>> $out/${tag}_bench_f02.err
timeout 300 python bench.py --size f09 --steps 5 --no-cpu > $out/${tag}_bench_f09.json 2>> $out/${tag}_bench_f02.err
timeout 300 python bench.py --size f19 --members 32 --steps 5 --no-cpu > $out/${tag}_bench_f19x32.json 2>> $out/${tag}_bench_f02.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench_f02.err
cat $out/${tag}_pytest.log $out/${tag}_smoke.log; cat $out/${tag}_launch_summary_soiltemp.txt | head -8; tail -3 $out/${tag}_bench_f02.err
for f in f02 f02_pre f09 f19x32 reference; do python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench_$f.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("$f", "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "bound", r.get("bound"), "frac", r.get("frac"),
          {k: round(v["ms"], 2) for k, v in (r.get("routines") or {}).items()})
except Exception as e:
    print("$f", "ERR", e)
PY
done
