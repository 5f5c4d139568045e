out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_balance.py -m gpu -q -x -k sink 2>&1 | tail -3
CTSM_B200_SINK_WARP=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > $out/r2n_bench_sink1b.json 2> $out/r2n_bench_sink1b.err
python - <<PY
import json
d = json.loads(open("$out/r2n_bench_sink1b.json").read().strip().splitlines()[-1])
r = d["roofline"]["routines"]
print("step %.2f ms" % d["ms_per_step"], {k: round(x["ms"], 3) for k, x in r.items()})
PY
