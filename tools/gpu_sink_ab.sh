#!/bin/bash
# tools/gpu_sink_ab.sh: parity + A/B of the warp-per-column plant sink / patch2col kernels against the thread-per-column ones.
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_balance.py tests/test_gpu_soilfluxes.py -m gpu -q -x 2>&1 | tail -5 > $out/r2n_pytest_sink.log
cat $out/r2n_pytest_sink.log
for v in 1 0; do
  CTSM_B200_SINK_WARP=$v timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > $out/r2n_bench_sink$v.json 2> $out/r2n_bench_sink$v.err
  python - <<PY
import json
d = json.loads(open("$out/r2n_bench_sink$v.json").read().strip().splitlines()[-1])
r = d["roofline"]["routines"]
print("SINK_WARP=$v", "step %.2f ms" % d["ms_per_step"], {k: round(x["ms"], 3) for k, x in r.items()})
PY
done
