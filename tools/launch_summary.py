#!/usr/bin/env python
"""tools/launch_summary.py CSV [OUT.json SIZE [STEPS]]

Per-kernel and per-routine totals of one step from an ncu launch list of `bench.py --steps 1 --warmup 3 --no-e2e --no-cpu`
(4 identical steps: the table is the mean over them).  Metrics per launch: gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum and, when present, the FP64 pipe counters sm__inst_executed_pipe_fp64.sum (warp instructions),
sm__thread_inst_executed_pipe_fp64_pred_on.sum (thread instructions), sm__pipe_fp64_cycles_active.sum (pipe-busy cycles
summed over the SMs).  OUT.json is what bench.py's `roofline.traffic` / `roofline.fp64` read."""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
steps = float(sys.argv[4]) if len(sys.argv) > 4 else 4.0
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
M = {"gpu__time_duration.sum": "ms", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr",
     "sm__inst_executed_pipe_fp64.sum": "f64w", "sm__thread_inst_executed_pipe_fp64_pred_on.sum": "f64t",
     "sm__pipe_fp64_cycles_active.sum": "f64c"}
acc = collections.defaultdict(lambda: collections.defaultdict(float))
n = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] not in M:
        continue
    name = r[kn].split("(")[0].split("::")[-1].replace("void ", "").split("<")[0]
    v = float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
    acc[name][M[r[mn]]] += v
    if r[mn] == "gpu__time_duration.sum":
        n[name] += 1


def routine_of(k):
    if k.startswith("soiltemp") or k == "patchmask_kernel":
        return "SoilTemperature"
    if "soilwater" in k:
        return "SoilWater"
    if "soilfluxes" in k:
        return "SoilFluxes"
    if "patch2col" in k:
        return "clm_drv_patch2col"
    if k.startswith("plantsink"):
        return "VertTranSink_HydStress"
    if k.startswith("balance"):
        return "BalanceCheck"
    if k.startswith("water"):
        return "WaterBalance"
    if k.startswith(("canopy_", "phs_", "nt_", "split_")):
        return "CanopyFluxes"
    return "other"


tt = sum(a["ms"] for a in acc.values())
print("%-26s %8s %10s %7s %10s %10s %12s %12s %9s" % ("kernel", "launches", "ms/step", "share", "DRAM rd GB", "DRAM wr GB",
                                                   "fp64 warp-i", "fp64 thr-i", "pipe busy"))
for name in sorted(acc, key=lambda k: -acc[k]["ms"]):
    a = acc[name]
    # pipe-busy share: busy cycles summed over the 4 x 148 SM sub-partitions / (4 x 148 x elapsed cycles); elapsed cycles are not in the list, so the
    # share is given against the nominal 1.965 GHz (an under-estimate when the clock sat lower)
    busy = a["f64c"] / (4.0 * 148.0 * a["ms"] * 1e-3 * 1.965e9) if a["ms"] > 0 else 0.0
    print("%-26s %8.1f %10.3f %6.1f%% %10.3f %10.3f %12.4g %12.4g %8.1f%%" % (
        name, n[name] / steps, a["ms"] / steps, 100 * a["ms"] / tt, a["rd"] / steps / 1e9, a["wr"] / steps / 1e9,
        a["f64w"] / steps, a["f64t"] / steps, 100 * busy))
tot = collections.defaultdict(float)
for a in acc.values():
    for k, v in a.items():
        tot[k] += v
print("%-26s %8.1f %10.3f %6.1f%% %10.3f %10.3f %12.4g %12.4g" % ("TOTAL", sum(n.values()) / steps, tt / steps, 100.0,
                                                                 tot["rd"] / steps / 1e9, tot["wr"] / steps / 1e9,
                                                                 tot["f64w"] / steps, tot["f64t"] / steps))
print("(ncu serialises launches and flushes caches between them: the SHARES are comparable with the event-timed step, the"
      " absolute times are not)")
per = collections.defaultdict(lambda: collections.defaultdict(float))
for k, a in acc.items():
    for m, v in a.items():
        per[routine_of(k)][m] += v / steps
print()
print("%-26s %10s %7s %10s %12s %12s %10s" % ("routine", "ms/step", "share", "DRAM GB", "fp64 warp-i", "fp64 thr-i", "lanes/inst"))
for r in sorted(per, key=lambda k: -per[k]["ms"]):
    a = per[r]
    print("%-26s %10.3f %6.1f%% %10.3f %12.4g %12.4g %10.1f" % (r, a["ms"], 100 * a["ms"] * steps / tt, (a["rd"] + a["wr"]) / 1e9,
                                                              a["f64w"], a["f64t"], a["f64t"] / a["f64w"] if a["f64w"] else 0.0))
if len(sys.argv) > 3:
    json.dump({"size": sys.argv[3],
               "source": "ncu launch list of `bench.py --steps 1 --warmup 3 --no-e2e --no-cpu` (--clock-control none), per-launch "
                         "metrics summed over the kernels of one call of each routine, mean of %d steps" % int(steps),
               "dram_bytes_per_call": {r: a["rd"] + a["wr"] for r, a in per.items()},
               "fp64_warp_inst_per_call": {r: a["f64w"] for r, a in per.items()},
               "fp64_thread_inst_per_call": {r: a["f64t"] for r, a in per.items()},
               "fp64_pipe_busy_cycles_per_call": {r: a["f64c"] for r, a in per.items()},
               "ncu_ms_per_call": {r: a["ms"] for r, a in per.items()},
               "launches_per_step": sum(n.values()) / steps},
              open(sys.argv[2], "w"), indent=1)
