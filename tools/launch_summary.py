#!/usr/bin/env python
"""Per-kernel totals of one step from an ncu launch list (gpu__time_duration.sum, dram__bytes_read/write.sum per launch
of `bench.py --steps 1 --warmup 3`: 4 identical steps, the table is the mean over them)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
idc = h.index("ID")
t = collections.defaultdict(float); rd = collections.defaultdict(float); wr = collections.defaultdict(float); n = collections.Counter()
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
steps = 4.0
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].split("::")[-1].replace("void ", "").rstrip("<")
    v = float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
    if r[mn] == "gpu__time_duration.sum":
        t[name] += v; n[name] += 1
    elif r[mn] == "dram__bytes_read.sum":
        rd[name] += v
    elif r[mn] == "dram__bytes_write.sum":
        wr[name] += v
tt = sum(t.values())
print("%-28s %9s %12s %8s %12s %12s" % ("kernel", "launches", "ms/step", "share", "DRAM rd GB", "DRAM wr GB"))
for name in sorted(t, key=lambda k: -t[k]):
    print("%-28s %9.1f %12.3f %7.1f%% %12.3f %12.3f" % (name, n[name] / steps, t[name] / steps, 100 * t[name] / tt,
                                                      rd[name] / steps / 1e9, wr[name] / steps / 1e9))
print("%-28s %9.1f %12.3f %7.1f%% %12.3f %12.3f" % ("TOTAL", sum(n.values()) / steps, tt / steps, 100.0,
                                                  sum(rd.values()) / steps / 1e9, sum(wr.values()) / steps / 1e9))
print("(ncu serialises launches and flushes caches between them: the SHARES are comparable with the event-timed step, the"
      " absolute times are not)")
if len(sys.argv) > 3:
    # DRAM bytes per call of each routine (bench.py roofline.traffic): tools/launch_summary.py CSV OUT.json SIZE
    import json
    routine_of = lambda k: ("SoilTemperature" if k in ("soiltemp_kernel", "patchmask_kernel") else "SoilWater" if "soilwater" in k else "SoilFluxes" if "soilfluxes" in k else "clm_drv_patch2col" if "patch2col" in k
                            else "VertTranSink_HydStress" if k.startswith("plantsink") else "BalanceCheck" if k.startswith("balance")
                            else "CanopyFluxes")
    per = collections.defaultdict(float)
    for k in t:
        per[routine_of(k)] += (rd[k] + wr[k]) / steps
    json.dump({"size": sys.argv[3], "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over the "
               "kernels of one call (mean of 4 steps of `bench.py --steps 1 --warmup 3`)", "dram_bytes_per_call": per},
              open(sys.argv[2], "w"), indent=1)

