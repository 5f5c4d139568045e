#!/usr/bin/env python
"""Workload statistics of the CanopyFluxes/PHS oracle (test infrastructure): how the work is distributed over
patches and ITERATION passes.  Used to design the GPU mapping."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctsm_b200 import abi, synthetic_canopy
from oracle import oracle
OL = oracle.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
sg, S = synthetic_canopy.make_full_case(n, seed=13)
prm = abi.default_params()
cnt = (C.c_longlong * 8).in_dll(OL, "oracle_phs_counters")
pp = np.zeros(sg.npatch, dtype=np.int64)
OL.oracle_phs_set_patch_counter.argtypes = [C.c_void_p]
OL.oracle_phs_set_patch_counter(pp.ctypes.data)
f = abi.make_struct("canopyfluxes", S, sg.bounds)
fe = sg.filters["exposedvegp"]
st = abi.Status()
OL.oracle_canopyfluxes(C.byref(prm), C.byref(sg.bounds), len(fe), abi.i32p(fe), C.byref(f), C.byref(st))
OL.oracle_phs_set_patch_counter(None)
c = list(cnt)
ni = S["num_iter"][fe - 1]
passes = ni.sum()
print("patches", len(fe), "patch-passes", int(passes), "mean passes", passes / len(fe))
print("calcstress calls %d, Newton iterations %d (%.1f per call), calls hitting itmax %d" % (c[0], c[1], c[1] / max(c[0], 1), c[2]))
print("ci_func calls %d (%.1f per patch-pass), brent calls %d, hybrid outer passes %d" % (c[3], c[3] / passes, c[4], c[5]))
w = pp[fe - 1]
print("Newton iterations per patch: percentiles 50/90/99/99.9/max", np.percentile(w, [50, 90, 99, 99.9, 100]))
print("per patch-pass mean Newton its:", w.sum() / passes)
night = S["parsun_z"][0, fe - 1] <= 0
print("night: mean its/patch", w[night].mean(), " day:", w[~night].mean())
cap = ni >= 41
print("capped patches", int(cap.sum()), "their Newton its/pass", (w[cap] / 41).mean() if cap.any() else 0, " others its/pass", (w[~cap] / ni[~cap]).mean())
print("share of all Newton iterations spent in capped patches: %.1f%%" % (100.0 * w[cap].sum() / w.sum()))
for lo, hi in ((3, 5), (6, 10), (11, 20), (21, 40), (41, 41)):
    m = (ni >= lo) & (ni <= hi)
    print("num_iter %2d-%2d: %6d patches (%.2f%%), Newton its/pass %.1f, night frac %.2f" % (lo, hi, m.sum(), 100 * m.mean(), (w[m] / ni[m]).mean() if m.any() else 0, night[m].mean() if m.any() else 0))
# characteristics of the heaviest patches
order = np.argsort(-w)[:12]
for i in order:
    p = fe[i] - 1; c = S["column"][p] - 1
    print("p=%d its=%d num_iter=%d night=%d ivt=%d elai=%.2f laisun=%.3f laisha=%.3f tsai=%.2f htop=%.1f fdry=%.2f smp[min,max]=(%.3g,%.3g) ksr_sum=%.3g vegwp=%s qaf-def=%.3g bsun=%.2f bsha=%.2f tran=%.3g" % (
        p, w[i], ni[i], night[i], S["itype"][p], S["elai"][p], S["laisun"][p], S["laisha"][p], S["tsai"][p], S["htop"][p], S["fdry"][p],
        S["smp_l"][:20, c].min(), S["smp_l"][:20, c].max(), S["k_soil_root"][:, p].sum(), np.round(S["vegwp"][:, p]), 0.0, S["bsun"][p], S["bsha"][p], S["qflx_tran_veg"][p]))
