#!/usr/bin/env python
"""Where do CUDA and oracle CanopyFluxes differ at large sizes?  Runs both on the same synthetic case and reports, per
patch, the relative error of t_veg / qflx_tran_veg / eflx_sh_veg against num_iter and the PHS iteration load.
usage: canopy_parity_diag.py SIZE [seed]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctsm_b200 import abi, synthetic_canopy
from oracle import oracle
from tests.util import copy_state
from tests.test_gpu_canopy import run_gpu

size = sys.argv[1] if len(sys.argv) > 1 else "f09"
size = int(size) if size.isdigit() else size
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20260103
sg, S = synthetic_canopy.make_full_case(size, seed=seed)
OL = oracle.lib()
OL.oracle_set_num_threads(len(os.sched_getaffinity(0)))
prm = abi.default_params()
from tests.test_gpu_canopy import canopy_sensitivity
got = copy_state(S)
ref, sens = canopy_sensitivity(sg, S, prm)
L = abi.lib()
ctx = C.c_void_p()
assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
rc, st = run_gpu(L, ctx, sg, got, abi.MEM_DEVICE)
assert rc == 0
fe = sg.filters["exposedvegp"] - 1
ni_r, ni_g = ref["num_iter"][fe], got["num_iter"][fe]
print("patches", len(fe), "num_iter differs on", int((ni_r != ni_g).sum()), "capped", int((ni_r >= 41).sum()))
def rel(name, floor):
    a, b = got[name][..., fe], ref[name][..., fe]
    if a.ndim == 2:
        return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor), axis=0)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
errs = {"t_veg": rel("t_veg", 1.0), "qflx_tran_veg": rel("qflx_tran_veg", 1e-7), "eflx_sh_veg": rel("eflx_sh_veg", 1.0),
        "vegwp": rel("vegwp", 1.0), "btran": rel("btran", 1e-2), "taf": rel("taf", 1.0)}
emax = np.max(np.stack(list(errs.values())), axis=0)
for lo, hi in ((3, 5), (6, 10), (11, 15), (16, 20), (21, 30), (31, 40), (41, 41)):
    m = (ni_r >= lo) & (ni_r <= hi)
    if m.any():
        e = emax[m]
        print("num_iter %2d-%2d: %7d patches  max err %.2e  >1e-10: %5d  >1e-8: %4d  >1e-6: %3d" % (
            lo, hi, m.sum(), e.max(), (e > 1e-10).sum(), (e > 1e-8).sum(), (e > 1e-6).sum()))
print("ill-conditioned by the probe (sens > 1e-11):", int((sens > 1e-11).sum()), " GPU beyond 1e-10:", int((emax > 1e-10).sum()),
      " beyond 1e-10 but NOT flagged:", int(((emax > 1e-10) & ~(sens > 1e-11)).sum()))
bad = np.nonzero((emax > 1e-10) & ~(sens > 1e-11))[0]
print("unflagged patches beyond 1e-10:", len(bad))
order = bad[np.argsort(-emax[bad])][:25]
night = S["parsun_z"][0, fe] <= 0
for i in order:
    p = fe[i]
    print("  p=%d sens=%.1e ni=%d/%d night=%d emax=%.2e %s btran=%.4f/%.4f bsun %.4f/%.4f tran=%.3e/%.3e tveg=%.6f/%.6f vegwp_root=%.1f" % (
        p + 1, sens[i], ni_r[i], ni_g[i], night[i], emax[i], {k: "%.1e" % v[i] for k, v in errs.items()},
        ref["btran"][p], got["btran"][p], ref["bsun"][p], got["bsun"][p], ref["qflx_tran_veg"][p], got["qflx_tran_veg"][p],
        ref["t_veg"][p], got["t_veg"][p], ref["vegwp"][3, p]))
L.ctsm_b200_finalize(ctx)
