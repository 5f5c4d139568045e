#!/usr/bin/env python
"""Debug aid (test infrastructure): per-field error table, CUDA CanopyFluxes vs oracle."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctsm_b200 import abi, synthetic_canopy
from oracle import oracle
from tests.util import copy_state
from tests.test_gpu_canopy import run_oracle, run_gpu

size, seed = int(sys.argv[1]), int(sys.argv[2])
L, OL = abi.lib(), oracle.lib()
prm = abi.default_params()
ctx = C.c_void_p()
assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
sg, S = synthetic_canopy.make_full_case(size, seed=seed)
ref, got = copy_state(S), copy_state(S)
print(run_oracle(OL, prm, sg, ref)[0], run_gpu(L, ctx, sg, got, abi.MEM_DEVICE)[0])
fe = sg.filters["exposedvegp"] - 1
ties = got["num_iter"][fe] != ref["num_iter"][fe]
print("patches", len(fe), "ties", int(ties.sum()), fe[ties][:10], got["num_iter"][fe][ties][:10], ref["num_iter"][fe][ties][:10])
cap = np.zeros(sg.npatch, dtype=bool); cap[fe[ref["num_iter"][fe] >= 41]] = True
print("capped", int(cap.sum()))
rows = []
for fs in abi.FIELDS["canopyfluxes"]:
    if fs.intent == "IN" or fs.ctype == "int":
        continue
    a, b = got[fs.name], ref[fs.name]
    fin = (np.abs(b) < 1e30) & (np.abs(a) < 1e30)
    if not fin.any():
        continue
    scale = float(np.max(np.abs(b[fin])))
    if fs.sub == 'PATCH':
        fin = fin & ~cap
    e = np.where(fin, np.abs(a - b) / np.maximum(np.abs(b), 1e-6 * scale + 1e-300), 0.0)
    idx = np.unravel_index(np.argmax(e), e.shape)
    rows.append((float(e.max()), fs.name, idx, float(a[idx]), float(b[idx]), int((e > 1e-10).sum())))
rows.sort(reverse=True)
for r in rows[:25]:
    print("%.3e %-18s idx=%s gpu=%.17g ref=%.17g n>1e-10=%d" % r)
# distribution of t_veg error
e = np.abs(got["t_veg"][fe] - ref["t_veg"][fe]) / ref["t_veg"][fe]
print("t_veg relerr percentiles", np.percentile(e, [50, 90, 99, 100]))
p = rows[0][2][-1]
print("worst patch", p, "night", S["parsun_z"][0, p] <= 0, "itype", S["itype"][p], "num_iter", got["num_iter"][p], ref["num_iter"][p],
      "laisun", S["laisun"][p], "elai", S["elai"][p])
for k in ("t_veg", "qflx_evap_veg", "qflx_tran_veg", "rssun", "rssha", "bsun", "bsha", "gs_mol_sun", "gs_mol_sha", "taf", "qaf", "ustar", "btran"):
    a, b = got[k][..., p], ref[k][..., p]
    print(k, a, b, np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
