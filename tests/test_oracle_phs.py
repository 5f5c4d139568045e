"""CPU pin of the oracle's plant-hydraulic-stress photosynthesis (oracle/oracle_phs.c) by tests/phs_python.py, an independent
per-patch Python restatement written from PhotosynthesisMod.F90:2704-5228.  Same libm, same operation order: identical bits."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from tests import phs_python as pp
from tests.util import copy_state

dp = C.POINTER(C.c_double)


def _bind(OL):
    P, F = C.POINTER(abi.Params), C.POINTER(abi.STRUCTS["canopyfluxes"])
    OL.oracle_phs_calcstress.argtypes = [P, F, C.c_int, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
    OL.oracle_phs_standalone.argtypes = [P, F, C.c_int, C.POINTER(C.c_int32)] + [dp] * 11
    return OL


def patch_inputs(S, prm, p, nt):
    """the patch's view of the state arrays (p: 0-based patch index), Fortran-indexed"""
    c = S["column"][p] - 1
    g = S["gridcell"][p] - 1
    t = S["itype"][p]
    lev = lambda a, n=4: {i: float(a[i - 1, t]) for i in range(1, n + 1)}       # (segment, pft) tables
    return pp.SimpleNamespace(
        laisun=float(S["laisun"][p]), laisha=float(S["laisha"][p]), htop=float(S["htop"][p]), tsai=float(S["tsai"][p]),
        elai=float(S["elai"][p]), esai=float(S["esai"][p]), fdry=float(S["fdry"][p]), forc_rho=float(S["forc_rho"][c]),
        forc_pbot=float(S["forc_pbot"][c]), tgcm=float(S["thm"][p]),
        psi50=lev(S["pft_psi50"]), ck=lev(S["pft_ck"]), kmax=lev(S["pft_kmax"]),
        k={j: float(S["k_soil_root"][j - 1, p]) for j in range(1, 21)}, smp={j: float(S["smp_l"][j - 1, c]) for j in range(1, 21)},
        z={j: float(S["z"][j + 11, c]) for j in range(1, 21)}, local_time_lt_noon=bool(S["local_time_lt_noon"][g]), c=c, g=g, t=t)


def test_calcstress_matches_python_restatement(oracle_lib):
    OL = _bind(oracle_lib)
    sg, S = synthetic_canopy.make_full_case(300, seed=1101)
    prm = abi.default_params()
    rng = np.random.Generator(np.random.PCG64(1102))
    fe = sg.filters["exposedvegp"]
    S["k_soil_root"] = rng.uniform(1.0e-9, 2.0e-6, S["k_soil_root"].shape) * (rng.random(S["k_soil_root"].shape) < 0.9)
    S["k_soil_root"][0] = 0.0
    dead = fe[::37] - 1
    S["k_soil_root"][:, dead] = 0.0                                  # no root-soil conductance at all: getvegwp's first branch
    nt = S["pft_psi50"].shape[1]
    f = abi.make_struct("canopyfluxes", S, sg.bounds)
    stats = {"night": 0, "itmax": 0, "flag": 0, "iters": 0}
    for n, p1 in enumerate(fe[:900]):
        p = p1 - 1
        P = patch_inputs(S, prm, p, nt)
        night = n % 3 == 0
        x0 = sorted(rng.uniform(-250000.0, -2000.0, 4))              # sun <= sha <= xyl <= root, as the clamps keep them
        x0 = [x0[0], x0[1], x0[2], x0[3]]
        if night:
            x0[0] = 1.0                                              # the reference's night sentinel
        gb_mol = float(rng.uniform(2.0e5, 3.0e6))
        gs_sun, gs_sha = (float(v) for v in rng.uniform(1.0e4, 6.0e5, 2))
        if n % 11 == 0:
            gs_sun = 0.0
        qsatl = float(rng.uniform(0.004, 0.03))
        qaf = qsatl - float(rng.uniform(-0.002, 0.012))
        xv = (C.c_double * 4)(*x0)
        bsun, bsha = C.c_double(-9.0), C.c_double(-9.0)
        S["qflx_tran_veg"][p] = -7.0
        rc = OL.oracle_phs_calcstress(C.byref(prm), C.byref(f), int(p1), xv, C.byref(bsun), C.byref(bsha), gb_mol, gs_sun, gs_sha, qsatl, qaf)
        assert rc == 0
        x = {i + 1: x0[i] for i in range(4)}
        r = pp.calcstress(P, x, gb_mol, gs_sun, gs_sha, qsatl, qaf)
        assert [x[i] for i in range(1, 5)] == list(xv), (p1, [x[i] for i in range(1, 5)], list(xv))
        assert (r.bsun, r.bsha) == (bsun.value, bsha.value), (p1, r.bsun, bsun.value, r.bsha, bsha.value)
        if r.night:
            assert r.tran == S["qflx_tran_veg"][p]
        else:
            assert S["qflx_tran_veg"][p] == -7.0
        assert [r.vegwp_pd[i] for i in range(1, 5)] == [S["vegwp_pd"][i, p] for i in range(4)]
        stats["night"] += r.night
        stats["itmax"] += r.iters > 50
        stats["flag"] += r.iters == 0
        stats["iters"] += r.iters
    assert stats["night"] > 100 and stats["flag"] > 5 and stats["iters"] > 2000, stats
