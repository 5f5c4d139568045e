"""CPU pin of the oracle's ozone routines (oracle/oracle_ozone.c) by a NumPy restatement written from
OzoneMod.F90: identical bits (no transcendentals)."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from tests.util import copy_state
from tests.test_oracle_preflux import case as preflux_case

RGAS = 6.02214e26 * 1.38065e-23


def case(n=600, seed=1001):
    sg, S = preflux_case(n, seed)
    synthetic_canopy.ozone_state(sg, S, np.random.Generator(np.random.PCG64(seed + 7)))
    return sg, S


def run_uptake(OL, prm, sg, S, fe=None, bounds=None):
    st = abi.Status()
    f = abi.make_struct("ozone", S, sg.bounds)
    fe = sg.filters["exposedvegp"] if fe is None else fe
    z = np.zeros(1, np.int32)
    return OL.oracle_calc_ozone_uptake(C.byref(prm), C.byref(bounds if bounds is not None else sg.bounds), len(fe), abi.i32p(fe if len(fe) else z),
                                       C.byref(f), C.byref(st))


def run_stress(OL, sg, S, method, luna=1, fe=None, fn=None):
    st = abi.Status()
    f = abi.make_struct("ozone", S, sg.bounds)
    fe = sg.filters["exposedvegp"] if fe is None else fe
    fn = sg.filters["noexposedvegp"] if fn is None else fn
    z = np.zeros(1, np.int32)
    return OL.oracle_calc_ozone_stress(C.byref(sg.bounds), len(fe), abi.i32p(fe if len(fe) else z), len(fn), abi.i32p(fn if len(fn) else z),
                                       method, luna, C.byref(f), C.byref(st))


def uptake_np(prm, sg, S0):
    """CalcOzoneUptakeOnePoint, OzoneMod.F90:470-509, array-at-a-time"""
    S = copy_state(S0)
    p = sg.filters["exposedvegp"] - 1
    c, g, t = S["column"][p] - 1, S["gridcell"][p] - 1, S["itype"][p]
    dtime = int(prm.dtime)
    conc = S["forc_o3"][g] * 1.e9 * (S["forc_pbot"][c] / (S["forc_th"][c] * RGAS * 0.001))
    tlai, old = S["tlai"][p], S["tlai_old"][p]
    for rs, nm in ((S["rssha"][p], "o3uptakesha"), (S["rssun"][p], "o3uptakesun")):
        flux = conc / (1.67 * rs + S["rb1"][p] + S["ram1"][p])
        crit = np.where(flux < 0.8, 0.0, flux - 0.8)
        perdt = crit * dtime * 0.000001
        with np.errstate(invalid="ignore", divide="ignore"):
            heal = np.where(tlai - old > 0, np.maximum(0.0, ((tlai - old) / tlai) * perdt), 0.0)
        leafturn = np.where(S["pft_evergreen"][t] == 1, 1.0 / (S["pft_leaf_long"][t] * 365.0 * 24.0), 0.0)
        decay = S[nm][p] * leafturn * (dtime / 3600.0)
        S[nm][p] = np.where(tlai > 0.5, np.maximum(0.0, S[nm][p] + perdt - decay - heal), 0.0)
    S["tlai_old"][p] = tlai
    return S


def stress_np(sg, S0, method):
    S = copy_state(S0)
    p, q = sg.filters["exposedvegp"] - 1, sg.filters["noexposedvegp"] - 1
    t = S["itype"][p]
    k = np.where(t > 3, np.where(S["pft_woody"][t] == 0, 2, 1), 0)
    tab = {"v": (np.array([0.8390, 0.8752, 0.8021]), np.array([0.0, 0.0, -0.0009])),
           "g": (np.array([0.7823, 0.9125, 0.7511]), np.array([0.0048, 0.0, 0.0])),
           "jmax": (np.array([1.0, 1.0, 1.0]), np.array([0.0, -0.0037, 0.0]))}
    for which in (("v", "g") if method == 1 else ("jmax",)):
        a, b = tab[which]
        for leaf in ("sha", "sun"):
            u = S["o3uptake" + leaf][p]
            S["o3coef" + which + leaf][p] = np.where(u == 0.0, 1.0, np.maximum(0.0, np.minimum(1.0, a[k] + b[k] * u)))
            S["o3coef" + which + leaf][q] = 1.0
    return S


def test_ozone_uptake_matches_numpy(oracle_lib):
    sg, S = case()
    prm = abi.default_params()
    ref = copy_state(S)
    assert run_uptake(oracle_lib, prm, sg, ref) == 0
    exp = uptake_np(prm, sg, S)
    for f in abi.FIELDS["ozone"]:
        assert np.array_equal(ref[f.name], exp[f.name], equal_nan=True), f.name
    p = sg.filters["exposedvegp"] - 1
    assert (ref["o3uptakesun"][p] == 0).any() and (ref["o3uptakesun"][p] > S["o3uptakesun"][p]).any() and (ref["o3uptakesun"][p] < S["o3uptakesun"][p]).any()
    assert (S["tlai"][p] <= 0.5).any()


@pytest.mark.parametrize("method", [1, 2])
def test_ozone_stress_matches_numpy(oracle_lib, method):
    sg, S = case(600, 1011)
    ref = copy_state(S)
    assert run_stress(oracle_lib, sg, ref, method) == 0
    exp = stress_np(sg, S, method)
    for f in abi.FIELDS["ozone"]:
        assert np.array_equal(ref[f.name], exp[f.name], equal_nan=True), f.name
    if method == 2:                                              # Falk only runs when LUNA does
        off = copy_state(S)
        assert run_stress(oracle_lib, sg, off, 2, luna=0) == 0
        for k in S:
            assert np.array_equal(off[k], S[k], equal_nan=True), k
    assert run_stress(oracle_lib, sg, copy_state(S), 3) == 2
