"""CPU pins of the oracle's surface-water / infiltration chain (oracle/oracle_hydrology.c):
a vectorised NumPy restatement written from the Fortran (not from the C), and the invariants the routines imply
(water routed at the surface is conserved, every branch of the chain is populated by the synthetic state)."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from tests.util import copy_state
from tests.test_oracle_preflux import case as preflux_case

DENICE = 917.0


def case(n=800, seed=601, wet_every=3):
    sg, S = preflux_case(n, seed, wet_every)
    synthetic_canopy.hydrology_state(sg, S, np.random.Generator(np.random.PCG64(seed + 4)))
    return sg, S


def run_infiltration(OL, prm, sg, S, bounds=None, fn=None, fh=None):
    st = abi.Status()
    f = abi.make_struct("infiltration", S, sg.bounds)
    fn = sg.filters["nolakec"] if fn is None else fn
    fh = sg.filters["hydrologyc"] if fh is None else fh
    return OL.oracle_hydrology_infiltration(C.byref(prm), C.byref(bounds if bounds is not None else sg.bounds), len(fn), abi.i32p(fn),
                                            len(fh), abi.i32p(fh), 0, C.byref(f), C.byref(st))


def infiltration_np(prm, sg, S0):
    """HydrologyNoDrainageMod.F90:297-337 in NumPy, array-at-a-time over filter_hydrologyc (0-based index arrays)."""
    S = copy_state(S0)
    c = sg.filters["hydrologyc"] - 1
    cn = sg.filters["nolakec"] - 1
    dt = prm.dtime
    # SetSoilWaterFractions (SoilHydrologyMod.F90:239-252); dz / h2osoi_ice rows: levels -11..25, soil level j at row j+11
    watsat = S["watsat"][:20, c]
    dz = S["dz"][12:32, c]
    ice = S["h2osoi_ice"][12:32, c]
    vol_ice = np.minimum(watsat, ice / (dz * DENICE))
    S["eff_porosity"][:20, c] = np.maximum(0.01, watsat - vol_ice)
    icefrac = np.minimum(1.0, vol_ice / watsat)
    S["icefrac"][:20, c] = icefrac
    # SetFloodc (:282-291)
    S["qflx_floodc"][cn] = S["forc_flood"][S["col_gridcell"][cn] - 1]
    # SaturatedExcessRunoff (SaturatedExcessRunoffMod.F90:254-281, :344-356)
    perched = (S["frost_table"][c] > S["zwt_perched"][c]) & (S["frost_table"][c] <= S["zwt"][c])
    fsat = S["wtfact"][c] * np.exp(-0.5 * prm.fff * np.where(perched, S["zwt_perched"][c], S["zwt"][c]))
    if prm.crop_fsat_equals_zero:
        fsat = np.where(S["lun_itype"][c] == 2, 0.0, fsat)
    S["fsat"][c] = fsat
    S["fcov"][c] = fsat
    sat_excess = fsat * S["qflx_rain_plus_snomelt"][c]
    S["qflx_sat_excess_surf"][c] = sat_excess
    # SetQflxInputs (SoilHydrologyMod.F90:339-362)
    top = S["qflx_rain_plus_snomelt"][c] + S["qflx_snow_h2osfc"][c] + S["qflx_floodc"][c]
    S["qflx_top_soil"][c] = top
    nosnow = S["snl"][c] >= 0
    fsno = np.where(nosnow, 0.0, S["frac_sno_eff"][c])
    evap = np.where(nosnow, S["qflx_liqevap_from_top_layer"][c], S["qflx_ev_soil_col"][c])
    fh = S["frac_h2osfc"][c]
    in_soil = (1.0 - fh) * (top - sat_excess)
    to_sfc = fh * (top - sat_excess)
    in_soil = in_soil - (1.0 - fsno - fh) * evap
    to_sfc = to_sfc - fh * S["qflx_ev_h2osfc_col"][c]
    S["qflx_in_soil"][c] = in_soil
    S["qflx_top_soil_to_h2osfc"][c] = to_sfc
    # InfiltrationExcessRunoff (InfiltrationExcessRunoffMod.F90:253-259, :296-300)
    q_unsat = np.min(10.0 ** (-prm.e_ice * icefrac[:3]) * S["hksat"][:3, c], axis=0)
    qinmax = (1.0 - fsat) * q_unsat
    S["qinmax"][c] = qinmax
    excess = np.maximum(0.0, in_soil - (1.0 - fh) * qinmax)
    S["qflx_infl_excess"][c] = excess
    # RouteInfiltrationExcess (SoilHydrologyMod.F90:399-419)
    veg = np.isin(S["lun_itype"][c], (1, 2))
    limited = np.where(veg, in_soil - excess, in_soil)
    if prm.h2osfcflag != 0:
        in_sfc = np.where(veg, to_sfc + excess, 0.0)
        excess_surf = np.zeros_like(excess)
    else:
        in_sfc = np.where(veg, to_sfc, 0.0)
        excess_surf = np.where(veg, excess, 0.0)
    S["qflx_in_soil_limited"][c] = limited
    S["qflx_in_h2osfc"][c] = in_sfc
    S["qflx_infl_excess_surf"][c] = excess_surf
    # UpdateH2osfc (SurfaceWaterMod.F90:387-428, :472-501, :541-552)
    h0, thr = S["h2osfc"][c], S["h2osfc_thresh"][c]
    fn = S["frac_h2osfc_nosnow"][c]
    with np.errstate(invalid="ignore"):
        clust = np.where(fn <= prm.pc, 0.0, np.abs(fn - prm.pc) ** prm.mu) if prm.h2osfcflag == 1 else np.zeros_like(fn)
    k_wet = 1.0e-4 * np.sin((np.pi / 180.0) * S["topo_slope"][c])
    surf = np.minimum(k_wet * clust * (h0 - thr), (h0 - thr) / dt)
    surf = np.where((h0 > thr) & (prm.h2osfcflag != 0), surf, 0.0)
    surf = np.where(surf < 1.0e-8, 0.0, surf)
    S["qflx_h2osfc_surf"][c] = surf
    part = h0 + (in_sfc - surf) * dt
    part = np.where(np.abs(part) < 1.0e-13 * np.abs(h0), 0.0, part)
    drain = np.where(part < 0.0, part / dt, np.minimum(fh * qinmax, part / dt) if prm.h2osfcflag != 0 else np.maximum(0.0, part / dt))
    S["qflx_h2osfc_drain"][c] = drain
    h1 = part - drain * dt
    h1 = np.where(np.abs(h1) < 1.0e-13 * np.abs(part), 0.0, h1)
    S["h2osfc"][c] = h1
    # Infiltration (:450-453), TotalSurfaceRunoff (:511-515)
    S["qflx_infl"][c] = limited + drain
    S["qflx_surf"][c] = sat_excess + excess_surf + surf
    return S


@pytest.mark.parametrize("h2osfcflag,crop0", [(1, 0), (0, 1)], ids=["default", "noh2osfc_cropfsat0"])
def test_infiltration_matches_numpy(oracle_lib, h2osfcflag, crop0):
    sg, S = case()
    prm = abi.default_params()
    prm.h2osfcflag, prm.crop_fsat_equals_zero = h2osfcflag, crop0
    ref = copy_state(S)
    assert run_infiltration(oracle_lib, prm, sg, ref) == 0
    exp = infiltration_np(prm, sg, S)
    worst = 0.0
    for fs in abi.FIELDS["infiltration"]:
        a, b = ref[fs.name], exp[fs.name]
        if fs.intent == "IN":
            assert np.array_equal(a, S[fs.name], equal_nan=True), fs.name
            continue
        fin = np.abs(b) < 1e30
        assert np.array_equal(fin, np.abs(a) < 1e30), fs.name
        if not fin.any():
            continue
        scale = float(np.max(np.abs(b[fin]))) + 1e-300
        e = float(np.max(np.abs(a[fin] - b[fin]))) / scale
        worst = max(worst, e)
        assert e <= 1e-13, (fs.name, e)                       # pow / exp / sin differ between libm and NumPy by an ulp at most
    c = sg.filters["hydrologyc"] - 1
    # every branch of the chain is exercised by the synthetic state
    perched = (S["frost_table"][c] > S["zwt_perched"][c]) & (S["frost_table"][c] <= S["zwt"][c])
    assert 0 < perched.sum() < len(c)
    assert (ref["qflx_infl_excess"][c] > 0).any() and (ref["qflx_infl_excess"][c] == 0).any()
    assert (ref["qflx_h2osfc_drain"][c] < 0).any()                             # the surface store driven negative
    assert (S["snl"][c] < 0).any() and (S["snl"][c] == 0).any()
    if h2osfcflag:
        assert (ref["qflx_h2osfc_surf"][c] > 0).any()
    else:
        assert np.all(ref["h2osfc"][c][ref["qflx_h2osfc_drain"][c] >= 0] == 0.0)  # h2osfcflag = 0: the store always drains completely


def test_infiltration_conserves_surface_water(oracle_lib):
    """What reaches the surface either enters the soil, stays in h2osfc, runs off or evaporates:
    qflx_top_soil - evaporation terms = qflx_infl + qflx_surf + d(h2osfc)/dt  (closed to rounding on soil / crop columns)."""
    sg, S = case(1500, 611)
    prm = abi.default_params()
    ref = copy_state(S)
    assert run_infiltration(oracle_lib, prm, sg, ref) == 0
    c = sg.filters["hydrologyc"] - 1
    c = c[np.isin(S["lun_itype"][c], (1, 2))]
    nosnow = S["snl"][c] >= 0
    fsno = np.where(nosnow, 0.0, S["frac_sno_eff"][c])
    evap = np.where(nosnow, S["qflx_liqevap_from_top_layer"][c], S["qflx_ev_soil_col"][c])
    fh = S["frac_h2osfc"][c]
    supply = ref["qflx_top_soil"][c] - (1.0 - fsno - fh) * evap - fh * S["qflx_ev_h2osfc_col"][c]
    used = ref["qflx_infl"][c] + ref["qflx_surf"][c] + (ref["h2osfc"][c] - S["h2osfc"][c]) / prm.dtime
    scale = np.abs(supply) + np.abs(ref["qflx_infl"][c]) + S["h2osfc"][c] / prm.dtime + 1e-12
    assert np.max(np.abs(supply - used) / scale) < 1e-12


def test_infiltration_urban_and_empty(oracle_lib):
    sg, S = case(200, 621)
    prm = abi.default_params()
    ref = copy_state(S)
    z = np.zeros(1, dtype=np.int32)
    assert run_infiltration(oracle_lib, prm, sg, ref, fn=z[:0], fh=z[:0]) == 0
    for k in S:
        assert np.array_equal(ref[k], S[k], equal_nan=True), k
    ref["lun_itype"][sg.filters["hydrologyc"][3] - 1] = 8
    assert run_infiltration(oracle_lib, prm, sg, ref) == 16
