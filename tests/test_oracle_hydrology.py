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
    # one column whose surface-water runoff falls between REAL(4) 1.0e-8 (= 9.99999994e-9, what SurfaceWaterMod.F90:499 compares
    # with) and the double 1.0e-8: it keeps its runoff (pc = 0.4, mu = 0.13889: the defaults)
    ck = sg.filters["hydrologyc"][5] - 1
    S["frac_h2osfc_nosnow"][ck], S["topo_slope"][ck] = 1.0, 30.0
    kw = 1.0e-4 * np.sin((np.pi / 180.0) * 30.0) * (1.0 - 0.4) ** 0.13889
    S["h2osfc_thresh"][ck] = 1.0
    S["h2osfc"][ck] = 1.0 + 9.99999997e-9 / kw
    S["frac_h2osfc"][ck] = max(S["frac_h2osfc"][ck], 0.05)
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
    surf = np.where(surf < float(np.float32(1.0e-8)), 0.0, surf)          # SurfaceWaterMod.F90:499: REAL(4) literal
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
    ck = sg.filters["hydrologyc"][5] - 1                     # (the column case() places between the REAL(4) and the double 1.0e-8)
    ref = copy_state(S)
    assert run_infiltration(oracle_lib, prm, sg, ref) == 0
    if h2osfcflag:
        assert 9.99999994e-9 < ref["qflx_h2osfc_surf"][ck] < 1.0e-8
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


# ----------------------------------------------------------------------------------------------------------------------
# PerchedWaterTable / ThetaBasedWaterTable / RenewCondensation and the closing diagnostics of HydrologyNoDrainage
SAT_LEV = float(np.float32(0.9))            # a default-kind literal in both routines (SoilHydrologyMod.F90:1556, :1981)
LO = 11                                      # row of soil level j in SNOSOI arrays: LO + j; in zi (SNOSOI0): LO + 1 + j


def wt_case(n=800, seed=631, saturate=True):
    from tests.test_oracle_snow import case as snow_case
    sg, S = snow_case(n, seed)
    synthetic_canopy.watertable_state(sg, S, np.random.Generator(np.random.PCG64(seed + 6)), saturate)
    return sg, S


def run_water_table(OL, prm, sg, S, fh=None, bounds=None):
    st = abi.Status()
    f = abi.make_struct("watertable", S, sg.bounds)
    fh = sg.filters["hydrologyc"] if fh is None else fh
    z = np.zeros(1, np.int32)
    rc = OL.oracle_water_table(C.byref(prm), C.byref(bounds if bounds is not None else sg.bounds), len(fh), abi.i32p(fh if len(fh) else z), 0,
                               C.byref(f), C.byref(st))
    return rc, st


def run_diagnostics(OL, prm, sg, S, fs, fns, bounds=None, fn=None, fh=None):
    st = abi.Status()
    f = abi.make_struct("hydrodiag", S, sg.bounds)
    fn = sg.filters["nolakec"] if fn is None else fn
    fh = sg.filters["hydrologyc"] if fh is None else fh
    z = np.zeros(1, np.int32)
    p = lambda a: abi.i32p(a if len(a) else z)
    rc = OL.oracle_hydrology_diagnostics(C.byref(prm), C.byref(bounds if bounds is not None else sg.bounds), len(fn), p(fn), len(fs), p(fs),
                                         len(fns), p(fns), len(fh), p(fh), 0, C.byref(f), C.byref(st))
    return rc, st


def water_table_np(prm, sg, S0):
    """SoilHydrologyMod.F90:1525-1641, :1933-2025, :2569-2678 with NumPy, column by column for the searches (they are sequential
    by construction) and array-at-a-time for RenewCondensation"""
    S = copy_state(S0)
    cols = sg.filters["hydrologyc"] - 1
    vol = S["h2osoi_liq"][LO + 1:LO + 21] / (S["dz"][LO + 1:LO + 21] * 1000.0) + S["h2osoi_ice"][LO + 1:LO + 21] / (S["dz"][LO + 1:LO + 21] * 917.0)
    sat = vol / S["watsat"][:20]                               # rows: soil levels 1..20
    t = S["t_soisno"][LO + 1:LO + 21]
    z, zi = S["z"][LO + 1:LO + 21], S["zi"][LO + 1:LO + 22]    # zi rows: levels 0..20
    for c in cols:
        warm = t[:, c] > 273.15
        k_frz = 20 if warm[0] else 1
        hit = np.nonzero(warm[:-1] & ~warm[1:])[0]             # level k-1 warm, level k frozen: k = hit + 2
        if len(hit):
            k_frz = int(hit[0]) + 2
        frost = zi[k_frz - 1, c]
        S["frost_table"][c] = frost
        perched = frost
        if S0["zwt"][c] < frost and not warm[k_frz - 1]:
            pass
        elif k_frz > 1:
            k_perch = 1
            for k in range(k_frz, 0, -1):
                S["h2osoi_vol"][k - 1, c] = vol[k - 1, c]
                if sat[k - 1, c] <= SAT_LEV:
                    k_perch = k
                    break
            if warm[k_frz - 1]:
                k_perch = k_frz
            if k_frz > k_perch:
                s1, s2 = sat[k_perch - 1, c], sat[k_perch, c]
                if s1 > s2:
                    perched = zi[k_perch - 1, c]
                else:
                    m = (z[k_perch, c] - z[k_perch - 1, c]) / (s2 - s1)
                    perched = max(0.0, m * SAT_LEV + (z[k_perch, c] - m * s2))
        S["zwt_perched"][c] = perched
        nb = int(S["nbedrock"][c])
        below = np.nonzero(sat[:nb, c][::-1] <= SAT_LEV)[0]    # searching upwards from bedrock
        if len(below):
            k_zwt = nb - int(below[0])
            S["h2osoi_vol"][k_zwt - 1:nb, c] = vol[k_zwt - 1:nb, c]
        else:
            k_zwt = 1
            S["h2osoi_vol"][:nb, c] = vol[:nb, c]
        if k_zwt == 1:
            zwt = zi[1, c]
        elif k_zwt < nb:
            s1, s2 = sat[k_zwt - 1, c], sat[k_zwt, c]
            m = (z[k_zwt, c] - z[k_zwt - 1, c]) / (s2 - s1)
            zwt = max(0.0, m * SAT_LEV + (z[k_zwt, c] - m * s2))
        else:
            zwt = zi[nb, c]
        S["zwt"][c] = zwt
    bare = cols[S["snl"][cols] + 1 >= 1]
    w = 1.0 - S["frac_h2osfc"][bare]
    S["h2osoi_liq"][LO + 1, bare] = S["h2osoi_liq"][LO + 1, bare] + w * S["qflx_liqdew_to_top_layer"][bare] * prm.dtime
    before = S["h2osoi_ice"][LO + 1, bare] + w * S["qflx_soliddew_to_top_layer"][bare] * prm.dtime
    after = before - w * S["qflx_solidevap_from_top_layer"][bare] * prm.dtime
    S["h2osoi_ice"][LO + 1, bare] = np.where(np.abs(after) < 1e-12 * np.abs(before), 0.0, after)
    return S


def test_water_table_matches_numpy(oracle_lib):
    sg, S = wt_case()
    prm = abi.default_params()
    ref = copy_state(S)
    rc, st = run_water_table(oracle_lib, prm, sg, ref)
    assert rc == 0, st.msg
    exp = water_table_np(prm, sg, S)
    for fs in abi.FIELDS["watertable"]:
        assert np.array_equal(ref[fs.name], exp[fs.name], equal_nan=True), fs.name     # no transcendentals: identical bits
    c = sg.filters["hydrologyc"] - 1
    # every branch is populated: frost table above / below the old water table, perched table by interpolation and at an interface,
    # theta-based table at the top, by interpolation, at bedrock; condensation on snow-free columns incl. exact sublimation
    assert ((ref["zwt_perched"][c] != ref["frost_table"][c]).sum() > 20) and (ref["zwt_perched"][c] == ref["frost_table"][c]).sum() > 20
    zi = S["zi"][LO + 1:LO + 22]
    at_bed = ref["zwt"][c] == zi[S["nbedrock"][c], c]
    at_top = ref["zwt"][c] == zi[1, c]
    assert at_bed.sum() > 10 and at_top.sum() > 10 and (~at_bed & ~at_top).sum() > 50
    bare = c[S["snl"][c] == 0]
    assert (ref["h2osoi_ice"][LO + 1, bare] == 0.0).sum() > 3 and (ref["h2osoi_liq"][LO + 1, bare] != S["h2osoi_liq"][LO + 1, bare]).any()
    snowy = c[S["snl"][c] < 0]
    assert np.array_equal(ref["h2osoi_ice"][LO + 1, snowy], S["h2osoi_ice"][LO + 1, snowy])
    # failure: sublimation far beyond the layer's ice
    bad = copy_state(S)
    bad["qflx_solidevap_from_top_layer"][bare[2]] = 10.0
    rc, st = run_water_table(oracle_lib, prm, sg, bad)
    assert rc == 18 and st.subgrid_index == bare[2] + 1 and b"RenewCondensation" in st.msg


def diagnostics_np(prm, sg, S0, fs, fns):
    """HydrologyNoDrainageMod.F90:420-757 in NumPy"""
    S = copy_state(S0)
    cn, ch, cs, cns = sg.filters["nolakec"] - 1, sg.filters["hydrologyc"] - 1, fs - 1, fns - 1
    S["snow_persistence"][cs] = S["snow_persistence"][cs] + prm.dtime
    S["snow_persistence"][cns] = 0.0
    lev = np.arange(-11, 1)[:, None]
    ice, liq, t = S["h2osoi_ice"], S["h2osoi_liq"], S["t_soisno"]
    for k in ("snowice", "snowliq", "t_sno_mul_mss"):
        S[k][cn] = 0.0
    si, sl, tm = np.zeros(len(cs)), np.zeros(len(cs)), np.zeros(len(cs))
    for j in range(12):
        on = lev[j] >= S["snl"][cs] + 1
        si = np.where(on, si + ice[j, cs], si)
        sl = np.where(on, sl + liq[j, cs], sl)
        tm = np.where(on, (tm + ice[j, cs] * t[j, cs]) + liq[j, cs] * 273.15, tm)
    S["snowice"][cs], S["snowliq"][cs], S["t_sno_mul_mss"][cs] = si, sl, tm
    b0, b1 = sg.bounds.begc - 1, sg.bounds.endc
    S["snowdp"][b0:b1] = S["snow_depth"][b0:b1] * S["frac_sno_eff"][b0:b1]
    zi = S["zi"][LO + 1:LO + 22][:, cn]                         # levels 0..20
    tt, dz = t[LO + 1:LO + 21][:, cn], S["dz"][LO + 1:LO + 21][:, cn]
    for name, depth in (("t_soi17cm", 0.17), ("t_soi10cm", 0.1)):
        acc = np.zeros(len(cn))
        for j in range(1, 21):
            full = zi[j] <= depth
            part = (zi[j] > depth) & (zi[j - 1] < depth)
            fracl = np.where(full, 1.0, (depth - zi[j - 1]) / dz[j - 1])
            acc = np.where(full | part, acc + tt[j - 1] * dz[j - 1] * fracl, acc)
        S[name][cn] = acc / depth
    S["tsl"][cn] = tt[0]
    snl = S["snl"][cn]
    ttop = t[snl + 1 + LO, cn]
    fse, fh, th = S["frac_sno_eff"][cn], S["frac_h2osfc"][cn], S["t_h2osfc"][cn]
    S["t_grnd"][cn] = np.where(snl < 0, fse * ttop + (1.0 - fse - fh) * tt[0] + fh * th, (1.0 - fh) * tt[0] + fh * th)
    rural = np.isin(S["lun_itype"][cn], (1, 2))
    S["t_grnd_r"][cn[rural]] = ttop[rural]
    d25 = S["dz"][LO + 1:LO + 26]
    S["h2osoi_vol"][:, cn] = (liq[LO + 1:LO + 26] / (d25 * 1000.0) + ice[LO + 1:LO + 26] / (d25 * 917.0))[:, cn]
    lq, dd = liq[LO + 1:LO + 26][:, ch], d25[:, ch]
    ws, sc, bs = S["watsat"][:, ch], S["sucsat"][:, ch], S["bsw"][:, ch]
    psi = sc * (-9.8e-6) * np.maximum(lq / (dd * 1000.0) / ws, 0.001) ** (-bs)
    S["soilpsi"][:, ch] = np.where(lq > 0.0, np.minimum(np.maximum(psi, -15.0), 0.0), -15.0)
    vol = S["h2osoi_vol"][:, ch]
    s_node = np.minimum(1.0, np.maximum(vol / ws, 0.01))
    S["smp_l"][:, ch] = np.maximum(S["smpmin"][ch][None, :], -sc * s_node ** (-bs))
    zb = S["z"][LO + 1:LO + 26][:, ch] + 0.5 * dd
    watdry = ws * (316230.0 / sc) ** (-1.0 / bs)
    rw, sw, rz = np.zeros(len(ch)), np.zeros(len(ch)), np.zeros(len(ch))
    for name, depth in (("wf", 0.05), ("wf2", 0.17)):           # the accumulators carry over from wf to wf2
        for j in range(25):
            on = zb[j] <= depth
            rw = np.where(on, rw + (vol[j] - watdry[j]) * dd[j], rw)
            sw = np.where(on, sw + (ws[j] - watdry[j]) * dd[j], sw)
            rz = np.where(on, rz + dd[j], rz)
        with np.errstate(invalid="ignore", divide="ignore"):
            S[name][ch] = np.where(rz != 0.0, (rw / rz) / (sw / rz), (vol[0] - watdry[0]) / (ws[0] - watdry[0]))
    S["h2osno_top"][cs] = ice[S["snl"][cs] + 1 + LO, cs] + liq[S["snl"][cs] + 1 + LO, cs]
    S["h2osno_top"][cns] = 0.0
    S["snw_rds"][:, cns] = 0.0
    for k in ("snot_top", "dTdz_top", "snw_rds_top", "sno_liq_top"):
        S[k][cns] = 1.0e36
    return S


def test_hydrology_diagnostics_match_numpy(oracle_lib):
    from tests.test_oracle_snow import snow_filters
    sg, S = wt_case(900, 641)
    S["dz"][LO + 1, ::7] = 0.1                                 # a first soil layer reaching below 0.05 m: wf's rz == 0 branch
    prm = abi.default_params()
    fs, fns = snow_filters(oracle_lib, sg, S)
    ref = copy_state(S)
    rc, st = run_diagnostics(oracle_lib, prm, sg, ref, fs, fns)
    assert rc == 0, st.msg
    exp = diagnostics_np(prm, sg, S, fs, fns)
    for f in abi.FIELDS["hydrodiag"]:
        a, b = ref[f.name], exp[f.name]
        if f.intent == "IN":
            assert np.array_equal(a, S[f.name], equal_nan=True), f.name
            continue
        fin = np.abs(b) < 1e30
        assert np.array_equal(fin, np.abs(a) < 1e30), f.name
        if f.name in ("soilpsi", "smp_l", "wf", "wf2"):         # pow: libm against NumPy, an ulp
            e = np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300))
            assert e <= 1e-13, (f.name, e)
        else:
            assert np.array_equal(a, b, equal_nan=True), f.name
    ch = sg.filters["hydrologyc"] - 1
    assert (S["z"][LO + 1, ch] + 0.5 * S["dz"][LO + 1, ch] > 0.05).sum() > 20
