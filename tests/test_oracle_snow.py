"""CPU pins of the oracle's snow routines (oracle/oracle_snow.c): an independent per-column Python
restatement written from the Fortran (tests/snow_python.py, not from the C), and the conservation laws the routines imply -
water, enthalpy, aerosol mass, the layer-thickness limits after CombineSnowLayers / DivideSnowLayers."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from tests.util import copy_state
from tests.test_oracle_hydrology import case as hydro_case

NSNO = 12
CPICE, CPLIQ, HFUS, TFRZ = 2.11727e3, 4.188e3, 3.337e5, 273.15
AER = ("bcphi", "bcpho", "ocphi", "ocpho", "dst1", "dst2", "dst3", "dst4")


def case(n=600, seed=801):
    sg, S = hydro_case(n, seed)
    synthetic_canopy.snow_state(sg, S, np.random.Generator(np.random.PCG64(seed + 5)))
    return sg, S


def snow_filters(OL, sg, S):
    fn = sg.filters["nolakec"]
    a, b = np.zeros(len(fn), np.int32), np.zeros(len(fn), np.int32)
    na, nb = C.c_int32(), C.c_int32()
    OL.oracle_build_snow_filter(len(fn), abi.i32p(fn), abi.i32p(S["snl"]), sg.bounds.begc, abi.i32p(a), C.byref(na), abi.i32p(b), C.byref(nb))
    return a[:na.value].copy(), b[:nb.value].copy()


def run_snow_water(OL, prm, sg, S, fs, fns, bounds=None):
    st = abi.Status()
    f = abi.make_struct("snowwater", S, sg.bounds)
    z = np.zeros(1, np.int32)
    rc = OL.oracle_snow_water(C.byref(prm), C.byref(bounds if bounds is not None else sg.bounds), len(fs), abi.i32p(fs if len(fs) else z),
                              len(fns), abi.i32p(fns if len(fns) else z), C.byref(f), C.byref(st))
    return rc, st


def run_snow_layers(OL, prm, sg, S, fs, bounds=None):
    st = abi.Status()
    f = abi.make_struct("snowlayers", S, sg.bounds)
    z = np.zeros(1, np.int32)
    rc = OL.oracle_snow_layers(C.byref(prm), C.byref(bounds if bounds is not None else sg.bounds), len(fs), abi.i32p(fs if len(fs) else z),
                               C.byref(f), C.byref(st))
    return rc, st


def dz_limits(OL, prm):
    a, b, c = (np.zeros(NSNO) for _ in range(3))
    dp = C.POINTER(C.c_double)
    OL.oracle_snow_dz_limits(C.byref(prm), a.ctypes.data_as(dp), b.ctypes.data_as(dp), c.ctypes.data_as(dp))
    return a, b, c


def test_build_snow_filter(oracle_lib):
    sg, S = case(300, 811)
    fs, fns = snow_filters(oracle_lib, sg, S)
    fn = sg.filters["nolakec"]
    assert np.array_equal(fs, fn[S["snl"][fn - 1] < 0]) and np.array_equal(fns, fn[S["snl"][fn - 1] >= 0])
    assert len(fs) > 50 and len(fns) > 50


def test_dz_limits_follow_the_namelist(oracle_lib):
    """InitSnowLayers, SnowHydrologyMod.F90:2985-3002, for nlevsno = 12 and the clm5 / clm6 namelist values"""
    dzmin, dzmax_l, dzmax_u = dz_limits(oracle_lib, abi.default_params())
    u = [0.02, 0.05]
    for j in range(2, 12):
        u.append(2 * u[-1] + 0.01)
    assert np.allclose(dzmax_u[:11], u[:11], rtol=0, atol=1e-15) and dzmax_u[11] > 1e300 and dzmax_l[11] > 1e300
    assert np.allclose(dzmin[:5], [0.010, 0.015, 0.025, 0.055, 0.115], rtol=0, atol=1e-15)
    assert np.allclose(dzmax_l[:4], [0.03, 0.07, 0.18, 0.41], rtol=0, atol=1e-15)


def test_snow_water_conserves_water_and_aerosol(oracle_lib):
    sg, S = case(800, 821)
    prm = abi.default_params()
    fs, fns = snow_filters(oracle_lib, sg, S)
    ref = copy_state(S)
    rc, st = run_snow_water(oracle_lib, prm, sg, ref, fs, fns)
    assert rc == 0, st.msg
    c = fs - 1
    dt = prm.dtime
    w0 = (S["h2osoi_ice"][:NSNO, c] + S["h2osoi_liq"][:NSNO, c]).sum(0)
    w1 = (ref["h2osoi_ice"][:NSNO, c] + ref["h2osoi_liq"][:NSNO, c]).sum(0)
    src = S["frac_sno_eff"][c] * (S["qflx_soliddew_to_top_layer"][c] - S["qflx_solidevap_from_top_layer"][c] + S["qflx_liq_grnd"][c]
                                  + S["qflx_liqdew_to_top_layer"][c] - S["qflx_liqevap_from_top_layer"][c]) * dt
    out = ref["qflx_snow_percolation"][NSNO - 1, c] * dt
    assert np.max(np.abs(w1 - (w0 + src - out)) / np.maximum(w0, 1.0)) < 1e-12       # (truncation to zero moves <= 1e-12 relative)
    assert (out > 0).sum() > 20 and (ref["qflx_snow_percolation"][:NSNO - 1, c] > 0).any()
    # aerosol: what leaves the pack is what the bottom layer loses; deposition adds forc_aer * dtime
    g = S["col_gridcell"][c] - 1
    dep = {"bcphi": S["forc_aer"][0, g] + S["forc_aer"][2, g], "bcpho": S["forc_aer"][1, g], "ocphi": S["forc_aer"][3, g] + S["forc_aer"][5, g],
           "ocpho": S["forc_aer"][4, g], "dst1": S["forc_aer"][6, g] + S["forc_aer"][7, g], "dst2": S["forc_aer"][8, g] + S["forc_aer"][9, g],
           "dst3": S["forc_aer"][10, g] + S["forc_aer"][11, g], "dst4": S["forc_aer"][12, g] + S["forc_aer"][13, g]}
    for a in AER:
        m0, m1 = S["mss_" + a][:, c].sum(0), ref["mss_" + a][:, c].sum(0)
        assert np.all(m1 <= m0 + dep[a] * dt + 1e-18) and np.all(ref["mss_" + a][:, c] >= 0.0)
        nodrain = out == 0
        assert np.max(np.abs(m1[nodrain] - m0[nodrain] - dep[a][nodrain] * dt) / m0[nodrain]) < 1e-12
    # layers never thinner than their water (PostPercolation_AdjustLayerThicknesses), no-snow columns reset
    act = np.arange(-NSNO + 1, 1)[:, None] >= (S["snl"][c] + 1)[None, :]
    need = ref["h2osoi_liq"][:NSNO, c] / 1000.0 + ref["h2osoi_ice"][:NSNO, c] / 917.0
    assert np.all(ref["dz"][:NSNO, c][act] >= need[act])
    cn = fns - 1
    bare = S["h2osno_no_layers"][cn] <= 0
    assert bare.any() and np.all(ref["snow_depth"][cn][bare] == 0) and np.all(ref["int_snow"][cn][bare] == 0)
    assert np.array_equal(ref["qflx_rain_plus_snomelt"][cn], S["qflx_liq_grnd"][cn] + S["qflx_snomelt"][cn])
    # the exact-sublimation columns were truncated to zero, not left at rounding noise
    top = S["snl"][c] + NSNO
    gone = S["qflx_solidevap_from_top_layer"][c] * dt * S["frac_sno_eff"][c] == S["h2osoi_ice"][top, c]
    assert gone.sum() > 3


def test_snow_water_reports_negative_top_layer(oracle_lib):
    sg, S = case(200, 831)
    prm = abi.default_params()
    fs, fns = snow_filters(oracle_lib, sg, S)
    bad = fs[11]
    S["qflx_solidevap_from_top_layer"][bad - 1] = 1.0
    rc, st = run_snow_water(oracle_lib, prm, sg, copy_state(S), fs, fns)
    assert rc == 18 and st.subgrid_index == bad and b"h2osoi_ice has gone significantly negative" in st.msg


def snow_totals(S, c):
    """per-column totals over the snow pack (+ what CombineSnowLayers may hand to soil layer 1 / h2osno_no_layers)"""
    snl = S["snl"][c]
    act = np.arange(-NSNO + 1, 1)[:, None] >= (snl + 1)[None, :]
    ice, liq, t = S["h2osoi_ice"][:NSNO, c], S["h2osoi_liq"][:NSNO, c], S["t_soisno"][:NSNO, c]
    water = np.where(act, ice + liq, 0.0).sum(0)
    enth = np.where(act, (CPICE * ice + CPLIQ * liq) * (t - TFRZ) + HFUS * liq, 0.0).sum(0)
    depth = np.where(act, S["dz"][:NSNO, c], 0.0).sum(0)
    aer = {a: np.where(act, S["mss_" + a][:, c], 0.0).sum(0) for a in AER}
    return water, enth, depth, aer


@pytest.mark.parametrize("method,wind,subgrid", [(2, 1, 1), (1, 0, 0)], ids=["vionnet_wind_subgrid", "anderson_nowind_iceold"])
def test_snow_layers_invariants(oracle_lib, method, wind, subgrid):
    sg, S = case(1500, 841)
    prm = abi.default_params()
    prm.snow_overburden_compaction_method, prm.wind_dependent_snow_density, prm.use_subgrid_fluxes = method, wind, subgrid
    fs, _ = snow_filters(oracle_lib, sg, S)
    ref = copy_state(S)
    rc, st = run_snow_layers(oracle_lib, prm, sg, ref, fs)
    assert rc == 0, st.msg
    c = fs - 1
    dzmin, dzmax_l, dzmax_u = dz_limits(oracle_lib, prm)
    w0, e0, d0, a0 = snow_totals(S, c)
    w1, e1, d1, a1 = snow_totals(ref, c)
    soil = np.isin(S["lun_itype"][c], (1, 2))
    # water: pack + soil layer 1 + unlayered snow is conserved on soil / crop columns; elsewhere liquid of a vanishing pack is dropped
    tot0 = w0 + S["h2osoi_liq"][NSNO, c] + S["h2osoi_ice"][NSNO, c] + 0.0
    gone = (ref["snl"][c] == 0)
    tot1 = w1 + ref["h2osoi_liq"][NSNO, c] + ref["h2osoi_ice"][NSNO, c] + np.where(gone, ref["h2osno_no_layers"][c], 0.0)
    assert np.max(np.abs(tot1 - tot0)[soil] / tot0[soil]) < 1e-13
    assert gone.sum() > 5 and (ref["snl"][c] < S["snl"][c]).sum() > 50 and (ref["snl"][c] > S["snl"][c]).sum() > 50
    # layers are only ever merged or split: compaction shrinks the depth, nothing else changes it (frac_sno_eff weighting aside)
    keep = ~gone
    if not subgrid:                          # (with use_subgrid_fluxes the melt term may be an expansion: "allowing for negative values for ddz3", :2018)
        assert np.all(d1[keep] <= d0[keep] * (1 + 1e-12))
    # enthalpy: Combo conserves it; a thin layer merged without temperature adjustment (:2246-2250) and a split with the
    # temperature-gradient rule (:2705-2712) do not, so only columns whose layer count is unchanged are exact
    same = keep & (ref["snl"][c] == S["snl"][c]) & (np.abs(d1 - d0) <= 1e-12 * d0) & (np.abs(w1 - w0) <= 1e-13 * w0)
    if same.any():
        assert np.max(np.abs(e1 - e0)[same] / np.maximum(np.abs(e0[same]), 1.0)) < 1e-9
    # aerosol mass: conserved by every merge and split except the bottom layer's merge into the soil (:2233 moves it only for j < 0)
    act0 = np.arange(-NSNO + 1, 1)[:, None] >= (S["snl"][c] + 1)[None, :]
    nothin = keep & np.all(np.where(act0, S["h2osoi_ice"][:NSNO, c] > 0.01, True), axis=0)
    assert nothin.sum() > 100
    for a in AER:
        assert np.max(np.abs(a1[a] - a0[a])[nothin] / a0[a][nothin]) < 1e-12
        assert np.all(a1[a][gone] == 0.0)
    # structure after the update: layers above snl+1 are zeroed, interfaces stack up from zi(0) = 0
    snl1 = ref["snl"][c]
    lev = np.arange(-NSNO + 1, 1)[:, None]
    empty = (lev <= snl1[None, :]) & (snl1[None, :] > -NSNO)
    for k in ("h2osoi_ice", "h2osoi_liq", "t_soisno", "dz", "z"):
        assert np.all(ref[k][:NSNO, c][empty] == 0.0), k
    act = lev >= (snl1 + 1)[None, :]
    zi = ref["zi"][:, c]                                     # rows: levels -12 .. 25
    assert np.allclose((zi[1:NSNO + 1] - zi[0:NSNO])[act], ref["dz"][:NSNO, c][act], rtol=1e-12, atol=1e-15)
    # thickness limits (in frac_sno_eff-weighted thickness, as DivideSnowLayers measures them): after the update no layer with
    # layers beneath exceeds dzmax_u, the bottom layer does not exceed dzmax_l
    fse = ref["frac_sno_eff"][c]
    for i, cc in enumerate(c):
        n = -snl1[i]
        if n == 0:
            continue
        d = ref["dz"][NSNO - n:NSNO, cc] * fse[i]
        assert np.all(d[:-1] <= dzmax_u[:n - 1] * (1 + 1e-12)), (cc, d)
        if n < NSNO:
            assert d[-1] <= dzmax_l[n - 1] * (1 + 1e-12), (cc, d)
    assert np.all(ref["snw_rds"][:, c][act] >= prm.snw_rds_min - 1e-9) and np.all(ref["snw_rds"][:, c][act] <= 1500.0)


def test_snow_layers_refuses_lake_and_urban(oracle_lib):
    sg, S = case(200, 851)
    prm = abi.default_params()
    fs, _ = snow_filters(oracle_lib, sg, S)
    T = copy_state(S)
    T["lun_itype"][fs[3] - 1] = 8
    rc, st = run_snow_layers(oracle_lib, prm, sg, T, fs)
    assert rc == 16 and st.subgrid_index == fs[3]
    T = copy_state(S)
    T["lun_itype"][fs[4] - 1] = 5
    rc, st = run_snow_layers(oracle_lib, prm, sg, T, fs)
    assert rc == 2
    T = copy_state(S)
    rc, st = run_snow_layers(oracle_lib, prm, sg, T, fs[:0])
    assert rc == 0
    for k in S:
        assert np.array_equal(T[k], S[k], equal_nan=True), k


SNO = ("swe_old", "snw_rds") + tuple("mss_" + a for a in AER)
SNOSOI = ("h2osoi_ice", "h2osoi_liq", "dz", "z", "t_soisno", "imelt", "frac_iceold")


def test_snow_water_matches_python_restatement(oracle_lib):
    """oracle_snow_water against tests/snow_python.py (written from the Fortran), column by column: identical bits"""
    from tests import snow_python as sp
    sg, S = case(500, 861)
    prm = abi.default_params()
    fs, fns = snow_filters(oracle_lib, sg, S)
    ref = copy_state(S)
    rc, st = run_snow_water(oracle_lib, prm, sg, ref, fs, fns)
    assert rc == 0
    scal = ("snl", "frac_sno_eff", "qflx_soliddew_to_top_layer", "qflx_solidevap_from_top_layer", "qflx_liq_grnd", "qflx_liqdew_to_top_layer",
            "qflx_liqevap_from_top_layer", "int_snow", "qflx_snow_drain")
    for c1 in fs:
        c = c1 - 1
        col = sp.column(S, c, tuple("mss_" + a for a in AER), ("h2osoi_ice", "h2osoi_liq", "dz"), scal)
        sp.snow_water_column(prm, col, S["forc_aer"][:, S["col_gridcell"][c] - 1].tolist())
        for k in ("h2osoi_ice", "h2osoi_liq", "dz", "qflx_snow_percolation") + tuple("mss_" + a for a in AER):
            assert np.array_equal(np.array(col[k].v), ref[k][:, c]), (c1, k)
        for k in ("int_snow", "qflx_snow_drain", "qflx_rain_plus_snomelt"):
            assert col[k] == ref[k][c], (c1, k)
    g = S["col_gridcell"] - 1                                # AerosolFluxes diagnostics, every column (AerosolMod.F90:728-750)
    assert np.array_equal(ref["flx_bc_dep"], S["forc_aer"][0, g] + S["forc_aer"][1, g] + S["forc_aer"][2, g])
    assert np.array_equal(ref["flx_oc_dep_phi"], S["forc_aer"][3, g] + S["forc_aer"][5, g])
    assert np.array_equal(ref["flx_dst_dep_dry3"], S["forc_aer"][11, g]) and np.array_equal(ref["flx_dst_dep_wet1"], S["forc_aer"][6, g])
    tot = S["forc_aer"][6, g]
    for k in range(7, 14):
        tot = tot + S["forc_aer"][k, g]
    assert np.array_equal(ref["flx_dst_dep"], tot)


@pytest.mark.parametrize("method,wind,subgrid", [(2, 1, 1), (1, 0, 0)], ids=["vionnet_wind_subgrid", "anderson_nowind_iceold"])
def test_snow_layers_match_python_restatement(oracle_lib, method, wind, subgrid):
    """oracle_snow_layers against tests/snow_python.py, column by column: identical bits in every element of every array,
    including the elements a shift leaves behind above the pack (aerosol masses, grain radii)"""
    from tests import snow_python as sp
    sg, S = case(700, 871)
    prm = abi.default_params()
    prm.snow_overburden_compaction_method, prm.wind_dependent_snow_density, prm.use_subgrid_fluxes = method, wind, subgrid
    fs, _ = snow_filters(oracle_lib, sg, S)
    ref = copy_state(S)
    rc, st = run_snow_layers(oracle_lib, prm, sg, ref, fs)
    assert rc == 0
    scal = ("snl", "lun_itype", "frac_sno_eff", "frac_sno", "frac_h2osfc", "int_snow", "n_melt", "snow_depth", "h2osno_no_layers")
    nmerge = nsplit = 0
    for c1 in fs:
        c = c1 - 1
        col = sp.column(S, c, SNO, SNOSOI, scal)
        zi = sp.Lev(-NSNO, S["zi"][:, c].tolist())
        sp.snow_layers_column(prm, col, zi, float(S["forc_wind"][S["col_gridcell"][c] - 1]))
        assert col["snl"] == ref["snl"][c], c1
        for k in ("h2osoi_ice", "h2osoi_liq", "dz", "z", "t_soisno", "snw_rds") + tuple("mss_" + a for a in AER):
            assert np.array_equal(np.array(col[k].v), ref[k][:, c]), (c1, k, np.array(col[k].v)[:13], ref[k][:13, c])
        assert np.array_equal(np.array(zi.v), ref["zi"][:, c]), c1
        for k in ("frac_sno_eff", "frac_sno", "int_snow", "snow_depth", "h2osno_no_layers", "qflx_sl_top_soil"):
            assert col[k] == ref[k][c], (c1, k)
        nmerge += col["snl"] > S["snl"][c]
        nsplit += col["snl"] < S["snl"][c]
    assert nmerge > 30 and nsplit > 30


# ----------------------------------------------------------------------------------------------------------------------
# SnowCapping
def capping_case(n=600, seed=881):
    sg, S = case(n, seed)
    rng = np.random.Generator(np.random.PCG64(seed + 9))
    nc = sg.ncol
    S["topo"] = rng.uniform(0.0, 3000.0, nc)
    deep = np.nonzero((S["snl"] < 0) & (rng.random(nc) < 0.15))[0]          # packs above h2osno_max: a very heavy bottom layer
    S["h2osoi_ice"][NSNO - 1, deep] = rng.uniform(9000.0, 14000.0, len(deep))
    S["dz"][NSNO - 1, deep] = S["h2osoi_ice"][NSNO - 1, deep] / rng.uniform(300.0, 800.0, len(deep))
    thin = deep[::4]                                                        # ... and some whose excess exceeds the bottom layer
    S["h2osoi_ice"][NSNO - 1, thin] = 40.0
    S["h2osoi_ice"][NSNO - 2, thin] = 12000.0
    S["snl"][thin] = np.minimum(S["snl"][thin], -2)
    for k in ("qflx_snwcp_ice", "qflx_snwcp_liq", "qflx_snwcp_discarded_ice", "qflx_snwcp_discarded_liq"):
        S[k] = np.full(nc, 1.0e36)
    return sg, S


def run_snow_capping(OL, prm, sg, S, fi, fs, nstep=100, bounds=None):
    st = abi.Status()
    f = abi.make_struct("snowcapping", S, sg.bounds)
    z = np.zeros(1, np.int32)
    rc = OL.oracle_snow_capping(C.byref(prm), C.byref(bounds if bounds is not None else sg.bounds), len(fi), abi.i32p(fi if len(fi) else z),
                                len(fs), abi.i32p(fs if len(fs) else z), C.byref(f), nstep, C.byref(st))
    return rc, st


def capping_np(prm, sg, S0, fi, fs, nstep):
    """SnowHydrologyMod.F90:3121-3693 in NumPy (bulk water)"""
    S = copy_state(S0)
    ci, c = fi - 1, fs - 1
    for k in ("qflx_snwcp_ice", "qflx_snwcp_liq", "qflx_snwcp_discarded_ice", "qflx_snwcp_discarded_liq"):
        S[k][ci] = 0.0
    tot = S["h2osno_no_layers"][c].copy()
    for j in range(NSNO):                                       # CalculateTotalH2osno: top to bottom over the active layers
        on = (j - NSNO + 1) >= S["snl"][c] + 1
        tot = np.where(on, tot + S["h2osoi_ice"][j, c] + S["h2osoi_liq"][j, c], tot)
    excess = np.where(tot > prm.h2osno_max, tot - prm.h2osno_max, 0.0)
    runoff = tot > prm.h2osno_max
    if (prm.reset_snow or prm.reset_snow_glc) and nstep <= 4 * NSNO:
        ice_lu = S["lun_itype"][c] == 4
        r1 = ~ice_lu & bool(prm.reset_snow) & (tot > 35.0)
        r2 = ice_lu & bool(prm.reset_snow_glc) & (tot > 35.0) & (S["topo"][c] <= prm.reset_snow_glc_ela)
        excess = np.where(r1 | r2, tot - 35.0, excess)
        runoff = np.where(r1 | r2, False, runoff)
    cap = excess > 0.0
    cc, runoff, excess = c[cap], runoff[cap], excess[cap]
    b = NSNO - 1
    ice, liq = S["h2osoi_ice"][b, cc], S["h2osoi_liq"][b, cc]
    rho = ice / S["dz"][b, cc]
    m = ice + liq
    take = np.minimum(excess, m * (1.0 - 1.e-3))
    icefrac = ice / m
    fi_, fl_ = take / prm.dtime * icefrac, take / prm.dtime * (1.0 - icefrac)
    S["qflx_snwcp_ice"][cc] = np.where(runoff, fi_, 0.0)
    S["qflx_snwcp_liq"][cc] = np.where(runoff, fl_, 0.0)
    S["qflx_snwcp_discarded_ice"][cc] = np.where(runoff, 0.0, fi_)
    S["qflx_snwcp_discarded_liq"][cc] = np.where(runoff, 0.0, fl_)
    adj = (m - take) / m
    S["h2osoi_ice"][b, cc] = ice - (S["qflx_snwcp_ice"][cc] + S["qflx_snwcp_discarded_ice"][cc]) * prm.dtime
    S["h2osoi_liq"][b, cc] = liq - (S["qflx_snwcp_liq"][cc] + S["qflx_snwcp_discarded_liq"][cc]) * prm.dtime
    S["dz"][b, cc] = np.where(rho > 1.0, S["h2osoi_ice"][b, cc] / rho, S["dz"][b, cc])
    for a in AER:
        S["mss_" + a][b, cc] = S["mss_" + a][b, cc] * adj
    return S, cap.sum()


@pytest.mark.parametrize("reset,reset_glc,nstep", [(0, 0, 100), (1, 1, 10), (1, 0, 100)], ids=["capping", "reset_active", "reset_expired"])
def test_snow_capping_matches_numpy(oracle_lib, reset, reset_glc, nstep):
    sg, S = capping_case()
    prm = abi.default_params()
    prm.reset_snow, prm.reset_snow_glc, prm.reset_snow_glc_ela = reset, reset_glc, 1500.0
    fs, _ = snow_filters(oracle_lib, sg, S)
    fi = sg.filters["nolakec"]
    ref = copy_state(S)
    rc, st = run_snow_capping(oracle_lib, prm, sg, ref, fi, fs, nstep)
    assert rc == 0, st.msg
    exp, ncap = capping_np(prm, sg, S, fi, fs, nstep)
    for f in abi.FIELDS["snowcapping"]:
        assert np.array_equal(ref[f.name], exp[f.name], equal_nan=True), f.name           # no transcendentals: identical bits
    assert ncap > (200 if (reset and nstep <= 48) else 10)
    c = fs - 1
    # what leaves the pack is what the fluxes carry
    w0 = (S["h2osoi_ice"][:NSNO, c] + S["h2osoi_liq"][:NSNO, c]).sum(0)
    w1 = (ref["h2osoi_ice"][:NSNO, c] + ref["h2osoi_liq"][:NSNO, c]).sum(0)
    out = (ref["qflx_snwcp_ice"][c] + ref["qflx_snwcp_liq"][c] + ref["qflx_snwcp_discarded_ice"][c] + ref["qflx_snwcp_discarded_liq"][c]) * prm.dtime
    assert np.max(np.abs(w0 - w1 - out) / w0) < 1e-12
    if not reset:
        assert np.all(ref["qflx_snwcp_discarded_ice"][c] == 0.0) and (ref["qflx_snwcp_ice"][c] > 0).sum() == ncap
