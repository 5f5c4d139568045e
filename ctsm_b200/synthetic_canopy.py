"""Synthetic CanopyFluxes + PHS inputs on the subgrid and
soil/snow state built by ctsm_b200.synthetic.

Parameter values marked "external" in SURVEY.md Appendix D are synthetic choices, not CTSM
parameter-file values (the production parameter file is not part of the reference checkout).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import abi
from .abi import NLEVSNO, NLEVGRND, NLEVSOI, ISTDLAK
from .synthetic import (Subgrid, vertical_grid, build_subgrid, soil_state, GRID_SIZES, TFRZ, DENH2O)

N_PFT_TABLE = abi.MXPFT + 1


def _qsat_np(T, p):
    """Vectorised QSat (QSatMod.F90:61-127); used only to build physically consistent synthetic forcing."""
    a = [6.11213476, 0.444007856, 0.143064234e-01, 0.264461437e-03, 0.305903558e-05, 0.196237241e-07,
         0.892344772e-10, -0.373208410e-12, 0.209339997e-15]
    c = [6.11123516, 0.503109514, 0.188369801e-01, 0.420547422e-03, 0.614396778e-05, 0.602780717e-07,
         0.387940929e-09, 0.149436277e-11, 0.262655803e-14]
    td = np.clip(T - TFRZ, -75.0, 100.0)
    ew = np.polyval(a[::-1], td)
    ei = np.polyval(c[::-1], td)
    es = np.where(td >= 0, ew, ei) * 100.0
    return 0.622 * es / (p - 0.378 * es), es


def pft_tables() -> Dict[str, np.ndarray]:
    """Per-PFT parameter tables (index 0:mxpft).  psi50/ck follow PhotosynthesisMod.F90:929-934
    (setParamsForTesting), root_radius/root_density the code comments at :2948-2949, dleaf the tech note;
    everything else is a SYNTHETIC, physically plausible choice."""
    n = N_PFT_TABLE
    T = {}
    tree = np.zeros(n, dtype=np.int32); tree[1:9] = 1
    shrub = np.zeros(n, dtype=np.int32); shrub[9:12] = 1
    c3 = np.ones(n); c3[14] = 0.0
    T["pft_is_tree"], T["pft_is_shrub"], T["pft_c3psn"] = tree, shrub, c3
    T["pft_crop"] = np.zeros(n)

    def per(vals, default):
        a = np.full(n, float(default))
        a[:len(vals)] = vals
        return a
    T["pft_slatop"] = per([0.01, 0.010, 0.008, 0.024, 0.012, 0.012, 0.030, 0.030, 0.030, 0.012, 0.030, 0.030,
                           0.030, 0.030, 0.030], 0.03)
    T["pft_leafcn"] = per([25, 58, 58, 26, 30, 30, 23, 23, 23, 36, 23, 23, 28, 28, 35], 25)
    T["pft_flnr"] = per([0.1, 0.051, 0.047, 0.055, 0.076, 0.076, 0.106, 0.106, 0.106, 0.033, 0.106, 0.106,
                         0.137, 0.137, 0.090], 0.1)
    T["pft_fnitr"] = per([1.0, 0.72, 0.78, 0.79, 0.83, 0.71, 0.66, 0.64, 0.70, 0.62, 0.60, 0.76, 0.68, 0.61, 0.64], 0.7)
    T["pft_dleaf"] = np.full(n, 0.04)
    T["pft_dbh"] = per([0.0, 0.30, 0.25, 0.25, 0.45, 0.35, 0.40, 0.30, 0.20, 0.04, 0.06, 0.03, 0.0, 0.0, 0.0], 0.0)
    T["pft_fbw"] = np.full(n, 0.5)
    T["pft_nstem"] = per([0.0, 0.08, 0.10, 0.10, 0.05, 0.06, 0.05, 0.08, 0.10, 0.5, 0.5, 0.5, 0.0, 0.0, 0.0], 0.0)
    T["pft_rstem_per_dbh"] = np.full(n, 100.0)
    T["pft_wood_density"] = per([0.0, 450, 450, 500, 600, 600, 600, 550, 500, 500, 500, 500, 0, 0, 0], 0.0)
    T["pft_z0v_Cr"] = np.full(n, 0.35)
    T["pft_z0v_Cs"] = np.full(n, 0.01)
    T["pft_z0v_c"] = np.full(n, 0.09)
    T["pft_z0v_cw"] = per([4.0] * 12 + [8.0, 8.0, 8.0], 4.0)
    T["pft_z0v_LAImax"] = np.full(n, 6.0)
    T["pft_smpso"] = per([-66000.0] * 12 + [-74000.0] * 3, -74000.0)
    T["pft_smpsc"] = per([-255000.0] * 12 + [-275000.0] * 3, -275000.0)
    T["pft_froot_leaf"] = per([0.0, 1.5, 1.5, 1.5, 1.0, 1.0, 1.0, 1.2, 1.2, 1.5, 1.5, 1.5, 2.0, 2.0, 2.0], 1.0)
    T["pft_root_radius"] = np.full(n, 0.29e-3)
    T["pft_root_density"] = np.full(n, 0.31e6)
    T["pft_mbbopt"] = np.where(c3 > 0.5, 9.0, 4.0)
    T["pft_medlynintercept"] = np.full(n, 100.0)
    T["pft_medlynslope"] = per([2.0, 2.35, 2.35, 2.35, 4.12, 4.12, 4.45, 4.45, 4.45, 4.70, 4.70, 4.70,
                                2.22, 5.25, 1.62], 4.0)
    T["pft_krmax"] = per([1e-9, 2e-9, 2e-9, 2e-9, 4e-9, 3e-9, 4e-9, 3e-9, 2e-9, 2e-9, 2e-9, 2e-9, 3e-9, 3e-9, 3e-9], 2e-9)
    T["pft_theta_cj"] = np.where(c3 > 0.5, 0.9393, 0.80)
    T["pft_kmax"] = np.full((4, n), 2.0e-8)
    psi = np.full(n, -340000.0); psi[1] = -150000.0; psi[2] = -530000.0; psi[3:13] = -400000.0; psi[0] = -150000.0
    T["pft_psi50"] = np.repeat(psi[None, :], 4, axis=0).copy()
    T["pft_ck"] = np.full((4, n), 3.95)
    return T


def canopy_state(sg: Subgrid, S: Dict[str, np.ndarray], rng: np.random.Generator, day_fraction: float = 0.5) -> None:
    """Adds the fields of group `canopyfluxes` (ctsm_b200_fields.def) to S, consistent with the soil/snow state
    already in S, and the exposedvegp / noexposedvegp filters to sg.filters."""
    nc, npch, ng = sg.ncol, sg.npatch, sg.ngrc
    g = lambda a, b, *sh: rng.uniform(a, b, size=sh)
    lo = -NLEVSNO + 1
    dzsoi, zisoi, zsoi = vertical_grid()
    S.update(pft_tables())
    # --- gridcell
    S["dayl"] = g(30000.0, 50000.0, ng)
    S["max_dayl"] = np.maximum(S["dayl"], g(45000.0, 55000.0, ng))
    spd, ang = g(0.5, 12.0, ng), g(0.0, 2 * np.pi, ng)
    S["forc_u"], S["forc_v"] = spd * np.cos(ang), spd * np.sin(ang)
    S["forc_pco2"] = np.full(ng, 40.0) * g(0.95, 1.05, ng)
    S["forc_po2"] = np.full(ng, 0.209 * 1.0e5)
    for nm in ("forc_hgt_t", "forc_hgt_u", "forc_hgt_q"):
        S[nm] = np.full(ng, 30.0)
    S["near_local_noon"] = (rng.random(ng) < 1.0 / 12.0).astype(np.int32)
    S["local_time_lt_noon"] = (rng.random(ng) < 0.5).astype(np.int32)
    is_day_g = rng.random(ng) < day_fraction
    # --- column
    t1 = S["t_soisno"][1 - lo]
    tg = S["t_grnd"]
    forc_t = np.clip(tg + rng.normal(0.0, 3.0, nc), 235.0, 318.0)
    pbot = g(85000.0, 101325.0, nc)
    S["forc_pbot"] = pbot
    S["forc_th"] = forc_t * (100000.0 / pbot) ** 0.286
    S["forc_rho"] = pbot / (287.04 * forc_t)
    qs_a, _ = _qsat_np(forc_t, pbot)
    S["forc_q"] = g(0.2, 0.95, nc) * qs_a
    S["thv"] = S["forc_th"] * (1.0 + 0.61 * S["forc_q"])
    # forc_lwrad, emg, htvp, t_h2osfc, t_grnd, frac_*, snow_depth, snl are already in S (soil_state)
    S["soilresis"] = 10.0 ** g(1.0, 3.3, nc)
    S["soilbeta"] = g(0.1, 1.0, nc)
    S["z0mg"] = np.where(S["frac_sno_eff"] > 0, 0.0024, 0.01)
    topsn = np.take_along_axis(S["t_soisno"], (S["snl"] + 1 - lo)[None, :], axis=0)[0]
    qgs, _ = _qsat_np(topsn, pbot); qgsoil, _ = _qsat_np(t1, pbot); qgh, _ = _qsat_np(S["t_h2osfc"], pbot)
    hr = g(0.5, 1.0, nc)
    S["qg_snow"], S["qg_soil"], S["qg_h2osfc"] = qgs, hr * qgsoil, qgh
    fs, fh = S["frac_sno_eff"], S["frac_h2osfc"]
    S["qg"] = fs * S["qg_snow"] + (1 - fs - fh) * S["qg_soil"] + fh * S["qg_h2osfc"]
    qg1, _ = _qsat_np(tg + 0.5, pbot); qg0, _ = _qsat_np(tg - 0.5, pbot)
    S["dqgdT"] = hr * (qg1 - qg0)
    # soil hydraulic state as SoilWater left it (CH78): s = liq/(dz*1000*watsat)
    soil = slice(NLEVSNO, NLEVSNO + NLEVGRND)
    s_l = np.clip(S["h2osoi_liq"][soil] / (S["dz"][soil] * DENH2O * S["watsat"]), 0.01, 1.0)
    S["smp_l"] = np.maximum(-S["sucsat"] * s_l ** (-S["bsw"]), -1.0e8)
    S["hk_l"] = S["hksat"] * s_l ** (2.0 * S["bsw"] + 3.0)
    S["h2osoi_liqvol"] = np.full((NLEVSNO + NLEVGRND, nc), 1.0e36)
    # --- patch
    pc = sg.patch_column - 1
    pg = sg.patch_gridcell - 1
    ivt = sg.patch_itype
    S["gridcell"] = sg.patch_gridcell.copy()
    S["itype"] = ivt.copy()
    S["patch_lakpoi"] = (sg.col_lun_itype[pc] == ISTDLAK).astype(np.int32)
    S["nrad"] = np.ones(npch, dtype=np.int32)
    is_day = is_day_g[pg]
    elai = g(0.1, 6.0, npch); esai = g(0.1, 1.0, npch)
    S["elai"], S["esai"] = elai, esai
    S["tlai"], S["tsai"] = elai * g(1.0, 1.2, npch), esai * g(1.0, 1.2, npch)
    tree, shrub = S["pft_is_tree"][ivt] > 0, S["pft_is_shrub"][ivt] > 0
    S["htop"] = np.where(tree, g(5.0, 35.0, npch), np.where(shrub, g(0.3, 2.0, npch), g(0.1, 1.0, npch)))
    fsun = (1.0 - np.exp(-0.5 * elai)) / (0.5 * elai)
    tiny_sun = rng.random(npch) < 0.02          # exercises the laisun ~ 0 (3x3) branch of calcstress
    laisun = np.where(is_day, np.where(tiny_sun, 5.0e-4, fsun * elai), 0.0)
    S["laisun"], S["laisha"] = laisun, elai - laisun
    S["laisun_z"], S["laisha_z"] = S["laisun"][None, :].copy(), S["laisha"][None, :].copy()
    S["tlai_z"] = S["tlai"][None, :].copy()
    S["parsun_z"] = np.where(is_day, g(20.0, 300.0, npch), 0.0)[None, :]
    S["parsha_z"] = np.where(is_day, g(5.0, 60.0, npch), 0.0)[None, :]
    S["sabv"] = np.where(is_day, g(20.0, 400.0, npch), 0.0)
    S["emv"] = 1.0 - np.exp(-(elai + esai))
    fwet = np.where(rng.random(npch) < 0.5, g(0.0, 0.3, npch), 0.0)
    S["fwet"], S["fdry"] = fwet, (1.0 - fwet) * elai / (elai + esai)
    ft_p = forc_t[pc]
    S["t_veg"] = ft_p + rng.normal(0.0, 2.0, npch)
    S["t_stem"] = ft_p + rng.normal(0.0, 2.0, npch)
    S["t_a10"] = ft_p + rng.normal(0.0, 3.0, npch)
    S["vcmaxcintsun"] = np.where(is_day, g(0.3, 2.0, npch), 0.0)
    S["vcmaxcintsha"] = g(0.3, 3.0, npch)
    for nm in ("o3coefvsun", "o3coefgsun", "o3coefvsha", "o3coefgsha"):
        S[nm] = np.ones(npch)
    S["froot_carbon"] = g(50.0, 300.0, npch)
    S["vcmx25_z"] = g(20.0, 80.0, npch)[None, :]
    S["jmx25_z"] = S["vcmx25_z"] * g(1.6, 2.0, npch)[None, :]
    beta_r = g(0.5, 2.5, npch)
    rf = np.exp(-zsoi[1:NLEVGRND + 1][:, None] / beta_r[None, :])
    rf[NLEVSOI:] = 0.0
    S["rootfr"] = rf / rf.sum(0, keepdims=True)
    S["displa"] = 0.67 * S["htop"]
    S["z0mv"] = 0.1 * S["htop"]
    S["thm"] = ft_p + 0.0098 * (30.0 + S["z0mv"] + S["displa"])
    S["snocan"] = np.where(rng.random(npch) < 0.2, g(0.0, 2.0, npch), 0.0)
    S["liqcan"] = np.where(rng.random(npch) < 0.3, g(0.0, 0.5, npch), 0.0)
    S["cgrnds"], S["cgrndl"] = np.zeros(npch), np.zeros(npch)
    S["qflx_tran_veg"] = g(0.0, 5.0e-5, npch)
    wroot = -g(2.0e4, 1.0e5, npch)
    wxyl = wroot - g(0.0, 2.0e4, npch) - 1000.0 * S["htop"]
    S["vegwp"] = np.stack([wxyl - g(0.0, 2.0e4, npch), wxyl - g(0.0, 2.0e4, npch), wxyl, wroot])
    # --- LUNA daily accumulators (Acc24_Climate_LUNA): some patches on their first day (spval: nothing accumulates),
    # the rest part-way through a day.  Own generator: the stream of `rng` stays what it was before these fields existed.
    r2 = np.random.Generator(np.random.PCG64(977 + 31 * npch))
    first_day = r2.random(npch) < 0.15
    nday = r2.integers(0, 24, npch).astype(np.int32)
    nnight = r2.integers(0, 24, npch).astype(np.int32)
    S["t_veg_day"] = np.where(first_day, 1.0e36, nday * r2.uniform(270.0, 300.0, npch))
    S["t_veg_night"] = np.where(first_day, 1.0e36, nnight * r2.uniform(265.0, 290.0, npch))
    S["ndaysteps"], S["nnightsteps"] = nday, nnight
    S["par24d_z"] = (nday * 1800.0 * r2.uniform(0.0, 150.0, npch))[None, :]
    S["par24x_z"] = r2.uniform(0.0, 300.0, npch)[None, :]
    S["fpsn24"] = nday * 1800.0 * r2.uniform(0.0, 10.0, npch)
    # --- outputs: recognisable fill so untouched elements are visible
    fill = 1.0e36
    for fs_ in abi.FIELDS["canopyfluxes"]:
        if fs_.name in S:
            continue
        n = sg.bounds.extent(fs_.sub)
        shape = (n,) if fs_.lev == "L1" else (fs_.nlev, n)
        S[fs_.name] = np.full(shape, fill) if fs_.ctype == "double" else np.full(shape, -9999, dtype=np.int32)
    for k, v in list(S.items()):
        S[k] = np.ascontiguousarray(v)
    # filters (filterMod.F90:595-648)
    nlu = sg.filters["nolakeurbanp"]
    ex = S["frac_veg_nosno"][nlu - 1] > 0
    sg.filters["exposedvegp"] = np.ascontiguousarray(nlu[ex])
    sg.filters["noexposedvegp"] = np.ascontiguousarray(nlu[~ex])


def make_full_case(size="tiny", seed: int = 20260101, special_every: int = 10, day_fraction: float = 0.5):
    """Soil/snow state + canopy state on one subgrid (configs 3 and 4)."""
    ngrc = GRID_SIZES[size] if isinstance(size, str) else int(size)
    rng = np.random.Generator(np.random.PCG64(seed))
    sg = build_subgrid(ngrc, rng, special_every)
    S = soil_state(sg, rng)
    canopy_state(sg, S, rng, day_fraction)
    return sg, S


def balance_state(sg: Subgrid, S: Dict[str, np.ndarray], rng: np.random.Generator, noise: float = 1.0e-11) -> None:
    """Adds the fields of groups `balancecheck` and `plantsink`: budgets that close up to `noise`
    (so the residuals are differences of O(1..100) operands, as in the model)."""
    nc, npch, ng = sg.ncol, sg.npatch, sg.ngrc
    g = lambda a, b, *sh: rng.uniform(a, b, size=sh)
    dtime = 1800.0
    first = np.searchsorted(sg.col_gridcell, np.arange(1, ng + 1), side="left")
    last = np.searchsorted(sg.col_gridcell, np.arange(1, ng + 1), side="right") - 1
    S["grc_coli"], S["grc_colf"] = (first + 1).astype(np.int32), (last + 1).astype(np.int32)
    S["col_active"] = sg.col_active.astype(np.int32)
    ncol_g = (last - first + 1)
    S["wtgcell"] = (1.0 / ncol_g)[sg.col_gridcell - 1]
    S["patch_active"] = sg.patch_active.astype(np.int32)
    S["npatches"] = (sg.col_patchf - sg.col_patchi + 1).astype(np.int32)
    colflux = ("forc_rain", "forc_snow", "qflx_flood", "qflx_sfc_irrig", "qflx_glcice_dyn_water_flux", "qflx_evap_tot",
               "qflx_surf", "qflx_qrgwl", "qflx_drain", "qflx_drain_perched", "qflx_ice_runoff",
               "qflx_snwcp_discarded_liq", "qflx_snwcp_discarded_ice")
    sign = (1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1, -1)
    net = np.zeros(nc)
    for nm, sgn in zip(colflux, sign):
        S[nm] = g(0.0, 3.0e-4, nc) * (rng.random(nc) < 0.6)
        net += sgn * S[nm]
    S["begwb"] = g(500.0, 3000.0, nc)
    S["endwb"] = S["begwb"] + net * dtime + rng.normal(0.0, noise, nc)
    grcflux = ("forc_rain_grc", "forc_snow_grc", "forc_flood_grc", "qflx_sfc_irrig_grc", "qflx_evap_tot_grc",
               "qflx_surf_grc", "qflx_qrgwl_grc", "qflx_drain_grc", "qflx_drain_perched_grc", "qflx_ice_runoff_grc")
    gsign = (1, 1, 1, 1, -1, -1, -1, -1, -1, -1)
    gnet = np.zeros(ng)
    for nm, sgn in zip(grcflux, gsign):
        S[nm] = g(0.0, 3.0e-4, ng)
        gnet += sgn * S[nm]
    agg = lambda a: np.add.reduceat(a * S["wtgcell"], first)
    gnet += agg(S["qflx_glcice_dyn_water_flux"]) - agg(S["qflx_snwcp_discarded_liq"]) - agg(S["qflx_snwcp_discarded_ice"])
    S["begwb_grc"] = g(500.0, 3000.0, ng)
    S["endwb_grc"] = S["begwb_grc"] + gnet * dtime + rng.normal(0.0, noise, ng)
    S["errh2o_grc"] = np.full(ng, 1.0e36)
    for nm in ("qflx_prec_grnd", "qflx_soliddew_to_top_layer", "qflx_liqdew_to_top_layer", "qflx_solidevap_from_top_layer",
               "qflx_liqevap_from_top_layer", "qflx_snwcp_ice", "qflx_snwcp_liq", "qflx_sl_top_soil",
               "qflx_snow_grnd", "qflx_liq_grnd", "qflx_snow_h2osfc"):
        S[nm] = g(0.0, 2.0e-4, nc) * (rng.random(nc) < 0.5)
    lo = -NLEVSNO + 1
    tot = S["h2osno_no_layers"].copy()
    for j in range(lo, 1):
        act = j >= S["snl"] + 1
        tot += np.where(act, S["h2osoi_ice"][j - lo] + S["h2osoi_liq"][j - lo], 0.0)
    fs = S["frac_sno_eff"]
    h2ice = np.where(np.abs(S["qflx_h2osfc_to_ice"]) < 1e30, S["qflx_h2osfc_to_ice"], 0.0)
    S["qflx_h2osfc_to_ice"] = h2ice
    S["qflx_snow_drain"] = np.where(np.abs(S["qflx_snow_drain"]) < 1e30, S["qflx_snow_drain"], 0.0)
    src = (S["qflx_snow_grnd"] - S["qflx_snow_h2osfc"]) + fs * (S["qflx_liq_grnd"] + S["qflx_soliddew_to_top_layer"]
                                                               + S["qflx_liqdew_to_top_layer"]) + h2ice
    snk = fs * (S["qflx_solidevap_from_top_layer"] + S["qflx_liqevap_from_top_layer"]) + S["qflx_snwcp_ice"] + S["qflx_snwcp_liq"] \
        + S["qflx_snwcp_discarded_ice"] + S["qflx_snwcp_discarded_liq"] + S["qflx_snow_drain"] + S["qflx_sl_top_soil"]
    lake = sg.col_lun_itype == ISTDLAK                       # BalanceCheckMod.F90:772-780
    src = np.where(lake, S["qflx_snow_grnd"] + fs * (S["qflx_liq_grnd"] + S["qflx_soliddew_to_top_layer"]
                                                     + S["qflx_liqdew_to_top_layer"]), src)
    S["h2osno_old"] = tot - (src - snk) * dtime + rng.normal(0.0, noise, nc)
    S["errsoi_col"] = rng.normal(0.0, 1.0e-7, nc)
    S["forc_solad"] = g(0.0, 300.0, 2, nc)
    S["forc_solai"] = g(0.0, 100.0, 2, ng)
    pc, pg = sg.patch_column - 1, sg.patch_gridcell - 1
    sol = S["forc_solad"][0, pc] + S["forc_solad"][1, pc] + S["forc_solai"][0, pg] + S["forc_solai"][1, pg]
    S["fsr"] = sol * g(0.1, 0.4, npch)
    S["fsa"] = sol - S["fsr"] + rng.normal(0.0, noise, npch)
    S["eflx_lwrad_out"] = g(250.0, 480.0, npch)
    S["eflx_lwrad_net"] = S["eflx_lwrad_out"] - S["forc_lwrad"][pc] + rng.normal(0.0, noise, npch)
    S["eflx_sh_tot"] = rng.normal(30.0, 60.0, npch)
    S["eflx_lh_tot"] = rng.normal(40.0, 60.0, npch)
    sabg_chk = np.where(np.abs(S["sabg_chk"]) < 1e30, S["sabg_chk"], 0.0)
    S["sabg_chk"] = sabg_chk
    dh = np.where(np.abs(S["dhsdt_canopy"]) < 1e30, S["dhsdt_canopy"], 0.0)
    S["dhsdt_canopy"] = dh
    S["eflx_soil_grnd"] = S["sabv"] + sabg_chk + S["forc_lwrad"][pc] - S["eflx_lwrad_out"] - S["eflx_sh_tot"] - S["eflx_lh_tot"] \
        - dh + rng.normal(0.0, noise, npch)
    for nm in ("errh2o", "errh2osno", "snow_sources", "snow_sinks", "qflx_phs_neg"):
        S[nm] = np.full(nc, 1.0e36)
    for nm in ("errsol", "errlon", "errseb", "netrad", "qflx_hydr_redist"):
        S[nm] = np.full(npch, 1.0e36)
    S["k_soil_root"] = np.where(np.abs(S["k_soil_root"]) < 1e30, S["k_soil_root"], 0.0)
    for k, v in list(S.items()):
        S[k] = np.ascontiguousarray(v)


def soilfluxes_state(sg: Subgrid, S: Dict[str, np.ndarray], rng: np.random.Generator) -> None:
    """Adds the fields of groups `soilfluxes` (SoilFluxesMod.F90:37) and `patch2col` (clm_driver.F90:1655).  t_ssbef / t_h2osfc_bef are the temperatures
    SoilTemperature saves on entry (SoilTemperatureMod.F90:290-300): copies of the current state, so a step that
    runs SoilTemperature first sees the true 'before' values.  Patch fluxes that CanopyFluxes leaves at spval on
    patches without exposed vegetation get the values BareGroundFluxes would give them (zero canopy fluxes)."""
    npch = sg.npatch
    g = lambda a, b, *sh: rng.uniform(a, b, size=sh)
    S["t_ssbef"] = S["t_soisno"].copy()
    S["t_h2osfc_bef"] = S["t_h2osfc"].copy()
    S["patch_active"] = sg.patch_active.astype(np.int32)
    fill = {"eflx_sh_veg": 0.0, "eflx_sh_stem": 0.0, "qflx_evap_veg": 0.0, "qflx_tran_veg": 0.0, "ulrad": 0.0}
    for nm, v in fill.items():
        S[nm] = np.where(np.abs(S[nm]) < 1e30, S[nm], v)
    for nm, (lo, hi) in {"dlrad": (0.0, 30.0), "cgrnds": (1.0, 12.0), "cgrndl": (1.0e-7, 4.0e-6), "t_skin": (250.0, 310.0)}.items():
        S[nm] = np.where(np.abs(S[nm]) < 1e30, S[nm], g(lo, hi, npch))
    for nm in ("xmf", "xmf_h2osfc", "c_h2osfc", "eflx_h2osfc_to_snow"):      # SoilTemperature outputs
        if not np.all(np.abs(S[nm]) < 1e30):
            S[nm] = np.where(np.abs(S[nm]) < 1e30, S[nm], 0.0 if nm != "c_h2osfc" else 1.0e-6)
    if not np.all(np.abs(S["fact"]) < 1e30):
        S["fact"] = np.where(np.abs(S["fact"]) < 1e30, S["fact"], g(1.0e-4, 5.0e-2, *S["fact"].shape))
    for fs in list(abi_fields("soilfluxes")) + list(abi_fields("patch2col")) + list(abi_fields("plantsinkdefault")):
        if fs.name not in S and fs.ctype == "double":       # (the topology integers come from balance_state)
            n = sg.ncol if fs.sub == "COL" else npch
            S[fs.name] = np.full(n if fs.lev == "L1" else (fs.nlev, n), 1.0e36, dtype=fs.dtype)
    for k, v in list(S.items()):
        S[k] = np.ascontiguousarray(v)


def bare_ground_state(sg: Subgrid, S: Dict[str, np.ndarray]) -> None:
    """What BareGroundFluxes leaves on patches without exposed vegetation (BareGroundFluxesMod.F90:294, :468): no effective
    roots, no transpiration.  The default root-water sink (use_hydrstress = .false.) reads both on every patch."""
    bare = np.ones(sg.npatch, dtype=bool)
    bare[sg.filters["exposedvegp"] - 1] = False
    S["rootr"][:, bare] = 0.0
    S["qflx_tran_veg"][bare] = 0.0


def preflux_state(sg: Subgrid, S: Dict[str, np.ndarray], rng: np.random.Generator) -> None:
    """Inputs of BiogeophysPreFluxCalcs / CalculateSurfaceHumidity / BareGroundFluxes that the rest of the step does not
    already carry: snow cover, air temperature, field capacity, the previous step's roughness
    lengths and the ZengWang2007 PFT ratios.  Outputs start from spval."""
    nc, npch = sg.ncol, sg.npatch
    g = lambda lo, hi, *shape: rng.uniform(lo, hi, shape)
    S["frac_sno"] = S["frac_sno_eff"].copy()                    # identical off urban landunits
    S["forc_t"] = S["forc_th"] - g(0.0, 0.4, nc)                # forc_th = forc_t*(psrf/pbot)**cappa, slightly warmer
    S["watfc"] = S["watsat"] * g(0.35, 0.8, *S["watsat"].shape)
    S["smpmin"] = np.full(nc, -1.0e8)
    S["z0m"] = g(0.01, 1.5, npch)                               # previous step (read where htop <= 1e-10)
    for nm in ("z0hg", "z0qg", "beta", "zii", "dsl", "soilalpha"):
        S[nm] = np.full(nc, 1.0e36)
    for nm in ("z0mg_p", "z0hg_p", "z0qg_p", "kbm1"):
        S[nm] = np.full(npch, 1.0e36)
    nt = len(S["pft_z0v_LAImax"])
    base = np.r_[0.0, rng.uniform(0.055, 0.12, abi_mod().MXPFT)]        # pftcon%z0mr / displar (noveg: 0)
    S["pft_z0mr"] = np.tile(base, nt // len(base))
    based = np.r_[0.0, rng.uniform(0.67, 0.68, abi_mod().MXPFT)]
    S["pft_displar"] = np.tile(based, nt // len(based))
    for grp in ("preflux", "surfacehumidity", "baregroundfluxes"):
        for fs in abi_fields(grp):
            if fs.name not in S and fs.ctype == "double":
                n = {"COL": nc, "PATCH": npch, "GRC": sg.ngrc}[fs.sub]
                S[fs.name] = np.full(n if fs.lev == "L1" else (fs.nlev, n), 1.0e36)
    for k, v in list(S.items()):
        S[k] = np.ascontiguousarray(v)


def waterbalance_state(sg: Subgrid, S: Dict[str, np.ndarray], rng: np.random.Generator) -> None:
    """Adds the fields of group `waterbalance` (BeginWaterColumnBalance, BalanceCheckMod.F90:171)."""
    nc = sg.ncol
    S["wa"] = rng.uniform(4000.0, 5000.0, nc)
    S["total_plant_stored_h2o"] = np.zeros(nc)
    S["excess_ice"] = np.where(rng.random((25, nc)) < 0.05, rng.uniform(0.0, 30.0, (25, nc)), 0.0)
    S["col_hydrologically_active"] = np.isin(sg.col_lun_itype, (1, 2)).astype(np.int32)
    S["patch_active"] = sg.patch_active.astype(np.int32)
    for nm in ("liqcan", "snocan"):
        S[nm] = np.where(np.abs(S[nm]) < 1e30, S[nm], 0.0)
    for fs in abi_fields("waterbalance"):
        if fs.name not in S:
            n = sg.ncol if fs.sub == "COL" else sg.npatch
            S[fs.name] = np.full(n if fs.lev == "L1" else (fs.nlev, n), 1.0e36, dtype=fs.dtype)
    for k, v in list(S.items()):
        S[k] = np.ascontiguousarray(v)


def watergrid_state(sg: Subgrid, S: Dict[str, np.ndarray], rng: np.random.Generator) -> None:
    """Adds the fields of group `watergridbalance` (WaterGridcellBalance, BalanceCheckMod.F90:132): lake geometry and ice
    fraction, dynbal baselines, gridcell -> column ranges and weights, what the dynbal dribblers still hold."""
    waterbalance_state(sg, S, rng)
    nc, ng = sg.ncol, sg.ngrc
    first = np.searchsorted(sg.col_gridcell, np.arange(1, ng + 1), side="left")
    last = np.searchsorted(sg.col_gridcell, np.arange(1, ng + 1), side="right") - 1
    S["grc_coli"], S["grc_colf"] = (first + 1).astype(np.int32), (last + 1).astype(np.int32)
    S["col_active"] = sg.col_active.astype(np.int32)
    w = rng.uniform(0.2, 1.0, nc)
    tot = np.add.reduceat(w, first)
    S["wtgcell"] = w / tot[sg.col_gridcell - 1]                    # column weights on the gridcell, summing to 1
    S["dynbal_baseline_liq"] = rng.uniform(0.0, 50.0, nc) * (rng.random(nc) < 0.3)
    S["dynbal_baseline_ice"] = rng.uniform(0.0, 20.0, nc) * (rng.random(nc) < 0.3)
    S["dz_lake"] = np.ascontiguousarray(np.broadcast_to(np.array([0.1, 1.0, 2.0, 3.0, 4.0, 5.0, 7.0, 7.0, 10.45, 10.45])[:, None], (10, nc)).copy())
    S["lake_icefrac"] = np.clip(rng.uniform(-0.5, 1.2, (10, nc)), 0.0, 1.0)
    S["qflx_liq_dynbal_left_to_dribble"] = rng.uniform(-5.0, 5.0, ng) * (rng.random(ng) < 0.2)
    S["qflx_ice_dynbal_left_to_dribble"] = rng.uniform(-1.0, 1.0, ng) * (rng.random(ng) < 0.2)
    S["begwb_grc"] = np.full(ng, 1.0e36)
    S["endwb_grc"] = np.full(ng, 1.0e36)
    for k in ("grc_coli", "grc_colf", "col_active", "wtgcell", "dynbal_baseline_liq", "dynbal_baseline_ice", "dz_lake",
              "lake_icefrac", "qflx_liq_dynbal_left_to_dribble", "qflx_ice_dynbal_left_to_dribble"):
        S[k] = np.ascontiguousarray(S[k])


def hydrology_state(sg: Subgrid, S: Dict[str, np.ndarray], rng: np.random.Generator) -> None:
    """Inputs of the surface-water / infiltration chain of HydrologyNoDrainage (HydrologyNoDrainageMod.F90:297-337; SURVEY.md 8f
    rank 3) that the rest of the step does not carry: flood water from the river model, water-table depths, surface-water
    threshold and fractions, the rain + snow-melt flux reaching the soil.  Every branch of the chain is populated: perched
    water table above / below the frost table, h2osfc above / below its threshold, h2osfc_partial driven negative (evaporation
    larger than the store), infiltration excess, columns with and without snow layers.  Outputs start from spval."""
    nc, ng = sg.ncol, sg.ngrc
    g = lambda lo, hi, *shape: rng.uniform(lo, hi, shape)
    S["col_gridcell"] = sg.col_gridcell.astype(np.int32)
    S["forc_flood"] = g(0.0, 2.0e-4, ng) * (rng.random(ng) < 0.2)
    S["topo_slope"] = g(0.05, 12.0, nc)                                  # degrees
    S["wtfact"] = g(0.1, 0.7, nc)
    S["zwt"] = g(0.3, 8.0, nc)
    S["zwt_perched"] = g(0.1, 6.0, nc)
    S["frost_table"] = g(0.1, 9.0, nc)
    S["h2osfc_thresh"] = g(0.5, 12.0, nc)
    S["h2osfc"] = np.where(rng.random(nc) < 0.5, g(0.0, 30.0, nc), 0.0)   # mm; half of the columns dry
    S["frac_h2osfc"] = np.where(S["h2osfc"] > 0.0, g(0.01, 0.9, nc), 0.0)
    S["frac_h2osfc_nosnow"] = np.minimum(1.0, S["frac_h2osfc"] * g(1.0, 1.3, nc))
    wet = rng.random(nc) < 0.6
    S["qflx_rain_plus_snomelt"] = np.where(wet, 10.0 ** g(-6.0, -2.3, nc), 0.0)     # up to ~18 mm/h: above qinmax on some columns
    S["qflx_snow_h2osfc"] = np.where(rng.random(nc) < 0.2, g(0.0, 2.0e-5, nc), 0.0)
    for nm in ("qflx_ev_soil_col", "qflx_ev_h2osfc_col", "qflx_liqevap_from_top_layer"):   # SoilFluxes outputs (spval before a step)
        S[nm] = np.where(np.abs(S[nm]) < 1e30, S[nm], g(-1.0e-5, 6.0e-5, nc))
    # evaporation that slightly over-empties the surface store (h2osfc_partial < 0, QflxH2osfcDrain's first branch): the deficit
    # stays below a few per cent of the store, a strongly negative qflx_infl would make SoilWater's adaptive step ill-conditioned
    big = (rng.random(nc) < 0.1) & (S["frac_h2osfc"] > 0.0)
    S["qflx_ev_h2osfc_col"] = np.where(big, g(1.0, 1.05, nc) * S["h2osfc"] / (1800.0 * np.maximum(S["frac_h2osfc"], 1e-3))
                                       + S["qflx_rain_plus_snomelt"] + 3.0e-4, S["qflx_ev_h2osfc_col"])
    for fs in abi_fields("infiltration"):
        if fs.name not in S and fs.ctype == "double":
            n = {"COL": nc, "PATCH": sg.npatch, "GRC": ng}[fs.sub]
            S[fs.name] = np.full(n if fs.lev == "L1" else (fs.nlev, n), 1.0e36)
    for k, v in list(S.items()):
        S[k] = np.ascontiguousarray(v)


def snow_state(sg: Subgrid, S: Dict[str, np.ndarray], rng: np.random.Generator) -> None:
    """Inputs of the snow routines of HydrologyNoDrainage (SnowWater, SnowCompaction, CombineSnowLayers, DivideSnowLayers;
    SURVEY.md 8f rank 3) that the rest of the step does not carry: aerosol masses and deposition, grain radii, the snow water
    before melt, melt flags, wind.  The layer structure soil_state draws (thickness 0.01 - 0.3 m in any order, density
    80 - 450 kg/m3) already violates the thickness limits in most columns, so combination and subdivision both run; on top of
    that: layers with almost no ice (<= 0.01 kg/m2: merged downwards, the bottom one into the soil), very light layers
    (< 50 kg/m3), one-layer packs thin enough to disappear, saturated (immobile) layers."""
    nc, ng = sg.ncol, sg.ngrc
    nsno = 12
    g = lambda lo, hi, *shape: rng.uniform(lo, hi, shape)
    snl = S["snl"]
    lev = np.arange(-nsno + 1, 1)[:, None]
    act = lev >= (snl + 1)[None, :]
    S["col_gridcell"] = sg.col_gridcell.astype(np.int32)
    ice, liq, dz = S["h2osoi_ice"], S["h2osoi_liq"], S["dz"]
    snowc = np.nonzero(snl < 0)[0]
    pick = lambda frac: snowc[rng.random(len(snowc)) < frac]
    for c in pick(0.12):                                      # an almost ice-free layer somewhere in the pack
        j = int(rng.integers(snl[c] + 1, 1))
        ice[j + nsno - 1, c] = rng.uniform(0.0, 0.01)
        liq[j + nsno - 1, c] = rng.uniform(0.0, 0.3)
    for c in pick(0.08):                                      # a very light layer
        j = int(rng.integers(snl[c] + 1, 1))
        ice[j + nsno - 1, c] = dz[j + nsno - 1, c] * rng.uniform(15.0, 45.0)
    for c in pick(0.06):                                      # a saturated layer (void <= 0.001)
        j = int(rng.integers(snl[c] + 1, 1))
        ice[j + nsno - 1, c] = dz[j + nsno - 1, c] * S["frac_sno_eff"][c] * 900.0
        liq[j + nsno - 1, c] = dz[j + nsno - 1, c] * S["frac_sno_eff"][c] * 30.0
    one = snowc[(snl[snowc] == -1)]
    for c in one[rng.random(len(one)) < 0.4]:                 # a pack about to disappear
        dz[nsno - 1, c] = rng.uniform(0.004, 0.02)
        ice[nsno - 1, c] = dz[nsno - 1, c] * rng.uniform(60.0, 300.0)
        S["zi"][nsno - 1, c] = -dz[nsno - 1, c]
        S["z"][nsno - 1, c] = -0.5 * dz[nsno - 1, c]
    wx = ice[:nsno] + liq[:nsno]
    S["swe_old"] = np.where(act, wx * g(0.9, 1.3, nsno, nc), 0.0)
    fio = np.zeros_like(ice)
    fio[:nsno] = np.where(act, ice[:nsno] / np.maximum(wx, 1e-30) * g(1.0, 1.1, nsno, nc), 0.0).clip(0.0, 1.0)
    S["frac_iceold"] = fio
    im = np.where(np.abs(S["imelt"]) < 3, S["imelt"], 0).astype(np.int32)
    im[:nsno] = np.where(act & (rng.random((nsno, nc)) < 0.4), 1, 0)
    S["imelt"] = im
    S["n_melt"] = 200.0 / np.maximum(10.0, g(5.0, 600.0, nc))
    S["forc_wind"] = g(0.3, 18.0, ng)
    S["snw_rds"] = np.where(act, g(54.526, 1500.0, nsno, nc), 0.0)
    for a in ("bcphi", "bcpho", "ocphi", "ocpho", "dst1", "dst2", "dst3", "dst4"):
        S["mss_" + a] = np.where(act, 10.0 ** g(-9.0, -5.0, nsno, nc), 0.0)
    S["forc_aer"] = 10.0 ** g(-14.0, -10.0, 14, ng)
    S["int_snow"] = np.maximum(S["int_snow"], (wx * act).sum(0) * g(1.0, 1.5, nc))
    # column water fluxes at the snow surface (SoilFluxes outputs: spval before a step); evaporation kept below the top layer's store
    top = np.clip(snl + nsno, 0, nsno - 1)
    ice_top, liq_top = ice[top, np.arange(nc)], liq[top, np.arange(nc)]
    fs = np.maximum(S["frac_sno_eff"], 1e-3)
    S["qflx_soliddew_to_top_layer"] = np.where(rng.random(nc) < 0.3, g(0.0, 2.0e-5, nc), 0.0)
    S["qflx_liqdew_to_top_layer"] = np.where(rng.random(nc) < 0.3, g(0.0, 2.0e-5, nc), 0.0)
    S["qflx_solidevap_from_top_layer"] = np.minimum(np.where(rng.random(nc) < 0.5, g(0.0, 4.0e-5, nc), 0.0), 0.5 * ice_top / (1800.0 * fs))
    S["qflx_liqevap_from_top_layer"] = np.minimum(np.where(rng.random(nc) < 0.5, g(0.0, 4.0e-5, nc), 0.0), 0.5 * liq_top / (1800.0 * fs))
    exact = snowc[rng.random(len(snowc)) < 0.05]              # sublimation that removes the top layer's ice to rounding (truncated to 0)
    S["qflx_solidevap_from_top_layer"][exact] = ice_top[exact] / (1800.0 * S["frac_sno_eff"][exact])
    S["qflx_soliddew_to_top_layer"][exact] = 0.0
    S["qflx_liq_grnd"] = np.where(rng.random(nc) < 0.4, 10.0 ** g(-6.0, -3.3, nc), 0.0)
    S["qflx_snomelt"] = np.where(rng.random(nc) < 0.4, g(0.0, 3.0e-4, nc), 0.0)
    S["qflx_snow_drain"] = np.where(snl < 0, S["qflx_snomelt"], 1.0e36)       # (SoilTemperature leaves the melt here on snow columns)
    for grp in ("snowwater", "snowlayers"):
        for fs_ in abi_fields(grp):
            if fs_.name not in S and fs_.ctype == "double":
                n = {"COL": nc, "PATCH": sg.npatch, "GRC": ng}[fs_.sub]
                S[fs_.name] = np.full(n if fs_.lev == "L1" else (fs_.nlev, n), 1.0e36)
    for k, v in list(S.items()):
        S[k] = np.ascontiguousarray(v)


def watertable_state(sg: Subgrid, S: Dict[str, np.ndarray], rng: np.random.Generator, saturate: bool = True) -> None:
    """Inputs of PerchedWaterTable / ThetaBasedWaterTable / RenewCondensation and of the diagnostics that close HydrologyNoDrainage
    (SURVEY.md 8f rank 3).  A third of the columns get nearly saturated layers below a random depth, so that the water tables are
    found by interpolation between layers (perched above a frozen layer, and the theta-based one above bedrock); a few per cent
    lose exactly their top-layer ice to sublimation (truncated to zero by RenewCondensation)."""
    nc = sg.ncol
    g = lambda lo, hi, *shape: rng.uniform(lo, hi, shape)
    lo = 11                                                     # row of soil level 1 in SNOSOI arrays is lo + 1
    dz, liq, ice = S["dz"], S["h2osoi_liq"], S["h2osoi_ice"]
    wet = np.nonzero(rng.random(nc) < (0.35 if saturate else 0.0))[0]       # (saturate = False: for steps that also run SoilWater,
    ktop = rng.integers(2, 16, size=len(wet))                              # whose adaptive solve is ill-conditioned on saturated layers)
    for c, k0 in zip(wet, ktop):
        for k in range(int(k0), 21):
            ratio = rng.uniform(0.91, 0.995)
            need = ratio * S["watsat"][k - 1, c] - ice[lo + k, c] / (dz[lo + k, c] * 917.0)
            if need > 0.0:
                liq[lo + k, c] = need * dz[lo + k, c] * 1000.0
    S["h2osoi_vol"] = g(0.05, 0.5, 25, nc)
    S["snow_persistence"] = g(0.0, 1.0e6, nc)
    nosnow = np.nonzero(S["snl"] == 0)[0]
    exact = nosnow[rng.random(len(nosnow)) < 0.05]
    S["qflx_soliddew_to_top_layer"][exact] = 0.0
    S["qflx_solidevap_from_top_layer"][exact] = ice[lo + 1, exact] / (1800.0 * (1.0 - S["frac_h2osfc"][exact]))
    for grp in ("watertable", "hydrodiag"):
        for fs_ in abi_fields(grp):
            if fs_.name not in S and fs_.ctype == "double":
                S[fs_.name] = np.full(nc if fs_.lev == "L1" else (fs_.nlev, nc), 1.0e36)
    for k, v in list(S.items()):
        S[k] = np.ascontiguousarray(v)


def ozone_state(sg: Subgrid, S: Dict[str, np.ndarray], rng: np.random.Generator) -> None:
    """Inputs of CalcOzoneUptake / CalcOzoneStress (OzoneMod.F90): ozone mixing ratio, the accumulated doses,
    last step's LAI, the three PFT flags.  Resistances CanopyFluxes has not produced yet (spval before a step) get plausible values."""
    npch, ng = sg.npatch, sg.ngrc
    g = lambda lo, hi, *shape: rng.uniform(lo, hi, shape)
    S["forc_o3"] = g(5.0e-9, 9.0e-8, ng)
    for nm, (lo, hi) in {"rssun": (40.0, 2.0e4), "rssha": (40.0, 2.0e4), "rb1": (5.0, 150.0), "ram1": (5.0, 300.0)}.items():
        S[nm] = np.where(np.abs(S[nm]) < 1e30, S[nm], g(lo, hi, npch))
    S["tlai_old"] = S["tlai"] * g(0.8, 1.2, npch)
    for nm in ("o3uptakesha", "o3uptakesun"):
        S[nm] = np.where(rng.random(npch) < 0.2, 0.0, g(0.0, 60.0, npch))
    nt = len(S["pft_z0v_LAImax"])
    S["pft_evergreen"] = (rng.random(nt) < 0.4).astype(np.float64)
    S["pft_leaf_long"] = g(0.5, 6.0, nt)
    S["pft_woody"] = (rng.random(nt) < 0.5).astype(np.float64)
    for nm in ("o3coefvsha", "o3coefvsun", "o3coefgsha", "o3coefgsun", "o3coefjmaxsha", "o3coefjmaxsun"):
        S[nm] = np.full(npch, 1.0e36)
    for k, v in list(S.items()):
        S[k] = np.ascontiguousarray(v)


def make_ensemble(sg: Subgrid, S: Dict[str, np.ndarray], nmember: int, rng: np.random.Generator, spread: float = 0.2):
    """Perturbed-parameter ensemble through the PFT tables (BASELINE config 5): the grid is split into `nmember` equal
    runs of gridcells, member m's patches get itype = m*(mxpft+1) + pft, and every pft_* table is extended to
    nmember*(mxpft+1) entries with member m's copy of medlynslope, kmax, psi50, ck and krmax scaled by U(1-spread,
    1+spread) factors (member 0 keeps the base table).  Returns the member index of every patch; run with
    ctsm_params_t.npft_table = nmember*(mxpft+1)."""
    npft = abi_mod().MXPFT + 1
    member_g = np.minimum((np.arange(sg.ngrc) * nmember) // max(sg.ngrc, 1), nmember - 1)
    member_p = member_g[sg.patch_gridcell - 1].astype(np.int32)
    S["itype"] = (S["itype"] + member_p * npft).astype(np.int32)
    perturbed = ("pft_medlynslope", "pft_kmax", "pft_psi50", "pft_ck", "pft_krmax")
    for k in [k for k in S if k.startswith("pft_")]:
        base = S[k]
        reps = [base]
        for m in range(1, nmember):
            reps.append(base * rng.uniform(1.0 - spread, 1.0 + spread) if k in perturbed else base)
        S[k] = np.ascontiguousarray(np.concatenate(reps, axis=-1).astype(base.dtype))
    return member_p


def abi_mod():
    from . import abi
    return abi


def abi_fields(group):
    from . import abi
    return abi.FIELDS[group]

