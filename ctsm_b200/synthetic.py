"""Deterministic synthetic subgrid + state for the hot-path routines.

Follows SURVEY.md section 8(d): G land gridcells, each with one vegetated
landunit (istsoil) holding one soil column with 15 patches (bare ground + 14
natural PFTs, Dirichlet(1) weights with >= 3 zero-weight inactive patches);
optionally every `special_every`-th gridcell also carries a land-ice column and
a lake column so that the filters are not trivially "everything".

Index conventions are the reference's (decompMod.F90:349-424): 1-based
proc-local g/l/c/p indices, g < l < c < p nesting, patches of a column are
contiguous (ColumnType.F90 patchi/patchf).  Arrays are stored level-major,
shape (nlev, n) == Fortran (n, nlev).

Parameter values marked "external" in SURVEY.md Appendix D are synthetic
choices, not CTSM parameter-file values.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict

import numpy as np

from . import abi
from .abi import NLEVSNO, NLEVGRND, NLEVSOI, ISTSOIL, ISTICE, ISTDLAK

GRID_SIZES = {"tiny": 64, "f19": 5500, "f09": 21000, "f02": 336000}
NPATCH_PER_SOILCOL = 15
TFRZ = 273.15
DENH2O, DENICE = 1000.0, 917.0
HVAP, HSUB = 2.501e6, 2.501e6 + 3.337e5


def vertical_grid():
    """20SL_8.5m soil structure, initVerticalMod.F90:291-315."""
    dzsoi = np.zeros(NLEVGRND + 1)
    for j in range(1, 5):
        dzsoi[j] = j * 0.02
    for j in range(5, 14):
        dzsoi[j] = dzsoi[4] + (j - 4) * 0.04
    for j in range(14, NLEVSOI + 1):
        dzsoi[j] = dzsoi[13] + (j - 13) * 0.10
    for j in range(NLEVSOI + 1, NLEVGRND + 1):
        dzsoi[j] = dzsoi[NLEVSOI] + (((j - NLEVSOI) * 25.0) ** 1.5) / 100.0
    zisoi = np.zeros(NLEVGRND + 1)
    for j in range(1, NLEVGRND + 1):
        zisoi[j] = np.sum(dzsoi[1:j + 1])
    zsoi = np.zeros(NLEVGRND + 1)
    for j in range(1, NLEVGRND + 1):
        zsoi[j] = 0.5 * (zisoi[j - 1] + zisoi[j])
    return dzsoi, zisoi, zsoi


@dataclass
class Subgrid:
    """Topology arrays (all 1-based indices stored in 0-based numpy arrays)."""
    bounds: abi.Bounds
    col_gridcell: np.ndarray
    col_lun_itype: np.ndarray
    col_active: np.ndarray
    col_patchi: np.ndarray
    col_patchf: np.ndarray
    patch_column: np.ndarray
    patch_gridcell: np.ndarray
    patch_itype: np.ndarray
    patch_wtcol: np.ndarray
    patch_active: np.ndarray
    filters: Dict[str, np.ndarray] = field(default_factory=dict)

    @property
    def ncol(self):
        return self.bounds.endc - self.bounds.begc + 1

    @property
    def npatch(self):
        return self.bounds.endp - self.bounds.begp + 1

    @property
    def ngrc(self):
        return self.bounds.endg - self.bounds.begg + 1


def build_subgrid(ngrc: int, rng: np.random.Generator, special_every: int = 10) -> Subgrid:
    """g -> l -> c -> p hierarchy with proc-local 1-based indices starting at 1."""
    has_special = np.zeros(ngrc, dtype=bool)
    if special_every > 0:
        has_special[special_every - 1::special_every] = True
    ncol_g = 1 + 2 * has_special.astype(np.int64)          # soil (+ ice + lake)
    ncol = int(ncol_g.sum())
    col_gridcell = np.repeat(np.arange(1, ngrc + 1, dtype=np.int32), ncol_g)
    first_col_of_g = np.concatenate([[0], np.cumsum(ncol_g)[:-1]])
    col_lun_itype = np.full(ncol, ISTSOIL, dtype=np.int32)
    sp = np.nonzero(has_special)[0]
    col_lun_itype[first_col_of_g[sp] + 1] = ISTICE
    col_lun_itype[first_col_of_g[sp] + 2] = ISTDLAK
    npatch_c = np.where(col_lun_itype == ISTSOIL, NPATCH_PER_SOILCOL, 1).astype(np.int64)
    npatch = int(npatch_c.sum())
    col_patchi = (np.concatenate([[0], np.cumsum(npatch_c)[:-1]]) + 1).astype(np.int32)
    col_patchf = (col_patchi + npatch_c - 1).astype(np.int32)
    patch_column = np.repeat(np.arange(1, ncol + 1, dtype=np.int32), npatch_c)
    patch_gridcell = col_gridcell[patch_column - 1]
    # pft index within soil columns: 0 = bare ground, 1..14 natural PFTs
    within = (np.arange(npatch, dtype=np.int64) - (col_patchi[patch_column - 1] - 1)).astype(np.int32)
    patch_itype = np.where(col_lun_itype[patch_column - 1] == ISTSOIL, within, 0).astype(np.int32)
    # weights: Dirichlet(1) over 15 patches with >= 3 zero weights
    soil_cols = np.nonzero(col_lun_itype == ISTSOIL)[0]
    w = rng.exponential(1.0, size=(soil_cols.size, NPATCH_PER_SOILCOL))
    nzero = rng.integers(3, 9, size=soil_cols.size)
    order = np.argsort(rng.random((soil_cols.size, NPATCH_PER_SOILCOL)), axis=1)
    rank = np.empty_like(order)
    np.put_along_axis(rank, order, np.arange(NPATCH_PER_SOILCOL)[None, :].repeat(soil_cols.size, 0), axis=1)
    w[rank < nzero[:, None]] = 0.0
    w /= w.sum(axis=1, keepdims=True)
    patch_wtcol = np.ones(npatch)
    idx = (col_patchi[soil_cols] - 1)[:, None] + np.arange(NPATCH_PER_SOILCOL)[None, :]
    patch_wtcol[idx.ravel()] = w.ravel()
    patch_active = patch_wtcol > 0.0
    col_active = np.ones(ncol, dtype=bool)
    b = abi.Bounds()
    b.begg, b.endg = 1, ngrc
    b.begl, b.endl = 1, ncol            # one landunit per column in this synthetic grid
    b.begc, b.endc = 1, ncol
    b.begp, b.endp = 1, npatch
    b.begCohort, b.endCohort = 1, 0
    b.level, b.clump_index = 1, -1
    sg = Subgrid(b, col_gridcell, col_lun_itype, col_active, col_patchi, col_patchf, patch_column,
                 patch_gridcell, patch_itype, patch_wtcol, patch_active)
    build_filters(sg)
    return sg


def build_filters(sg: Subgrid) -> None:
    """Python restatement of the filterMod.F90:303-592 lists the hot path uses
    (ascending index order, 1-based): nolakec, nolakep, hydrologyc, nolakeurbanp.
    (The oracle's C restatement is checked against this in tests.)"""
    lake_c = sg.col_lun_itype == ISTDLAK
    c1 = np.arange(1, sg.ncol + 1, dtype=np.int32)
    p1 = np.arange(1, sg.npatch + 1, dtype=np.int32)
    sg.filters["nolakec"] = c1[sg.col_active & ~lake_c]
    sg.filters["lakec"] = c1[sg.col_active & lake_c]
    hyd = np.isin(sg.col_lun_itype, (ISTSOIL, abi.ISTCROP))
    sg.filters["hydrologyc"] = c1[sg.col_active & hyd]
    lake_p = lake_c[sg.patch_column - 1]
    sg.filters["nolakep"] = p1[sg.patch_active & ~lake_p]
    sg.filters["nolakeurbanp"] = sg.filters["nolakep"].copy()   # no urban landunits in the synthetic grid


def _lev(name):
    lo, n = abi.LEV[name]
    return lo, n


def soil_state(sg: Subgrid, rng: np.random.Generator) -> Dict[str, np.ndarray]:
    """Column/patch state for SoilTemperature + SoilWater."""
    nc, npch = sg.ncol, sg.npatch
    S = {}
    dzsoi, zisoi, zsoi = vertical_grid()
    lo, nl = _lev("SNOSOI")              # -11 .. 25
    snl_class = rng.random(nc)
    snl = np.zeros(nc, dtype=np.int32)
    m1 = (snl_class >= 0.4) & (snl_class < 0.6)
    m2 = snl_class >= 0.6
    snl[m1] = -rng.integers(1, 5, size=int(m1.sum()))
    snl[m2] = -rng.integers(5, NLEVSNO + 1, size=int(m2.sum()))
    S["snl"] = snl
    S["lun_itype"] = sg.col_lun_itype.copy()
    S["patchi"], S["patchf"] = sg.col_patchi.copy(), sg.col_patchf.copy()
    S["column"] = sg.patch_column.copy()
    S["wtcol"] = sg.patch_wtcol.copy()
    S["nbedrock"] = rng.integers(5, NLEVSOI + 1, size=nc).astype(np.int32)

    dz = np.zeros((nl, nc)); z = np.zeros((nl, nc)); zi = np.zeros((nl + 1, nc))
    for j in range(1, NLEVGRND + 1):
        dz[j - lo] = dzsoi[j]; z[j - lo] = zsoi[j]; zi[j - (lo - 1)] = zisoi[j]
    zi[0 - (lo - 1)] = 0.0
    dzs = rng.uniform(0.01, 0.3, size=(NLEVSNO, nc))
    for j in range(0, -NLEVSNO, -1):                 # j = 0, -1, ..., -11
        act = j >= snl + 1
        dz[j - lo] = np.where(act, dzs[-j], 0.0)
        zi[(j - 1) - (lo - 1)] = zi[j - (lo - 1)] - dz[j - lo]
        z[j - lo] = np.where(act, 0.5 * (zi[(j - 1) - (lo - 1)] + zi[j - (lo - 1)]), 0.0)
    S["dz"], S["z"], S["zi"] = dz, z, zi

    g = lambda a, b, *sh: rng.uniform(a, b, size=sh)
    # soil texture varies smoothly with depth within a column (per-column base + small per-layer noise)
    def prof(a, b, rel):
        base = g(a, b, nc)[None, :]
        return base * (1.0 + rel * rng.standard_normal((NLEVGRND, nc)).clip(-2.5, 2.5))
    S["watsat"] = prof(0.35, 0.55, 0.02)
    S["tkmg"] = prof(1.5, 4.0, 0.05)
    S["tkdry"] = prof(0.15, 0.4, 0.05)
    S["csol"] = prof(2.0e6, 2.5e6, 0.02)
    S["bsw"] = prof(3.0, 12.0, 0.03)
    S["sucsat"] = prof(50.0, 400.0, 0.05)
    S["hksat"] = 10.0 ** g(-4.0, -2.0, nc)[None, :] * (1.0 + 0.1 * rng.standard_normal((NLEVGRND, nc)).clip(-2.5, 2.5))

    # temperature profile: surface 255..300 K relaxing to a deep value, + noise
    tsurf = g(255.0, 300.0, nc); tdeep = g(268.0, 290.0, nc)
    t = np.zeros((nl, nc))
    for j in range(1, NLEVGRND + 1):
        wgt = np.exp(-zsoi[j] / 1.5)
        t[j - lo] = wgt * tsurf + (1 - wgt) * tdeep + rng.normal(0, 0.5, nc)
    for j in range(0, -NLEVSNO, -1):
        act = j >= snl + 1
        tsn = np.minimum(tsurf, 273.15 + rng.normal(-3.0, 4.0, nc))
        t[j - lo] = np.where(act, tsn, 0.0)
    S["t_soisno"] = t
    # smooth saturation profile per column (a rough random profile would force the
    # Richards solver to its 60 s minimum substep everywhere, which is not a realistic mix)
    s0 = g(0.2, 0.85, nc); amp = g(0.0, 0.12, nc); ph = g(0.0, 6.28, nc)
    zz = zsoi[1:NLEVGRND + 1][:, None]
    sat = np.clip(s0[None, :] + amp[None, :] * np.sin(1.3 * zz + ph[None, :]) + rng.normal(0, 0.01, (NLEVGRND, nc)),
                  0.1, 0.95)
    liq = np.zeros((nl, nc)); ice = np.zeros((nl, nc))
    for j in range(1, NLEVGRND + 1):
        vol = sat[j - 1] * S["watsat"][j - 1]
        frozen = t[j - lo] < TFRZ
        fice = np.where(frozen, g(0.3, 0.9, nc), 0.0)
        liq[j - lo] = vol * (1 - fice) * dzsoi[j] * DENH2O
        ice[j - lo] = vol * fice * dzsoi[j] * DENICE
    rho_sno = g(80.0, 450.0, NLEVSNO, nc)
    for j in range(0, -NLEVSNO, -1):
        act = j >= snl + 1
        wet = t[j - lo] >= TFRZ - 0.5
        ice[j - lo] = np.where(act, dz[j - lo] * rho_sno[-j], 0.0)
        liq[j - lo] = np.where(act & wet, 0.08 * ice[j - lo] * rng.random(nc), 0.0)
    S["h2osoi_liq"], S["h2osoi_ice"] = liq, ice

    thin = (snl == 0) & (rng.random(nc) < 0.25)
    S["h2osno_no_layers"] = np.where(thin, g(0.1, 8.0, nc), 0.0)
    S["frac_sno_eff"] = np.where(snl < 0, g(0.3, 1.0, nc), np.where(thin, g(0.02, 0.3, nc), 0.0))
    sd = np.zeros(nc)
    for j in range(0, -NLEVSNO, -1):
        sd += dz[j - lo]
    S["snow_depth"] = np.where(snl < 0, sd, S["h2osno_no_layers"] / 250.0)
    S["int_snow"] = (ice[:NLEVSNO].sum(0) + liq[:NLEVSNO].sum(0) + S["h2osno_no_layers"]) * g(1.0, 1.5, nc)
    S["snomelt_accum"] = g(0.0, 0.01, nc)
    wetsfc = rng.random(nc) < 0.3
    S["frac_h2osfc"] = np.where(wetsfc, g(0.001, 0.3, nc), 0.0)
    S["h2osfc"] = np.where(wetsfc, 10.0 ** g(-7.0, 1.3, nc), 0.0)
    S["t_h2osfc"] = t[1 - lo] + rng.normal(0.0, 1.0, nc)
    top = np.take_along_axis(t, (snl + 1 - lo)[None, :], axis=0)[0]
    S["t_grnd"] = (S["frac_sno_eff"] * top + (1 - S["frac_sno_eff"] - S["frac_h2osfc"]) * t[1 - lo]
                   + S["frac_h2osfc"] * S["t_h2osfc"])
    S["forc_lwrad"] = g(150.0, 450.0, nc)
    S["emg"] = g(0.96, 0.97, nc)
    S["htvp"] = np.where(t[1 - lo] > TFRZ, HVAP, HSUB)
    S["eflx_bot"] = np.zeros(nc)
    S["eflx_snomelt_r"] = np.zeros(nc)

    # outputs (pre-filled with a recognisable value so untouched elements are visible)
    fill = 1.0e36
    for name, lev in (("thk", "SNOSOI"), ("bw", "SNO"), ("fact", "SNOSOI"), ("eflx_fgr", "GRND"),
                      ("qflx_snomelt_lyr", "SNO"), ("qflx_snofrz_lyr", "SNO"),
                      ("smp_l", "GRND"), ("hk_l", "GRND"), ("qin", "SOI"), ("qout", "SOI")):
        S[name] = np.full((abi.LEV[lev][1], nc), fill)
    S["imelt"] = np.full((nl, nc), -9999, dtype=np.int32)
    for name in ("c_h2osfc", "xmf", "xmf_h2osfc", "eflx_fgr12", "qflx_h2osfc_to_ice", "eflx_h2osfc_to_snow",
                 "qflx_snow_drain", "qflx_snofrz", "qflx_snomelt", "eflx_snomelt", "qcharge", "num_substeps"):
        S[name] = np.full(nc, fill)

    # patch level (ComputeGroundHeatFluxAndDeriv inputs)
    bare = sg.patch_itype == 0
    S["frac_veg_nosno"] = np.where(bare, 0, (rng.random(npch) < 0.85).astype(np.int32)).astype(np.int32)
    sabg = g(0.0, 300.0, npch) * (rng.random(npch) < 0.6)
    S["sabg"] = sabg
    S["sabg_soil"] = sabg * g(0.9, 1.0, npch)
    S["sabg_snow"] = sabg * g(0.2, 0.6, npch)
    l1, n1 = _lev("SNO1")
    wl = rng.random((n1, npch))
    psnl = snl[sg.patch_column - 1]
    levs = np.arange(l1, l1 + n1)[:, None]
    wl = np.where(levs >= (psnl + 1)[None, :], wl, 0.0)
    wl /= wl.sum(0, keepdims=True)
    S["sabg_lyr"] = wl * sabg[None, :]
    S["dlrad"] = g(0.0, 60.0, npch)
    S["cgrnd"] = g(5.0, 30.0, npch)
    for nm in ("eflx_sh_grnd", "eflx_sh_snow", "eflx_sh_soil", "eflx_sh_h2osfc"):
        S[nm] = rng.normal(20.0, 30.0, npch)
    for nm in ("qflx_evap_soi", "qflx_ev_snow", "qflx_ev_soil", "qflx_ev_h2osfc"):
        S[nm] = rng.normal(1.0e-5, 2.0e-5, npch)
    for nm in ("eflx_gnet", "dgnetdT", "sabg_chk"):
        S[nm] = np.full(npch, fill)

    # SoilWater inputs (SoilHydrologyMod icefrac / SoilMoistStress eff_porosity definitions)
    vol_ice = np.minimum(S["watsat"], ice[NLEVSNO:NLEVSNO + NLEVGRND] / (dz[NLEVSNO:NLEVSNO + NLEVGRND] * DENICE))
    S["icefrac"] = np.minimum(1.0, vol_ice / S["watsat"])
    S["eff_porosity"] = np.maximum(0.01, S["watsat"] - vol_ice)
    S["qflx_infl"] = g(0.0, 5.0e-4, nc) * (rng.random(nc) < 0.7)
    rootfr = np.exp(-zsoi[1:NLEVSOI + 1] / 0.8)[:, None] * np.ones((1, nc))
    rootfr /= rootfr.sum(0, keepdims=True)
    S["qflx_rootsoi"] = rootfr * g(0.0, 2.0e-5, nc)[None, :]
    return S


def make_case(size="tiny", seed: int = 20260101, special_every: int = 10):
    ngrc = GRID_SIZES[size] if isinstance(size, str) else int(size)
    rng = np.random.Generator(np.random.PCG64(seed))
    sg = build_subgrid(ngrc, rng, special_every)
    S = soil_state(sg, rng)
    return sg, S
