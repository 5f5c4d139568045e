"""CPU checks of the SoilFluxes oracle (SoilFluxesMod.F90:37-521): the reference's tests hold no vector for it (parity
unpinned), so it is checked against an independent numpy restatement of its formulas, against the identities the routine
guarantees, and for clump-decomposition independence."""
import ctypes as C

import numpy as np

from ctsm_b200 import abi, synthetic_canopy
from oracle import oracle


def _case(n=400, seed=71):
    OL = oracle.lib()
    prm = abi.default_params()
    sg, S = synthetic_canopy.make_full_case(n, seed=seed)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(seed + 2)))
    st = abi.Status()
    fe = sg.filters["exposedvegp"]
    f = abi.make_struct("canopyfluxes", S, sg.bounds)
    assert OL.oracle_canopyfluxes(C.byref(prm), C.byref(sg.bounds), len(fe), abi.i32p(fe), C.byref(f), C.byref(st)) == 0
    fp, fc = sg.filters["nolakep"], sg.filters["nolakec"]
    ft = abi.make_struct("soiltemperature", S, sg.bounds)
    assert OL.oracle_soiltemperature(C.byref(prm), C.byref(sg.bounds), len(fp), abi.i32p(fp), len(fc), abi.i32p(fc),
                                     C.byref(ft), C.byref(st)) == 0
    return OL, prm, sg, S


def _run(OL, prm, sg, S, bounds=None, fc=None, fp=None):
    st = abi.Status()
    f = abi.make_struct("soilfluxes", S, sg.bounds)
    fc = sg.filters["nolakec"] if fc is None else fc
    fp = sg.filters["nolakep"] if fp is None else fp
    b = sg.bounds if bounds is None else bounds
    return OL.oracle_soilfluxes(C.byref(prm), C.byref(b), len(fc), abi.i32p(fc), len(fp), abi.i32p(fp), C.byref(f), C.byref(st))


def test_soilfluxes_formulas_and_identities():
    OL, prm, sg, S = _case()
    before = {k: v.copy() for k, v in S.items()}
    assert _run(OL, prm, sg, S) == 0
    p = sg.filters["nolakep"] - 1
    c = S["column"][p] - 1
    lo = -abi.NLEVSNO + 1 if hasattr(abi, "NLEVSNO") else -11
    snl = S["snl"][c]
    tss_top = before["t_ssbef"][snl + 1 - lo, c]
    tss1 = before["t_ssbef"][1 - lo, c]
    fs, fh = S["frac_sno_eff"][c], S["frac_h2osfc"][c]
    tg0 = np.where(snl < 0, fs * tss_top + (1 - fs - fh) * tss1 + fh * S["t_h2osfc_bef"][c], (1 - fh) * tss1 + fh * S["t_h2osfc_bef"][c])
    tinc = S["t_grnd"][c] - tg0
    # flux correction :188-206 (before the evaporation limit touches the limited patches)
    ev_soil = before["qflx_ev_soil"][p] + tinc * before["cgrndl"][p]
    assert np.array_equal(S["qflx_ev_soil"][p], ev_soil)
    hvap, sb = 2.501e6, 5.67e-8
    htvp, emg, lw = S["htvp"][c], S["emg"][c], S["forc_lwrad"][c]
    assert np.array_equal(S["eflx_lh_tot"][p], hvap * S["qflx_evap_veg"][p] + htvp * S["qflx_evap_soi"][p])
    assert np.array_equal(S["eflx_lwrad_net"][p], S["eflx_lwrad_out"][p] - lw)
    assert np.array_equal(S["eflx_sh_tot"][p], (S["eflx_sh_veg"][p] + S["eflx_sh_grnd"][p]) + S["eflx_sh_stem"][p])
    lw_grnd = fs * tss_top ** 4 + (1.0 - fs - fh) * tss1 ** 4 + fh * S["t_h2osfc_bef"][c] ** 4
    fv = S["frac_veg_nosno"][p]
    want = ((1.0 - fs) * S["sabg_soil"][p] + fs * S["sabg_snow"][p]) + S["dlrad"][p] + (1 - fv) * emg * lw - emg * sb * lw_grnd \
        - emg * sb * tg0 ** 3 * (4.0 * tinc) - (S["eflx_sh_grnd"][p] + S["qflx_evap_soi"][p] * htvp)
    scale = np.abs(emg * sb * lw_grnd) + np.abs(S["eflx_sh_grnd"][p]) + 1.0
    assert np.max(np.abs(S["eflx_soil_grnd"][p] - want) / scale) < 1e-13
    # the partition closes wherever the snow limit (:283-292) did not act
    part = (S["qflx_liqevap_from_top_layer_patch"] + S["qflx_solidevap_from_top_layer_patch"]
            - S["qflx_liqdew_to_top_layer_patch"] - S["qflx_soliddew_to_top_layer_patch"])[p]
    close = np.abs(part - S["qflx_ev_snow"][p]) <= 1e-12 * np.maximum(np.abs(S["qflx_ev_snow"][p]), 1e-12)
    assert close.mean() > 0.9
    assert np.all(S["qflx_liqevap_from_top_layer_patch"][p] >= 0) and np.all(S["qflx_soliddew_to_top_layer_patch"][p] >= 0)
    # bare-ground skin temperature only where there is no exposed vegetation
    bare = fv == 0
    assert np.allclose(S["t_skin"][p][bare], np.sqrt(np.sqrt(lw_grnd[bare])), rtol=1e-15)
    assert np.array_equal(S["t_skin"][p][~bare], before["t_skin"][p][~bare])
    # p2c :312-318
    for ci in (sg.filters["nolakec"] - 1)[:50]:
        pi, pf = S["patchi"][ci] - 1, S["patchf"][ci]
        s = 0.0
        for q in range(pi, pf):
            if S["patch_active"][q]:
                s = s + S["errsoi_patch"][q] * S["wtcol"][q]
        assert S["errsoi_col"][ci] == s
    # fields outside the filter are untouched
    mask = np.ones(sg.npatch, bool); mask[p] = False
    for k in ("eflx_soil_grnd", "eflx_sh_grnd", "eflx_lwrad_out", "errsoi_patch"):
        assert np.array_equal(S[k][mask], before[k][mask], equal_nan=True)


def test_soilfluxes_is_decomposition_independent():
    OL, prm, sg, S = _case(300, 73)
    whole = {k: v.copy() for k, v in S.items()}
    parts = {k: v.copy() for k, v in S.items()}
    assert _run(OL, prm, sg, whole) == 0
    clumps, keep = oracle.make_clumps(sg, 5)
    f = abi.make_struct("soilfluxes", parts, sg.bounds)
    st = abi.Status()
    for k in clumps:
        assert OL.oracle_soilfluxes(C.byref(prm), C.byref(k.bounds), k.num_nolakec, k.filter_nolakec, k.num_nolakep,
                                    k.filter_nolakep, C.byref(f), C.byref(st)) == 0
    for fs in abi.FIELDS["soilfluxes"]:
        assert np.array_equal(whole[fs.name], parts[fs.name], equal_nan=True), fs.name


def test_soilfluxes_refuses_urban_columns():
    OL, prm, sg, S = _case(64, 75)
    c = int(sg.filters["nolakec"][2])
    S["lun_itype"][c - 1] = 7
    assert _run(OL, prm, sg, S) == 16


def test_patch2col_is_the_weighted_column_mean():
    OL, prm, sg, S = _case(200, 77)
    assert _run(OL, prm, sg, S) == 0
    allc = np.arange(sg.bounds.begc, sg.bounds.endc + 1, dtype=np.int32)
    fc = sg.filters["nolakec"]
    f = abi.make_struct("patch2col", S, sg.bounds)
    assert OL.oracle_patch2col(C.byref(sg.bounds), len(allc), abi.i32p(allc), len(fc), abi.i32p(fc), C.byref(f)) == 0
    w = S["wtcol"] * S["patch_active"]
    for pname, cname, flt in (("qflx_ev_snow", "qflx_ev_snow_col", fc), ("qflx_evap_tot_patch", "qflx_evap_tot", fc),
                              ("qflx_liqdew_to_top_layer_patch", "qflx_liqdew_to_top_layer", fc),
                              ("qflx_evap_soi", "qflx_evap_soi_col", allc)):
        src = np.where(np.abs(S[pname]) < 1e30, S[pname], 0.0) * np.where(np.abs(S[pname]) < 1e30, w, 0.0)
        for ci in (flt - 1)[:80]:
            pi, pf = S["patchi"][ci] - 1, S["patchf"][ci]
            act = S["patch_active"][pi:pf] != 0
            if not np.all(np.abs(S[pname][pi:pf][act]) < 1e30):
                continue
            want = float(np.sum(src[pi:pf]))
            assert abs(S[cname][ci] - want) <= 1e-14 * max(1.0, np.abs(src[pi:pf]).sum()) + 1e-300, (cname, ci)
