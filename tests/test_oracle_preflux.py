"""CPU pins of the oracle's BiogeophysPreFluxCalcs / CalculateSurfaceHumidity / BareGroundFluxes (oracle/oracle_preflux.c,
SURVEY.md 8f rank 2): vectorised NumPy restatements written from the Fortran (not from the C), and the invariants the
routines imply."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from tests.util import copy_state

TFRZ, DENH2O, DENICE = 273.15, 1000.0, 917.0
RGAS = 6.02214e26 * 1.38065e-23
ROVERG = RGAS / 18.016 / 9.80616 * 1000.0


def case(n=800, seed=401, wet_every=0):
    sg, S = synthetic_canopy.make_full_case(n, seed=seed)
    synthetic_canopy.balance_state(sg, S, np.random.Generator(np.random.PCG64(seed + 1)), 1.0e-11)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(seed + 2)))
    synthetic_canopy.preflux_state(sg, S, np.random.Generator(np.random.PCG64(seed + 3)))
    if wet_every:                                   # turn some glacier columns into wetlands (istwet = 6)
        ice = np.nonzero(S["lun_itype"] == 4)[0]
        S["lun_itype"][ice[::wet_every]] = 6
    return sg, S


def run_preflux(OL, prm, sg, S, flags=0):
    st = abi.Status()
    f = abi.make_struct("preflux", S, sg.bounds)
    fc, fp = sg.filters["nolakec"], sg.filters["nolakep"]
    return OL.oracle_biogeophys_pre_flux_calcs(C.byref(prm), C.byref(sg.bounds), len(fc), abi.i32p(fc), len(fp), abi.i32p(fp), 0,
                                               flags, C.byref(f), C.byref(st))


def run_humidity(OL, sg, S):
    st = abi.Status()
    f = abi.make_struct("surfacehumidity", S, sg.bounds)
    fc = sg.filters["nolakec"]
    return OL.oracle_calculate_surface_humidity(C.byref(sg.bounds), len(fc), abi.i32p(fc), C.byref(f), C.byref(st))


def run_bare(OL, prm, sg, S):
    st = abi.Status()
    f = abi.make_struct("baregroundfluxes", S, sg.bounds)
    fp = sg.filters["noexposedvegp"]
    return OL.oracle_bare_ground_fluxes(C.byref(prm), C.byref(sg.bounds), len(fp), abi.i32p(fp), C.byref(f), C.byref(st))


def qsat_np(T, p):
    """QSatMod.F90:61-127 in NumPy (Flatau et al. polynomials, Horner form as in the source)"""
    a = [6.11213476, 0.444007856, 0.143064234e-01, 0.264461437e-03, 0.305903558e-05, 0.196237241e-07, 0.892344772e-10,
         -0.373208410e-12, 0.209339997e-15]
    b = [0.444017302, 0.286064092e-01, 0.794683137e-03, 0.121211669e-04, 0.103354611e-06, 0.404125005e-09, -0.788037859e-12,
         -0.114596802e-13, 0.381294516e-16]
    c = [6.11123516, 0.503109514, 0.188369801e-01, 0.420547422e-03, 0.614396778e-05, 0.602780717e-07, 0.387940929e-09,
         0.149436277e-11, 0.262655803e-14]
    d = [0.503277922, 0.377289173e-01, 0.126801703e-02, 0.249468427e-04, 0.313703411e-06, 0.257180651e-08, 0.133268878e-10,
         0.394116744e-13, 0.498070196e-16]
    td = np.minimum(100.0, np.maximum(-75.0, T - TFRZ))

    def horner(k):
        r = np.full_like(td, k[8])
        for i in range(7, -1, -1):
            r = k[i] + td * r
        return r
    es = np.where(td >= 0.0, horner(a), horner(c)) * 100.0
    dd = np.where(td >= 0.0, horner(b), horner(d)) * 100.0
    vp = 1.0 / (p - 0.378 * es)
    vp1 = 0.622 * vp
    return es * vp1, es, dd * (vp1 * vp) * p


@pytest.mark.parametrize("method,resis", [(2, 1), (1, 0)], ids=["meier2022_sl14", "zengwang2007_leepielke"])
def test_preflux_matches_numpy(oracle_lib, method, resis):
    sg, S = case(wet_every=3)
    prm = abi.default_params()
    prm.z0param_method, prm.soil_resis_method = method, resis
    S["htop"][sg.filters["nolakep"][::17] - 1] = 0.0            # the htop <= 1e-10 branch (reads the previous z0mg)
    S0 = copy_state(S)
    assert run_preflux(oracle_lib, prm, sg, S) == 0
    fc, fp = sg.filters["nolakec"] - 1, sg.filters["nolakep"] - 1
    lt, fs = S0["lun_itype"][fc], S0["frac_sno"][fc]
    # ground roughness (FrictionVelocityMod.F90:601-637)
    if method == 1:
        z0mg = np.where(fs > 0.0, prm.zsno, prm.zlnd)
    else:
        sm = S0["snomelt_accum"][fc]
        snow = np.where(sm < 1.0e-5, np.exp(-1.4 * np.pi * 0.5 + (-0.31)) * 1.0e-3,
                        np.exp(1.4 * np.arctan((np.log10(np.maximum(sm, 1e-300)) + 0.23) / 0.08) + (-0.31)) * 1.0e-3)
        z0mg = np.where(fs > 0.0, snow, np.where(lt == 4, prm.zglc, prm.zlnd))
    np.testing.assert_allclose(S["z0mg"][fc], z0mg, rtol=1e-14)
    np.testing.assert_array_equal(S["z0hg"][fc], S["z0mg"][fc])
    np.testing.assert_array_equal(S["z0qg"][fc], S["z0mg"][fc])
    # SetZ0mDisp (:120-219)
    ivt, htop = S0["itype"][fp], S0["htop"][fp]
    if method == 1:
        np.testing.assert_array_equal(S["z0m"][fp], S0["pft_z0mr"][ivt] * htop)
        np.testing.assert_array_equal(S["displa"][fp], S0["pft_displar"][ivt] * htop)
    else:
        lm = S0["pft_z0v_LAImax"][ivt]
        with np.errstate(all="ignore"):
            dis = htop * (1.0 - (1.0 - np.exp(-(7.5 * lm) ** 0.5)) / (7.5 * lm) ** 0.5)
            uu = 4.0 * (S0["pft_z0v_Cs"][ivt] + S0["pft_z0v_Cr"][ivt] * lm / 2.0) ** (-0.5) / lm / S0["pft_z0v_c"][ivt]
            cw = S0["pft_z0v_cw"][ivt]
            z0 = htop * (1.0 - dis / np.where(htop > 0, htop, 1.0)) * np.exp(-0.4 * uu + np.log(cw) - 1.0 + 1.0 / cw)
        z0 = np.where(htop <= 1.0e-10, S0["z0mg"][S0["column"][fp] - 1], z0)
        veg = ivt != 0
        np.testing.assert_allclose(S["displa"][fp][veg], dis[veg], rtol=1e-13)
        np.testing.assert_allclose(S["z0m"][fp][veg], z0[veg], rtol=1e-13)
        assert np.all(S["z0m"][fp][~veg] == 0.0) and np.all(S["displa"][fp][~veg] == 0.0)
        assert np.any(htop[veg] <= 1e-10)
    # forcing heights (:657-681): vegetated soil patches use their own roughness, everything else the ground's
    col = S0["column"][fp] - 1
    gi = S0["gridcell"][fp] - 1
    ltp = S0["lun_itype"][col]
    vegp = ((ltp == 1) | (ltp == 2)) & (S0["frac_veg_nosno"][fp] != 0)
    z = np.where(vegp, S["z0mv"][fp], S["z0mg"][col])
    np.testing.assert_array_equal(S["forc_hgt_u_patch"][fp], S0["forc_hgt_u"][gi] + z + S["displa"][fp])
    np.testing.assert_array_equal(S["z0mv"][fp], S["z0m"][fp])
    # initial temperature and energy variables (:304-401)
    np.testing.assert_array_equal(S["t_ssbef"][:, fc], S0["t_soisno"][:, fc])
    snl = S0["snl"][fc]
    ttop = S0["t_soisno"][snl + 12, fc]                       # level snl+1 -> row snl+1-(-11) = snl+12
    t1 = S0["t_soisno"][12, fc]
    fse, fh, th = S0["frac_sno_eff"][fc], S0["frac_h2osfc"][fc], S0["t_h2osfc"][fc]
    tg = np.where(snl < 0, fse * ttop + (1.0 - fse - fh) * t1 + fh * th, (1 - fh) * t1 + fh * th)
    np.testing.assert_array_equal(S["t_grnd"][fc], tg)
    np.testing.assert_array_equal(S["emg"][fc], np.where(lt == 4, 0.97, (1.0 - fs) * 0.96 + fs * 0.97))
    sub = (S0["h2osoi_liq"][snl + 12, fc] <= 0.0) & (S0["h2osoi_ice"][snl + 12, fc] > 0.0)
    np.testing.assert_array_equal(S["htvp"][fc], np.where(sub, 2.501e6 + 3.337e5, 2.501e6))
    np.testing.assert_array_equal(S["thv"][fc], S0["forc_th"][fc] * (1.0 + 0.61 * S0["forc_q"][fc]))
    np.testing.assert_allclose(S["emv"][fp], 1.0 - np.exp(-(S0["elai"][fp] + S0["esai"][fp])), rtol=1e-15)
    np.testing.assert_array_equal(S["thm"][fp], S0["forc_t"][col] + 0.0098 * S["forc_hgt_t_patch"][fp])
    for k in ("eflx_sh_tot", "eflx_lh_tot", "eflx_sh_veg", "cgrnd", "cgrnds", "cgrndl"):
        assert np.all(S[k][fp] == 0.0), k
    # soil evaporative resistance (SurfaceResistanceMod.F90:279-313, :388-424)
    soil = (lt == 1) | (lt == 2)
    liq1, ice1, dz1, ws1 = S0["h2osoi_liq"][12, fc], S0["h2osoi_ice"][12, fc], S0["dz"][12, fc], S0["watsat"][0, fc]
    if resis == 0:
        wx = (liq1 / DENH2O + ice1 / DENICE) / dz1
        wfc = S0["watfc"][0, fc]
        ffc = np.maximum(np.minimum(1.0, wx / wfc), 0.01)
        beta = np.where(wx < wfc, (1.0 - fs - fh) * 0.25 * (1.0 - np.cos(np.pi * ffc)) ** 2 + fs + fh, 1.0)
        np.testing.assert_allclose(S["soilbeta"][fc][soil], beta[soil], rtol=1e-13)
        assert np.all(S["soilbeta"][fc][~soil] == 1.0)
        np.testing.assert_array_equal(S["soilresis"][fc], S0["soilresis"][fc])      # untouched by this method
    else:
        f32 = lambda v: float(np.float32(v))
        bsw1, suc1 = S0["bsw"][0, fc], S0["sucsat"][0, fc]
        vwc = np.maximum(liq1, 1.0e-6) / (dz1 * DENH2O)
        epor = np.maximum(0.01, ws1 - np.minimum(ws1, ice1 / (dz1 * DENICE)))
        aird = ws1 * (suc1 / 1.0e7) ** (1.0 / bsw1)
        d0 = f32(2.12e-5) * (t1 / f32(273.15)) ** 1.75
        eps = ws1 - aird
        dg = eps * d0 * (eps / ws1) ** (3.0 / np.maximum(3.0, bsw1))
        dsl = prm.d_max * np.maximum(0.001, prm.frac_sat_soil_dsl_init * epor - vwc) / np.maximum(0.001, prm.frac_sat_soil_dsl_init * ws1 - aird)
        dsl = np.minimum(np.maximum(dsl, 0.0), 200.0)
        sr = np.minimum(1.0e6, dsl / (dg * eps * 1.0e3) + 20.0)
        np.testing.assert_allclose(S["dsl"][fc][soil], dsl[soil], rtol=1e-12)
        np.testing.assert_allclose(S["soilresis"][fc][soil], sr[soil], rtol=1e-12)
        assert np.all(S["soilresis"][fc][~soil] == 0.0) and (~soil).any()
    # nothing outside the filters was written
    outc = np.ones(sg.ncol, dtype=bool); outc[fc] = False
    outp = np.ones(sg.npatch, dtype=bool); outp[fp] = False
    for fsn in abi.FIELDS["preflux"]:
        if fsn.intent != "IN" and fsn.sub in ("COL", "PATCH"):
            m = outc if fsn.sub == "COL" else outp
            assert np.array_equal(S[fsn.name][..., m], S0[fsn.name][..., m], equal_nan=True), fsn.name


def test_preflux_first_steps_flag_and_urban_refusal(oracle_lib):
    sg, S = case(200, 411)
    prm = abi.default_params()
    assert run_preflux(oracle_lib, prm, sg, S, flags=1) == 0          # CTSM_TIME_FIRST_STEPS: z0m = displa = 0 (:174-178)
    fp = sg.filters["nolakep"] - 1
    assert np.all(S["z0m"][fp] == 0.0) and np.all(S["displa"][fp] == 0.0)
    S["lun_itype"][sg.filters["nolakec"][3] - 1] = 8
    assert run_preflux(oracle_lib, prm, sg, S) == 16                  # CTSM_ERR_URBAN


def test_surface_humidity_matches_numpy(oracle_lib):
    sg, S = case(wet_every=2)
    prm = abi.default_params()
    assert run_preflux(oracle_lib, prm, sg, S) == 0
    S0 = copy_state(S)
    assert run_humidity(oracle_lib, sg, S) == 0
    fc = sg.filters["nolakec"] - 1
    lt, snl = S0["lun_itype"][fc], S0["snl"][fc]
    soil = (lt == 1) | (lt == 2)
    t1, pbot, fq = S0["t_soisno"][12, fc], S0["forc_pbot"][fc], S0["forc_q"][fc]
    fse, fh = S0["frac_sno_eff"][fc], S0["frac_h2osfc"][fc]
    wx = (S0["h2osoi_liq"][12, fc] / DENH2O + S0["h2osoi_ice"][12, fc] / DENICE) / S0["dz"][12, fc]
    fac = np.maximum(np.minimum(1.0, wx / S0["watsat"][0, fc]), 0.01)
    psit = np.maximum(S0["smpmin"][fc], -S0["sucsat"][0, fc] * fac ** (-S0["bsw"][0, fc]))
    hr = np.exp(psit / ROVERG / t1)
    qred = (1.0 - fse - fh) * hr + fse + fh
    np.testing.assert_allclose(S["soilalpha"][fc][soil], qred[soil], rtol=1e-13)
    assert np.all(S["soilalpha"][fc][~soil] == 1.0e36)
    qs, _, dq = qsat_np(t1, pbot)
    clip = (qs > fq) & (fq > hr * qs)
    qs_s, dq_s = np.where(clip, fq, qs), np.where(clip, 0.0, dq)
    qg_soil = hr * qs_s
    qsn, _, dqsn = qsat_np(S0["t_soisno"][snl + 12, fc], pbot)
    qg_snow = np.where(snl < 0, qsn, qg_soil)
    dqg = np.where(snl < 0, fse * dqsn + (1.0 - fse - fh) * hr * dq_s, (1.0 - fh) * hr * dq_s)
    qh, _, dqh = qsat_np(S0["t_h2osfc"][fc], pbot)
    qg_h = np.where(fh > 0.0, qh, qg_soil)
    dqg = np.where(fh > 0.0, dqg + fh * dqh, dqg)
    qg = fse * qg_snow + (1.0 - fse - fh) * qg_soil + fh * qg_h
    for name, want in (("qg_soil", qg_soil), ("qg_snow", qg_snow), ("qg_h2osfc", qg_h), ("dqgdT", dqg), ("qg", qg)):
        np.testing.assert_allclose(S[name][fc][soil], want[soil], rtol=1e-12, err_msg=name)
    # wetland / glacier: saturated at t_grnd unless the air is between (:218-234)
    qs2, _, dq2 = qsat_np(S0["t_grnd"][fc], pbot)
    clip2 = (qs2 > fq) & (fq > qs2)          # qred = 1: never true
    assert not clip2.any()
    ns = ~soil
    np.testing.assert_allclose(S["qg"][fc][ns], qs2[ns], rtol=1e-12)
    np.testing.assert_allclose(S["dqgdT"][fc][ns], dq2[ns], rtol=1e-12)
    np.testing.assert_array_equal(S["qg_snow"][fc][ns], S["qg"][fc][ns])
    assert ns.sum() > 10 and (lt == 6).any()


@pytest.mark.parametrize("method,resis", [(2, 1), (1, 0)], ids=["meier2022_sl14", "zengwang2007_leepielke"])
def test_bare_ground_fluxes_invariants(oracle_lib, method, resis):
    sg, S = case(wet_every=3)
    prm = abi.default_params()
    prm.z0param_method, prm.soil_resis_method = method, resis
    assert run_preflux(oracle_lib, prm, sg, S) == 0
    assert run_humidity(oracle_lib, sg, S) == 0
    S0 = copy_state(S)
    assert run_bare(oracle_lib, prm, sg, S) == 0
    fp = sg.filters["noexposedvegp"] - 1
    col = S0["column"][fp] - 1
    gi = S0["gridcell"][fp] - 1
    assert len(fp) > 500
    assert np.all(S["num_iter"][fp] == 3.0)                                     # niters (:82)
    assert np.all(S["btran"][fp] == 0.0) and np.all(S["rootr"][:, fp] == 0.0) and np.all(S["qflx_tran_veg"][fp] == 0.0)
    np.testing.assert_array_equal(S["t_veg"][fp], S0["forc_t"][col])
    for k in ("displa", "z0mv", "dlrad", "ulrad", "dhsdt_canopy", "eflx_sh_stem"):
        assert np.all(S[k][fp] == 0.0), k
    # flux identities (:431-457)
    rho, thm = S0["forc_rho"][col], S0["thm"][fp]
    raih = S["cgrnds"][fp]
    dth = thm - S0["t_grnd"][col]
    np.testing.assert_array_equal(S["eflx_sh_grnd"][fp], -raih * dth)
    np.testing.assert_array_equal(S["eflx_sh_tot"][fp], S["eflx_sh_grnd"][fp])
    np.testing.assert_array_equal(S["eflx_sh_soil"][fp], -raih * (thm - S0["t_soisno"][12, col]))
    np.testing.assert_array_equal(S["cgrnd"][fp], S["cgrnds"][fp] + S0["htvp"][col] * S["cgrndl"][fp])
    dqh = S0["forc_q"][col] - S0["qg"][col]
    raiw = np.where(S0["dqgdT"][col] != 0.0, S["cgrndl"][fp] / np.where(S0["dqgdT"][col] != 0.0, S0["dqgdT"][col], 1.0), np.nan)
    ok = np.isfinite(raiw)
    np.testing.assert_allclose(S["qflx_evap_soi"][fp][ok], (-raiw * dqh)[ok], rtol=1e-12, atol=1e-30)
    np.testing.assert_array_equal(S["qflx_evap_tot_patch"][fp], S["qflx_evap_soi"][fp])
    # aerodynamics: ram1 = um/ustar^2, stress against the wind (:404, :443-444)
    np.testing.assert_allclose(S["ram1"][fp], 1.0 / (S["ustar"][fp] * S["ustar"][fp] / S["um"][fp]), rtol=1e-15)
    np.testing.assert_array_equal(S["taux"][fp], -rho * S0["forc_u"][gi] / S["ram1"][fp])
    assert np.all((S["zeta"][fp] >= -100.0) & (S["zeta"][fp] <= prm.zetamaxstable) & (np.abs(S["zeta"][fp]) >= 0.01))
    np.testing.assert_allclose(S["obu"][fp] * S["zeta"][fp], S0["forc_hgt_u_patch"][fp], rtol=1e-14)     # zldis = the entry height
    # scalar roughness and forcing heights (:352-365)
    z0m, z0h = S["z0mg_p"][fp], S["z0hg_p"][fp]
    np.testing.assert_array_equal(z0m, S0["z0mg"][col])
    np.testing.assert_array_equal(S["z0qg_p"][fp], z0h)
    np.testing.assert_array_equal(S["forc_hgt_t_patch"][fp], S0["forc_hgt_t"][gi] + z0h + 0.0)
    np.testing.assert_allclose(S["kbm1"][fp], np.log(z0m / z0h), rtol=1e-13)
    if method == 1:
        np.testing.assert_allclose(z0h, z0m / np.exp(prm.a_coef * (S["ustar"][fp] * z0m / 1.5e-5) ** prm.a_exp), rtol=1e-12)
    # the column keeps the roughness of its last patch in the filter (:468-469)
    last = np.r_[col[1:] != col[:-1], True]
    np.testing.assert_array_equal(S["z0hg"][col[last]], z0h[last])
    # 2 m diagnostics: humidity bounded, rural copies only on soil / crop
    assert np.all((S["rh_ref2m"][fp] >= 0.0) & (S["rh_ref2m"][fp] <= 100.0))
    lt = S0["lun_itype"][col]
    rural = (lt == 1) | (lt == 2)
    np.testing.assert_array_equal(S["t_ref2m_r"][fp][rural], S["t_ref2m"][fp][rural])
    np.testing.assert_array_equal(S["t_ref2m_r"][fp][~rural], S0["t_ref2m_r"][fp][~rural])
    np.testing.assert_array_equal(S["tc_ref2m"][fp], S["t_ref2m"][fp] - TFRZ)
    assert (~rural).sum() > 10
    # patches with exposed vegetation were not touched
    ex = sg.filters["exposedvegp"] - 1
    for fsn in abi.FIELDS["baregroundfluxes"]:
        if fsn.intent != "IN" and fsn.sub == "PATCH":
            assert np.array_equal(S[fsn.name][..., ex], S0[fsn.name][..., ex], equal_nan=True), fsn.name
