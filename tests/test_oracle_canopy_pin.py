"""CPU pin of the oracle's surface-layer helpers and of its CanopyFluxes (oracle/oracle_canopy.c) by tests/canopy_python.py, plain
Python written from FrictionVelocityMod.F90, QSatMod.F90 and CanopyFluxesMod.F90: identical bits."""
import ctypes as C

import numpy as np
import pytest

from tests import canopy_python as cp


class FricVel(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("ustar", "temp1", "temp2", "temp12m", "temp22m", "fm", "vds", "u10_clm", "va", "u10", "fv")]


def test_friction_velocity_monin_obukhov_qsat_match_python(oracle_lib):
    OL = oracle_lib
    d = C.c_double
    OL.oracle_friction_velocity_point.argtypes = [d] * 8 + [C.c_int, d, d, C.POINTER(FricVel)]
    OL.oracle_friction_velocity_point.restype = None
    OL.oracle_moninobukini.argtypes = [d] * 6 + [C.POINTER(d)] * 2
    OL.oracle_moninobukini.restype = None
    OL.oracle_qsat.argtypes = [d, d] + [C.POINTER(d)] * 3
    OL.oracle_qsat.restype = None
    rng = np.random.Generator(np.random.PCG64(1301))
    regimes = {"very_unstable": 0, "unstable": 0, "stable": 0, "very_stable": 0, "tall": 0, "same_hgt": 0}
    for n in range(6000):
        hgt_u = float(rng.uniform(2.5, 60.0))
        same = n % 3 == 0
        hgt_t = hgt_u if same else float(rng.uniform(2.5, 60.0))
        hgt_q = hgt_t if n % 2 == 0 else float(rng.uniform(2.5, 60.0))
        displa = float(rng.uniform(0.0, 0.6 * min(hgt_u, hgt_t, hgt_q)))
        z0m = float(10 ** rng.uniform(-4, 0.3))
        z0h = z0m if n % 5 == 0 else float(z0m * 10 ** rng.uniform(-3, 0))
        z0q = z0h if n % 2 == 0 else float(z0m * 10 ** rng.uniform(-3, 0))
        zl = hgt_u - displa
        zeta = float(rng.choice([-1.0, 1.0]) * 10 ** rng.uniform(-2.5, 1.3))
        if n % 97 == 0:
            zeta = float(rng.choice([-1.574, -0.465, 1.0]))       # the regime boundaries themselves
        obu = zl / zeta
        it = int(rng.integers(1, 5))
        ur, um = float(rng.uniform(0.1, 25.0)), float(rng.uniform(0.1, 25.0))
        fm0 = float(rng.uniform(-5, 3))
        o = FricVel(fm=fm0)
        OL.oracle_friction_velocity_point(hgt_u, hgt_t, hgt_q, displa, z0m, z0h, z0q, obu, it, ur, um, C.byref(o))
        r = cp.friction_velocity(hgt_u, hgt_t, hgt_q, displa, z0m, z0h, z0q, obu, it, ur, um, fm0)
        for k, _ in FricVel._fields_:
            assert getattr(o, k) == getattr(r, k), (n, k, getattr(o, k), getattr(r, k))
        z = zl / obu
        regimes["very_unstable"] += z < -1.574
        regimes["unstable"] += -1.574 <= z < 0
        regimes["stable"] += 0 <= z <= 1
        regimes["very_stable"] += z > 1
        regimes["tall"] += zl - z0m > 10.0
        regimes["same_hgt"] += (hgt_q == hgt_t and z0q == z0h)
        # MoninObukIni
        thv, dthv = float(rng.uniform(240, 320)), float(rng.uniform(-12, 12) if n % 50 else 0.0)
        zms = float(rng.choice([0.5, 2.0]))
        um_o, obu_o = d(), d()
        OL.oracle_moninobukini(zms, ur, thv, dthv, zl, z0m, C.byref(um_o), C.byref(obu_o))
        assert (um_o.value, obu_o.value) == cp.monin_obuk_ini(zms, ur, thv, dthv, zl, z0m)
        # QSat
        T, p = float(rng.uniform(180, 390)), float(rng.uniform(5.0e4, 1.05e5))
        if n % 101 == 0:
            T = 273.15
        qs, es, qsdT = d(), d(), d()
        OL.oracle_qsat(T, p, C.byref(qs), C.byref(es), C.byref(qsdT))
        assert (qs.value, es.value, qsdT.value) == cp.qsat(T, p)[:3]
    assert min(regimes.values()) > 300, regimes


def canopy_patch_inputs(S, p):
    """everything CanopyFluxes reads for patch p (0-based) and its column / gridcell, Fortran-indexed, on top of the PHS inputs"""
    from tests.test_oracle_phs import phs_patch_inputs
    n = S["itype"].shape[0]
    dummy = {k: np.zeros(n) for k in ("esat_tv", "eair", "oair", "cair", "rb", "dayl_factor", "qsatl", "qaf", "bsun", "bsha")}
    P = phs_patch_inputs(S, p, dummy)
    c, g, t = P.c, P.g, P.t
    fp, fc, fg, ft = (lambda k: float(S[k][p])), (lambda k: float(S[k][c])), (lambda k: float(S[k][g])), (lambda k: float(S[k][t]))
    grnd = lambda k, lo=1: {j: float(S[k][j - lo, c]) for j in range(1, 26)}
    P.__dict__.update(
        dayl=fg("dayl"), max_dayl=fg("max_dayl"), forc_u=fg("forc_u"), forc_v=fg("forc_v"), forc_pco2=fg("forc_pco2"),
        forc_po2=fg("forc_po2"), forc_hgt_t=fg("forc_hgt_t"), forc_hgt_u=fg("forc_hgt_u"), forc_hgt_q=fg("forc_hgt_q"),
        forc_lwrad=fc("forc_lwrad"), forc_th=fc("forc_th"), forc_q=fc("forc_q"), z0mg=fc("z0mg"), t_h2osfc=fc("t_h2osfc"),
        t_grnd=fc("t_grnd"), thv=fc("thv"), emg=fc("emg"), frac_h2osfc=fc("frac_h2osfc"), frac_sno=fc("frac_sno_eff"),
        snow_depth=fc("snow_depth"), qg_snow=fc("qg_snow"), qg_soil=fc("qg_soil"), qg_h2osfc=fc("qg_h2osfc"), qg=fc("qg"),
        dqgdT=fc("dqgdT"), htvp=fc("htvp"), snl=int(S["snl"][c]), soilresis=fc("soilresis"), soilbeta=fc("soilbeta"),
        t_soisno={j: float(S["t_soisno"][j + 11, c]) for j in range(-11, 26)}, h2osoi_ice=grnd("h2osoi_ice", -11),
        h2osoi_liq=grnd("h2osoi_liq", -11), watsat=grnd("watsat"), bsw=grnd("bsw"), sucsat=grnd("sucsat"), dz=grnd("dz", -11),
        rootfr={j: float(S["rootfr"][j - 1, p]) for j in range(1, 26)},
        frac_veg_nosno=int(S["frac_veg_nosno"][p]), thm=fp("thm"), sabv=fp("sabv"), emv=fp("emv"), fwet=fp("fwet"), t_stem=fp("t_stem"),
        displa=fp("displa"), z0mv=fp("z0mv"), snocan=fp("snocan"), liqcan=fp("liqcan"), cgrnds=fp("cgrnds"), cgrndl=fp("cgrndl"),
        stem_biomass=fp("stem_biomass"), leaf_biomass=fp("leaf_biomass"), dleaf=ft("pft_dleaf"), dbh_param=ft("pft_dbh"),
        fbw=ft("pft_fbw"), nstem=ft("pft_nstem"), rstem_per_dbh=ft("pft_rstem_per_dbh"), wood_density=ft("pft_wood_density"),
        is_tree=bool(S["pft_is_tree"][t]), is_shrub=bool(S["pft_is_shrub"][t]), z0v_Cr=ft("pft_z0v_Cr"), z0v_Cs=ft("pft_z0v_Cs"),
        z0v_c=ft("pft_z0v_c"), z0v_cw=ft("pft_z0v_cw"), z0v_LAImax=ft("pft_z0v_LAImax"), smpso=ft("pft_smpso"), smpsc=ft("pft_smpsc"))
    return P


def _canopy_pin(oracle_lib, seed, ngrid, npatch, crop_every=0, **switches):
    from types import SimpleNamespace
    from ctsm_b200 import abi, synthetic_canopy
    from tests import phs_python as pp
    from tests.test_oracle_phs import phs_output_pairs
    from tests.util import copy_state
    sg, S = synthetic_canopy.make_full_case(ngrid, seed=seed)
    prm = abi.default_params()
    for k, v in switches.items():
        setattr(prm, k, v)
    if crop_every:
        S["pft_crop"][np.unique(S["itype"])[1::crop_every]] = 1.0
    S0 = copy_state(S)
    fe = sg.filters["exposedvegp"]
    f = abi.make_struct("canopyfluxes", S, sg.bounds)
    st = abi.Status()
    oracle_lib.oracle_canopyfluxes.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Bounds), C.c_int, C.POINTER(C.c_int32),
                                               C.POINTER(abi.STRUCTS["canopyfluxes"]), C.POINTER(abi.Status)]
    assert oracle_lib.oracle_canopyfluxes(C.byref(prm), C.byref(sg.bounds), len(fe), abi.i32p(fe), C.byref(f), C.byref(st)) == 0
    M = SimpleNamespace(**{k: getattr(prm, k) for k, _ in abi.Params._fields_ if not k.startswith("reserved")})
    patch_fields = ("t_veg t_stem displa z0mv z0hv z0qv forc_hgt_u_patch forc_hgt_t_patch forc_hgt_q_patch snocan liqcan cgrnds cgrndl "
                    "cgrnd qflx_tran_veg stem_biomass leaf_biomass dleaf_patch dhsdt_canopy btran taux tauy dlrad ulrad eflx_sh_snow "
                    "eflx_sh_h2osfc eflx_sh_soil eflx_sh_stem eflx_sh_veg eflx_sh_grnd ram1 rb1 rah1 rah2 raw1 raw2 ustar um uaf taf qaf "
                    "obu zeta vpd u10 u10_clm fv va vds t_ref2m t_ref2m_r t_skin q_ref2m rh_ref2m rh_ref2m_r rh_af vpd_ref2m iwue_ln "
                    "qflx_evap_veg qflx_evap_soi qflx_ev_snow qflx_ev_soil qflx_ev_h2osfc fpsn fpsn_wc fpsn_wj fpsn_wp").split()
    stats = {"patches": 0, "iters": 0, "max_iter": 0, "capped": 0, "clipped": 0}
    for p1 in fe[:npatch]:
        p = int(p1) - 1
        P = canopy_patch_inputs(S0, p)
        P.gs_mol_patch = float(S0["gs_mol"][0, p])
        O = cp.canopy_fluxes_patch(P, M, pp.photosynthesis_hydraulic_stress if prm.use_hydrstress else pp.photosynthesis)
        got = lambda k, *i: float(S[k][(*i, p)])
        pairs = [(k, getattr(O, k), got(k)) for k in patch_fields]
        pairs += [("num_iter", O.num_iter, int(S["num_iter"][p]))]
        day = P.par_z[1] > 0.0
        if prm.use_hydrstress:
            pairs += [("bsun", O.phs.bsun, got("bsun")), ("bsha", O.phs.bsha, got("bsha"))]
            pairs += phs_output_pairs(O.phs, got, day, prm.stomatalcond_mtd)
        else:
            for W, sfx in ((O.psn_sun, "sun"), (O.psn_sha, "sha")):
                pairs += [("lmr%s_z" % sfx, W.lmr_z, got("lmr%s_z" % sfx, 0)), ("psn%s_z" % sfx, W.psn_z, got("psn%s_z" % sfx, 0)),
                          ("rs%s_z" % sfx, W.rs_z, got("rs%s_z" % sfx, 0)), ("ci%s_z" % sfx, W.ci_z, got("ci%s_z" % sfx, 0)),
                          ("gs_mol_" + sfx, W.gs_mol_phase, got("gs_mol_" + sfx, 0)), ("psn" + sfx, W.psn, got("psn" + sfx)),
                          ("psn%s_wc" % sfx, W.psn_wc, got("psn%s_wc" % sfx)), ("psn%s_wj" % sfx, W.psn_wj, got("psn%s_wj" % sfx)),
                          ("psn%s_wp" % sfx, W.psn_wp, got("psn%s_wp" % sfx)), ("lmr" + sfx, W.lmr, got("lmr" + sfx)),
                          ("rs" + sfx, W.rs, got("rs" + sfx))]
                if W.par_z > 0.0:
                    pairs.append(("gs_mol_%s_ln" % sfx, W.gs_mol_ln, got("gs_mol_%s_ln" % sfx, 0)))
            W = O.psn_sha                                                 # the arrays both phases share hold the shaded call's values
            pairs += [("ac", W.ac, got("ac", 0)), ("aj", W.aj, got("aj", 0)), ("ap", W.ap, got("ap", 0)), ("ag", W.ag, got("ag", 0)),
                      ("an", W.an, got("an", 0)), ("vcmax_z", W.vcmax_z, got("vcmax_z", 0)), ("tpu_z", W.tpu_z, got("tpu_z", 0)),
                      ("kp_z", W.kp_z, got("kp_z", 0)), ("c3flag", float(W.c3flag), got("c3flag")), ("qe", W.qe, got("qe")),
                      ("kc", W.kc, got("kc")), ("ko", W.ko, got("ko")), ("cp", W.cp, got("cp")), ("lnca", W.lnc, got("lnca")),
                      ("gb_mol", W.gb_mol, got("gb_mol")),
                      # luvcmax25top / lujmax25top / lutpu25top are written by PhotosynthesisHydraulicStress only (:3266-3268)
                      ("luvcmax25top", float(S0["luvcmax25top"][p]), got("luvcmax25top"))]
            if W.par_z > 0.0:
                pairs.append(("gs_mol", W.gs_mol, got("gs_mol", 0)))
                if prm.stomatalcond_mtd == 2:
                    pairs.append(("vpd_can", W.vpd_can, got("vpd_can")))
            stats["brent"] = stats.get("brent", 0) + O.psn_sun.brent_calls + O.psn_sha.brent_calls
        for j in range(1, 26):
            pairs += [("rootr", O.rootr[j], got("rootr", j - 1)), ("eff_porosity", O.eff_porosity[j], float(S["eff_porosity"][j - 1, P.c])),
                      ("h2osoi_liqvol", O.h2osoi_liqvol[j], float(S["h2osoi_liqvol"][j + 11, P.c]))]
            if O.rresis[j] is not None:
                pairs.append(("rresis", O.rresis[j], got("rresis", j - 1)))
        bad = [(k, a, b) for k, a, b in pairs if a != b]
        assert not bad, (int(p1), day, O.num_iter, bad[:8])
        stats["patches"] += 1
        stats["iters"] += O.num_iter
        stats["max_iter"] = max(stats["max_iter"], O.num_iter)
        stats["capped"] += O.num_iter > prm.itmax_canopy_fluxes
        stats["crop"] = stats.get("crop", 0) + (P.crop != 0)
    return stats


def test_canopyfluxes_matches_python_restatement(oracle_lib):
    """CanopyFluxes with plant hydraulic stress, clm6_0 switches (Meier2022 roughness, biomass heat storage, SL14 soil resistance,
    Medlyn): every patch / column array element the routine writes, identical bits, with identical iteration counts"""
    stats = _canopy_pin(oracle_lib, 1401, 400, 900)
    print("CanopyFluxes pin:", stats)
    assert stats["patches"] == 900 and stats["max_iter"] >= 10


def test_canopyfluxes_matches_python_restatement_other_switches(oracle_lib):
    """ZengWang2007 roughness, no biomass heat storage, Lee-Pielke soil beta, under-canopy stability, Ball-Berry"""
    stats = _canopy_pin(oracle_lib, 1402, 300, 500, z0param_method=1, use_biomass_heat_storage=0, soil_resis_method=0,
                        use_undercanopy_stability=1, stomatalcond_mtd=1)
    print("CanopyFluxes pin (other switches):", stats)
    assert stats["patches"] == 500


@pytest.mark.parametrize("hyd", [1, 0])
def test_canopyfluxes_matches_python_restatement_photosynthesis_switches(oracle_lib, hyd):
    """use_luna, light_inhibit and modifyphoto_and_lmr_forcrop off, crop types present: the prescribed-Vcmax, uninhibited-respiration
    and crop branches of both photosynthesis routines inside the restated CanopyFluxes"""
    stats = _canopy_pin(oracle_lib, 1410 + hyd, 250, 450, use_hydrstress=hyd, use_luna=0, light_inhibit=0, modifyphoto_and_lmr_forcrop=0,
                        crop_every=3)
    print("CanopyFluxes pin (photosynthesis switches, use_hydrstress=%d):" % hyd, stats)
    assert stats["patches"] == 450 and stats["crop"] > 30


@pytest.mark.parametrize("mtd", [2, 1])
def test_canopyfluxes_without_hydraulic_stress_matches_python_restatement(oracle_lib, mtd):
    """use_hydrstress = .false.: Photosynthesis / hybrid / brent / ci_func for sunlit then shaded leaves, btran from
    calc_root_moist_stress, transpiration from the potential evaporation - identical bits, Medlyn and Ball-Berry"""
    stats = _canopy_pin(oracle_lib, 1403 + mtd, 300, 600, use_hydrstress=0, stomatalcond_mtd=mtd)
    print("CanopyFluxes pin (no PHS, stomatalcond_mtd=%d):" % mtd, stats)
    assert stats["patches"] == 600 and stats["brent"] > 0


def test_vert_tran_sink_hydstress_matches_python_restatement(oracle_lib):
    """Compute_EffecRootFrac_And_VertTranSink_HydStress (SoilWaterPlantSinkMod.F90:236-328), written from the Fortran column by
    column: qflx_rootsoi, qflx_phs_neg and qflx_hydr_redist, identical bits"""
    from ctsm_b200 import abi, synthetic_canopy
    from tests.util import copy_state
    sg, S = synthetic_canopy.make_full_case(300, seed=1501)
    rng = np.random.Generator(np.random.PCG64(1502))
    synthetic_canopy.balance_state(sg, S, rng, 1e-11)
    S["k_soil_root"] = rng.uniform(1.0e-9, 2.0e-6, S["k_soil_root"].shape) * (rng.random(S["k_soil_root"].shape) < 0.9)
    S["vegwp"][3] = rng.uniform(-250000.0, -2000.0, S["vegwp"].shape[1])
    S["wtcol"][::23] = 0.0
    S0 = copy_state(S)
    fh = sg.filters["hydrologyc"]
    f = abi.make_struct("plantsink", S, sg.bounds)
    oracle_lib.oracle_vert_tran_sink_hydstress.argtypes = [C.POINTER(abi.Bounds), C.c_int, C.POINTER(C.c_int32),
                                                          C.POINTER(abi.STRUCTS["plantsink"])]
    assert oracle_lib.oracle_vert_tran_sink_hydstress(C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(f)) == 0
    neg = 0
    for c1 in fh:
        c = int(c1) - 1
        qflx_phs_neg = 0.0
        redist = {}
        for j in range(1, 21):
            grav2 = float(S0["z"][j + 11, c]) * 1000.0
            temp = 0.0
            for p1 in range(int(S0["patchi"][c]), int(S0["patchi"][c]) + int(S0["npatches"][c])):
                p = p1 - 1
                if j == 1:
                    redist[p] = 0.0
                if S0["patch_active"][p] and S0["frac_veg_nosno"][p] > 0:
                    if S0["wtcol"][p] > 0.0:
                        patchflux = float(S0["k_soil_root"][j - 1, p]) * (float(S0["smp_l"][j - 1, c]) - float(S0["vegwp"][3, p]) - grav2)
                        if patchflux < 0:
                            redist[p] = redist[p] + patchflux
                        temp = temp + patchflux * float(S0["wtcol"][p])
            assert S["qflx_rootsoi"][j - 1, c] == temp, (c1, j)
            if temp < 0.0:
                qflx_phs_neg = qflx_phs_neg + temp
        assert S["qflx_phs_neg"][c] == qflx_phs_neg
        neg += qflx_phs_neg < 0.0
        for p, v in redist.items():
            assert S["qflx_hydr_redist"][p] == v
    assert neg > 20


@pytest.mark.parametrize("with_urban", [False, True])
def test_balancecheck_residuals_match_numpy(oracle_lib, with_urban):
    """the residual formulas of BalanceCheck / EnergyBalanceCheck (BalanceCheckMod.F90:565-600 errh2o_col, :648-694 c2g + errh2o_grc,
    :746-800 snow sources / sinks / errh2osno, :948-1000 errsol / errlon / errseb / netrad), restated in NumPy from the Fortran"""
    from ctsm_b200 import abi, synthetic_canopy
    from oracle import oracle
    from tests.util import copy_state
    sg, S = synthetic_canopy.make_full_case(500, seed=1601)
    rng = np.random.Generator(np.random.PCG64(1602))
    synthetic_canopy.balance_state(sg, S, rng, 1e-6)
    S["lun_itype"][::13] = 5                                       # some deep-lake columns: the other snow-source form
    if with_urban:
        S["lun_itype"][::17] = 7                                   # an urban type: the first (generic) snow form
    S0 = copy_state(S)
    prm = abi.default_params()
    prm.balance_skip_steps = 3                                     # BalanceCheckInit at dtime = 1800 s (test_Balance.pf)
    L = oracle.lib()
    rep, st = abi.BalanceReport(), abi.Status()
    allc = np.arange(1, sg.ncol + 1, dtype=np.int32)
    f = abi.make_struct("balancecheck", S, sg.bounds)
    rc = L.oracle_balancecheck(C.byref(prm), C.byref(sg.bounds), len(allc), abi.i32p(allc), C.byref(f), 1, C.byref(rep), C.byref(st))
    assert rc == 0, st.msg
    I, dtime = S0, prm.dtime
    act = I["col_active"] != 0
    errh2o = np.where(act, I["endwb"] - I["begwb"] - (I["forc_rain"] + I["forc_snow"] + I["qflx_flood"] + I["qflx_sfc_irrig"]
                      + I["qflx_glcice_dyn_water_flux"] - I["qflx_evap_tot"] - I["qflx_surf"] - I["qflx_qrgwl"] - I["qflx_drain"]
                      - I["qflx_drain_perched"] - I["qflx_ice_runoff"] - I["qflx_snwcp_discarded_liq"]
                      - I["qflx_snwcp_discarded_ice"]) * dtime, 0.0)
    assert np.array_equal(S["errh2o"], errh2o)
    # snow balance
    nlevsno = prm.nlevsno
    lev = np.arange(-nlevsno + 1, 1)[:, None]
    insnow = lev >= (I["snl"] + 1)[None, :]
    h2osno_total = I["h2osno_no_layers"].copy()
    for j in range(nlevsno):                                       # layer order of CalculateTotalH2osno (WaterStateType.F90)
        h2osno_total = np.where(insnow[j], h2osno_total + I["h2osoi_ice"][j] + I["h2osoi_liq"][j], h2osno_total)
    lt = I["lun_itype"]
    src = I["qflx_prec_grnd"] + I["qflx_soliddew_to_top_layer"] + I["qflx_liqdew_to_top_layer"]
    snk = (I["qflx_solidevap_from_top_layer"] + I["qflx_liqevap_from_top_layer"] + I["qflx_snow_drain"] + I["qflx_snwcp_ice"]
           + I["qflx_snwcp_liq"] + I["qflx_snwcp_discarded_ice"] + I["qflx_snwcp_discarded_liq"] + I["qflx_sl_top_soil"])
    lak = lt == 5
    src = np.where(lak, I["qflx_snow_grnd"] + I["frac_sno_eff"] * (I["qflx_liq_grnd"] + I["qflx_soliddew_to_top_layer"]
                                                                   + I["qflx_liqdew_to_top_layer"]), src)
    snk = np.where(lak, I["frac_sno_eff"] * (I["qflx_solidevap_from_top_layer"] + I["qflx_liqevap_from_top_layer"]) + I["qflx_snwcp_ice"]
                   + I["qflx_snwcp_liq"] + I["qflx_snwcp_discarded_ice"] + I["qflx_snwcp_discarded_liq"] + I["qflx_snow_drain"]
                   + I["qflx_sl_top_soil"], snk)
    soil = (lt == 1) | (lt == 2) | (lt == 6) | (lt == 4)
    src = np.where(soil, (I["qflx_snow_grnd"] - I["qflx_snow_h2osfc"]) + I["frac_sno_eff"] * (I["qflx_liq_grnd"]
                   + I["qflx_soliddew_to_top_layer"] + I["qflx_liqdew_to_top_layer"]) + I["qflx_h2osfc_to_ice"], src)
    snk = np.where(soil, I["frac_sno_eff"] * (I["qflx_solidevap_from_top_layer"] + I["qflx_liqevap_from_top_layer"]) + I["qflx_snwcp_ice"]
                   + I["qflx_snwcp_liq"] + I["qflx_snwcp_discarded_ice"] + I["qflx_snwcp_discarded_liq"] + I["qflx_snow_drain"]
                   + I["qflx_sl_top_soil"], snk)
    has = act & (I["snl"] < 0)
    assert np.array_equal(S["snow_sources"][act], np.where(has, src, 0.0)[act])
    assert np.array_equal(S["snow_sinks"][act], np.where(has, snk, 0.0)[act])
    assert np.array_equal(S["errh2osno"], np.where(has, (h2osno_total - I["h2osno_old"]) - (src - snk) * dtime, 0.0))
    assert has.sum() > 50 and (lak & has).sum() > 3 and (not with_urban or ((lt == 7) & has).sum() > 3)
    # energy
    pa = I["patch_active"] != 0
    c, g = I["column"] - 1, I["gridcell"] - 1
    urb = (lt[c] >= 7) & (lt[c] <= 9)
    errsol = I["fsa"] + I["fsr"] - (I["forc_solad"][0, c] + I["forc_solad"][1, c] + I["forc_solai"][0, g] + I["forc_solai"][1, g])
    errlon = I["eflx_lwrad_out"] - I["eflx_lwrad_net"] - I["forc_lwrad"][c]
    errseb = (I["sabv"] + I["sabg_chk"] + I["forc_lwrad"][c] - I["eflx_lwrad_out"] - I["eflx_sh_tot"] - I["eflx_lh_tot"]
              - I["eflx_soil_grnd"] - I["dhsdt_canopy"])
    sel = pa & ~urb
    assert np.array_equal(S["errsol"][sel], errsol[sel]) and np.array_equal(S["errlon"][sel], errlon[sel])
    assert np.array_equal(S["errseb"][sel], errseb[sel])
    assert np.array_equal(S["netrad"][pa], (I["fsa"] - I["eflx_lwrad_net"])[pa])
    assert np.all(S["errsol"][~pa] == 0.0) and np.all(S["errseb"][~pa] == 0.0)
    # gridcell residual with the three c2g averages ('urbanf' / 'unity' scales are 1 off urban landunits)
    ng = I["begwb_grc"].shape[0]
    def c2g(x):
        out, sumwt = np.full(ng, 1.0e36), np.zeros(ng)
        for ci in range(x.shape[0]):
            if act[ci] and I["wtgcell"][ci] != 0.0 and x[ci] != 1.0e36:
                gi = sg.col_gridcell[ci] - 1
                if sumwt[gi] == 0.0:
                    out[gi] = 0.0
                out[gi] = out[gi] + x[ci] * 1.0 * 1.0 * I["wtgcell"][ci]
                sumwt[gi] = sumwt[gi] + I["wtgcell"][ci]
        nz = sumwt != 0.0
        out[nz] = out[nz] / sumwt[nz]
        return out
    if not with_urban:
        errg = I["endwb_grc"] - I["begwb_grc"] - (I["forc_rain_grc"] + I["forc_snow_grc"] + I["forc_flood_grc"] + I["qflx_sfc_irrig_grc"]
               + c2g(I["qflx_glcice_dyn_water_flux"]) - I["qflx_evap_tot_grc"] - I["qflx_surf_grc"] - I["qflx_qrgwl_grc"]
               - I["qflx_drain_grc"] - I["qflx_drain_perched_grc"] - I["qflx_ice_runoff_grc"] - c2g(I["qflx_snwcp_discarded_liq"])
               - c2g(I["qflx_snwcp_discarded_ice"])) * dtime
        assert np.array_equal(S["errh2o_grc"], errg)


@pytest.mark.parametrize("method,resis", [(2, 1), (1, 0)], ids=["meier2022_sl14", "zengwang2007_leepielke"])
def test_bare_ground_fluxes_matches_python_restatement(oracle_lib, method, resis):
    """BareGroundFluxes (three Monin-Obukhov passes, scalar roughness of the ground, the dew-point switch of the latent-heat
    conductance, fluxes and 2 m diagnostics) for every patch without exposed vegetation: identical bits; the column's z0hg / z0qg
    take the value of the column's last such patch"""
    from types import SimpleNamespace
    from ctsm_b200 import abi
    from tests.test_oracle_preflux import case, run_preflux, run_humidity, run_bare
    from tests.util import copy_state
    sg, S = case(n=500, seed=1701, wet_every=3)
    prm = abi.default_params()
    prm.z0param_method, prm.soil_resis_method = method, resis
    assert run_preflux(oracle_lib, prm, sg, S) == 0
    assert run_humidity(oracle_lib, sg, S) == 0
    rng = np.random.Generator(np.random.PCG64(1702))
    S["t_grnd"] = S["t_grnd"] + rng.uniform(-6.0, 3.0, S["t_grnd"].shape)      # both sides of the dew point and of freezing
    S0 = copy_state(S)
    assert run_bare(oracle_lib, prm, sg, S) == 0
    M = SimpleNamespace(**{k: getattr(prm, k) for k, _ in abi.Params._fields_ if not k.startswith("reserved")})
    fields = ("btran t_veg rssun rssha displa z0mv z0hv z0qv dlrad ulrad dhsdt_canopy eflx_sh_stem z0mg_p z0hg_p z0qg_p kbm1 um obu zeta "
              "ustar vds u10 u10_clm va fv ram1 cgrnds cgrndl cgrnd taux tauy eflx_sh_grnd eflx_sh_tot eflx_sh_snow eflx_sh_soil "
              "eflx_sh_h2osfc qflx_tran_veg qflx_evap_veg qflx_evap_soi qflx_evap_tot_patch qflx_ev_snow qflx_ev_soil qflx_ev_h2osfc "
              "t_ref2m q_ref2m rh_ref2m forc_hgt_u_patch forc_hgt_t_patch forc_hgt_q_patch").split()
    fp = sg.filters["noexposedvegp"]
    last = {}
    stats = {"dew": 0, "nodew": 0, "stable": 0, "unstable": 0}
    for p1 in fp:
        p = int(p1) - 1
        c, g = int(S0["column"][p]) - 1, int(S0["gridcell"][p]) - 1
        fc, fg, fpp = (lambda k: float(S0[k][c])), (lambda k: float(S0[k][g])), (lambda k: float(S0[k][p]))
        P = SimpleNamespace(
            forc_u=fg("forc_u"), forc_v=fg("forc_v"), forc_hgt_u=fg("forc_hgt_u"), forc_hgt_t=fg("forc_hgt_t"), forc_hgt_q=fg("forc_hgt_q"),
            snl=int(S0["snl"][c]), forc_t=fc("forc_t"), forc_th=fc("forc_th"), forc_q=fc("forc_q"), forc_pbot=fc("forc_pbot"),
            forc_rho=fc("forc_rho"), t_grnd=fc("t_grnd"), t_h2osfc=fc("t_h2osfc"), thv=fc("thv"), beta=fc("beta"), zii=fc("zii"),
            qg=fc("qg"), qg_snow=fc("qg_snow"), qg_soil=fc("qg_soil"), qg_h2osfc=fc("qg_h2osfc"), dqgdT=fc("dqgdT"), htvp=fc("htvp"),
            soilbeta=fc("soilbeta"), soilresis=fc("soilresis"), z0mg=fc("z0mg"),
            z0hg=fc("z0hg"), z0qg=fc("z0qg"),                  # every patch starts from the column's value at entry (:309-311)
            t_soisno={j: float(S0["t_soisno"][j + 11, c]) for j in range(-11, 26)}, thm=fpp("thm"),
            forc_hgt_u_patch=fpp("forc_hgt_u_patch"), forc_hgt_t_patch=fpp("forc_hgt_t_patch"), forc_hgt_q_patch=fpp("forc_hgt_q_patch"))
        O = cp.bare_ground_fluxes_patch(P, M)
        bad = [(k, getattr(O, k), float(S[k][p])) for k in fields if getattr(O, k) != float(S[k][p])]
        assert not bad, (int(p1), bad[:6])
        assert S["num_iter"][p] == O.num_iter
        assert np.all(S["rootr"][:, p] == 0.0) and np.all(S["rresis"][:, p] == 0.0)
        if S0["lun_itype"][c] in (1, 2):
            assert S["t_ref2m_r"][p] == O.t_ref2m and S["rh_ref2m_r"][p] == O.rh_ref2m
        last[c] = (O.z0hg_p, O.z0qg_p)
        stats["dew"] += O.cgrndl != 0.0
        stats["nodew"] += O.cgrndl == 0.0
        stats["stable"] += O.zeta > 0
        stats["unstable"] += O.zeta < 0
    for c, (zh, zq) in last.items():
        assert S["z0hg"][c] == zh and S["z0qg"][c] == zq
    assert min(stats["stable"], stats["unstable"]) > 50 and (resis == 1 or min(stats["dew"], stats["nodew"]) > 20), stats
