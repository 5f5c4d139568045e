"""CPU pin of the oracle's plant-hydraulic-stress photosynthesis (oracle/oracle_phs.c) by tests/phs_python.py, an independent
per-patch Python restatement written from PhotosynthesisMod.F90:2704-5228.  Same libm, same operation order: identical bits."""
import ctypes as C
from types import SimpleNamespace

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


def phs_patch_inputs(S, p, work):
    """everything PhotosynthesisHydraulicStress reads for patch p (0-based) of the state S, Fortran-indexed"""
    P = patch_inputs(S, None, p, None)
    c, g, t = P.c, P.g, P.t
    f = lambda k: float(S[k][p])
    soil = lambda a, i, lo: {j: float(a[j - lo, i]) for j in range(1, 21)}
    P.__dict__.update(
        froot_carbon=f("froot_carbon"), rootfr=soil(S["rootfr"], p, 1), dz=soil(S["dz"], c, -11), hksat=soil(S["hksat"], c, 1),
        hk_l=soil(S["hk_l"], c, 1), root_radius=float(S["pft_root_radius"][t]), root_density=float(S["pft_root_density"][t]),
        froot_leaf=float(S["pft_froot_leaf"][t]), krmax=float(S["pft_krmax"][t]), tlai=f("tlai"), c3psn=float(S["pft_c3psn"][t]),
        crop=float(S["pft_crop"][t]), leafcn=float(S["pft_leafcn"][t]), flnr=float(S["pft_flnr"][t]), fnitr=float(S["pft_fnitr"][t]),
        slatop=float(S["pft_slatop"][t]), mbbopt=float(S["pft_mbbopt"][t]), medlynintercept=float(S["pft_medlynintercept"][t]),
        medlynslope=float(S["pft_medlynslope"][t]), theta_cj=float(S["pft_theta_cj"][t]), t_veg=f("t_veg"), t10=f("t_a10"),
        nrad=int(S["nrad"][p]), tlai_z=float(S["tlai_z"][0, p]), vcmaxcintsun=f("vcmaxcintsun"), vcmaxcintsha=f("vcmaxcintsha"),
        vcmx25_z=float(S["vcmx25_z"][0, p]), jmx25_z=float(S["jmx25_z"][0, p]),
        par_z={1: float(S["parsun_z"][0, p]), 2: float(S["parsha_z"][0, p])},
        lai_z={1: float(S["laisun_z"][0, p]), 2: float(S["laisha_z"][0, p])},
        o3coefv={1: f("o3coefvsun"), 2: f("o3coefvsha")}, o3coefg={1: f("o3coefgsun"), 2: f("o3coefgsha")},
        vegwp={i: float(S["vegwp"][i - 1, p]) for i in range(1, 5)},
        gs_mol={1: float(S["gs_mol_sun"][0, p]), 2: float(S["gs_mol_sha"][0, p])},
        near_local_noon=bool(S["near_local_noon"][g]), bsun_in=float(work["bsun"][p]), bsha_in=float(work["bsha"][p]),
        **{k: float(work[k][p]) for k in ("esat_tv", "eair", "oair", "cair", "rb", "dayl_factor", "qsatl", "qaf")})
    return P


def phs_output_pairs(W, got, day, mtd):
    """(name, restatement, oracle) for every array element PhotosynthesisHydraulicStress writes for one patch; got(name, *level)
    reads the oracle's state"""
    pairs = [("c3flag", float(W.c3flag), got("c3flag")), ("qe", W.qe, got("qe")), ("kc", W.kc, got("kc")), ("ko", W.ko, got("ko")),
             ("cp", W.cp, got("cp")), ("lnca", W.lnc, got("lnca")), ("luvcmax25top", W.luvcmax25top, got("luvcmax25top")),
             ("lujmax25top", W.lujmax25top, got("lujmax25top")), ("lutpu25top", W.lutpu25top, got("lutpu25top")),
             ("gb_mol", W.gb_mol, got("gb_mol")), ("qflx_tran_veg", W.qflx_tran_veg, got("qflx_tran_veg"))]
    for j in range(1, 21):
        pairs += [("k_soil_root", W.k_soil_root[j], got("k_soil_root", j - 1)),
                  ("root_conductance", W.root_conductance[j], got("root_conductance", j - 1)),
                  ("soil_conductance", W.soil_conductance[j], got("soil_conductance", j - 1))]
    for i in range(1, 5):
        pairs += [("vegwp", W.vegwp[i], got("vegwp", i - 1)), ("vegwp_pd", W.vegwp_pd[i], got("vegwp_pd", i - 1))]
        if day:
            pairs.append(("vegwp_ln", W.vegwp_ln[i], got("vegwp_ln", i - 1)))
    for s, sfx in ((1, "sun"), (2, "sha")):
        pairs += [("ac_phs", W.ac[s], got("ac_phs", s - 1)), ("aj_phs", W.aj[s], got("aj_phs", s - 1)),
                  ("ap_phs", W.ap[s], got("ap_phs", s - 1)), ("ag_phs", W.ag[s], got("ag_phs", s - 1)),
                  ("vcmax_z_phs", W.vcmax_z[s], got("vcmax_z_phs", s - 1)), ("tpu_z_phs", W.tpu_z[s], got("tpu_z_phs", s - 1)),
                  ("kp_z_phs", W.kp_z[s], got("kp_z_phs", s - 1)), ("an_" + sfx, W.an[s], got("an_" + sfx, 0)),
                  ("lmr%s_z" % sfx, W.lmr_z[s], got("lmr%s_z" % sfx, 0)), ("psn%s_z" % sfx, W.psn_z[s], got("psn%s_z" % sfx, 0)),
                  ("rs%s_z" % sfx, W.rs_z[s], got("rs%s_z" % sfx, 0)), ("ci%s_z" % sfx, W.ci_z[s], got("ci%s_z" % sfx, 0)),
                  ("gs_mol_" + sfx, W.gs_mol[s], got("gs_mol_" + sfx, 0)), ("psn" + sfx, W.psn[s], got("psn" + sfx)),
                  ("psn%s_wc" % sfx, W.psn_wc[s], got("psn%s_wc" % sfx)), ("psn%s_wj" % sfx, W.psn_wj[s], got("psn%s_wj" % sfx)),
                  ("psn%s_wp" % sfx, W.psn_wp[s], got("psn%s_wp" % sfx)), ("lmr" + sfx, W.lmr[s], got("lmr" + sfx)),
                  ("rs" + sfx, W.rs[s], got("rs" + sfx))]
        if day:
            pairs.append(("gs_mol_%s_ln" % sfx, W.gs_mol_ln[s], got("gs_mol_%s_ln" % sfx, 0)))
    if day and mtd == 2:
        pairs.append(("vpd_can", W.vpd_can, got("vpd_can")))
    return pairs


@pytest.mark.parametrize("mtd,seed", [(2, 1201), (1, 1202)])
def test_photosynthesis_hydraulic_stress_matches_python_restatement(oracle_lib, mtd, seed):
    """the whole of PhotosynthesisHydraulicStress (root-soil conductances, the vcmax / jmax / tpu / lmr temperature response, the
    night branch, hybrid_PHS + brent_PHS + ci_func_PHS around calcstress, the canopy sums): every array the routine writes,
    identical bits, Medlyn and Ball-Berry"""
    OL = _bind(oracle_lib)
    sg, S = synthetic_canopy.make_full_case(900, seed=seed)
    prm = abi.default_params()
    prm.stomatalcond_mtd = mtd
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    fe = np.ascontiguousarray(sg.filters["exposedvegp"][:2400], dtype=np.int32)
    np_ = S["itype"].shape[0]
    S["pft_crop"][np.unique(S["itype"])[1::4]] = 1.0                  # crop types: the bsun-free leaf respiration branches
    dim = fe[::7] - 1                                                 # dim light: net leaf uptake below zero on one or both leaves
    S["parsun_z"][0, dim] *= 0.02
    S["parsha_z"][0, dim] *= rng.choice([0.002, 0.02, 1.0], dim.size)
    col, grc = S["column"] - 1, S["gridcell"] - 1
    tv = S["t_veg"]
    work = {"esat_tv": 611.0 * np.exp(17.27 * (tv - 273.15) / (tv - 35.85))}
    work["eair"] = work["esat_tv"] * rng.uniform(0.25, 1.1, np_)
    work["oair"] = S["forc_po2"][grc].copy()
    work["cair"] = S["forc_pco2"][grc].copy()
    work["rb"] = rng.uniform(4.0, 80.0, np_)
    work["dayl_factor"] = rng.uniform(0.05, 1.0, np_) * (rng.random(np_) > 0.03)
    work["qsatl"] = rng.uniform(0.004, 0.03, np_)
    work["qaf"] = work["qsatl"] - rng.uniform(-0.002, 0.012, np_)
    work["bsun"], work["bsha"], work["btran"] = np.full(np_, -9.0), np.full(np_, -9.0), np.full(np_, -9.0)
    S["laisha"][fe[::29] - 1] = 0.0005                                # below tol_lai: the 3x3 branch of spacA
    S["laisha_z"][0, fe[::29] - 1] = 0.0005
    S0 = copy_state(S)
    M = SimpleNamespace(**{k: getattr(prm, k) for k, _ in abi.Params._fields_ if not k.startswith("reserved")})
    f = abi.make_struct("canopyfluxes", S, sg.bounds)
    wp = lambda k: work[k].ctypes.data_as(dp)
    rc = OL.oracle_phs_standalone(C.byref(prm), C.byref(f), len(fe), fe.ctypes.data_as(C.POINTER(C.c_int32)), wp("esat_tv"), wp("eair"),
                                  wp("oair"), wp("cair"), wp("rb"), wp("bsun"), wp("bsha"), wp("btran"), wp("dayl_factor"),
                                  wp("qsatl"), wp("qaf"))
    assert rc == 0
    work0 = {k: v.copy() for k, v in work.items()}
    work0["bsun"][:] = -9.0
    work0["bsha"][:] = -9.0
    stats = {"night": 0, "day": 0, "brent": 0, "itmax": 0, "calcstress": 0, "c4": 0, "crop": 0, "an_neg": 0}
    for p1 in fe:
        p = int(p1) - 1
        P = phs_patch_inputs(S0, p, work0)
        W = pp.photosynthesis_hydraulic_stress(P, M)
        day = P.par_z[1] > 0.0
        got = lambda k, *i: float(S[k][(*i, p)])
        pairs = phs_output_pairs(W, got, day, mtd)
        pairs += [("bsun", W.bsun, work["bsun"][p]), ("bsha", W.bsha, work["bsha"][p]), ("btran", W.btran, work["btran"][p])]
        bad = [(k, a, b) for k, a, b in pairs if a != b]
        assert not bad, (int(p1), day, bad[:6])
        stats["day"] += day
        stats["night"] += not day
        stats["brent"] += W.brent_calls
        stats["itmax"] += W.itmax_exits
        stats["calcstress"] += W.calcstress_calls
        stats["c4"] += not W.c3flag
        stats["crop"] += P.crop != 0
        stats["an_neg"] += day and (W.an[1] < 0.0 or W.an[2] < 0.0)
    print("PHS pin (stomatalcond_mtd=%d):" % mtd, stats)
    assert stats["day"] > 150 and stats["night"] > 150 and stats["brent"] > 8 and stats["c4"] > 10 and stats["crop"] > 10 and stats["an_neg"] > 20, stats
