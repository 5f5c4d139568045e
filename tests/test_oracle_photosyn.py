"""CPU pins of the oracle's soil-moisture-stress configuration (use_hydrstress = .false.):
Photosynthesis / hybrid / brent / ci_func (oracle/oracle_photosyn.c) checked through relations its OUTPUTS must satisfy,
evaluated here in NumPy from the published equations, and Compute_EffecRootFrac_And_VertTranSink_Default
(oracle/oracle_balance.c) against an independent vectorised NumPy restatement, bit for bit."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from tests.util import copy_state

RGAS, TFRZ = 6.02214e26 * 1.38065e-23, 273.15


def _run(OL, prm, sg, S):
    st = abi.Status()
    f = abi.make_struct("canopyfluxes", S, sg.bounds)
    fe = sg.filters["exposedvegp"]
    rc = OL.oracle_canopyfluxes(C.byref(prm), C.byref(sg.bounds), len(fe), abi.i32p(fe), C.byref(f), C.byref(st))
    return rc, st


def _smooth_min(theta, a, b):
    """smaller root of theta x^2 - (a+b) x + a b = 0, the co-limitation of PhotosynthesisMod.F90:2612-2626"""
    bq, cq = -(a + b), a * b
    disc = np.sqrt(np.maximum(bq * bq - 4.0 * theta * cq, 0.0))
    q = -0.5 * (bq + np.where(bq >= 0.0, disc, -disc))
    r1 = q / theta
    r2 = np.where(q != 0.0, cq / np.where(q != 0.0, q, 1.0), 1.0e36)
    return np.minimum(r1, r2)


@pytest.mark.parametrize("mtd", [1, 2], ids=["ballberry", "medlyn"])
def test_nophs_photosynthesis_outputs_satisfy_the_leaf_equations(oracle_lib, mtd):
    prm = abi.default_params()
    prm.use_hydrstress, prm.stomatalcond_mtd = 0, mtd
    sg, S = synthetic_canopy.make_full_case(600, seed=91)
    S0 = copy_state(S)
    rc, st = _run(oracle_lib, prm, sg, S)
    assert rc == 0, st.msg
    # the reference's own consistency check of the Ball-Berry conductance (:1991-2004) and the canopy energy closure (CanopyFluxes :1746)
    assert st.n_warnings == 0
    fe = sg.filters["exposedvegp"] - 1
    ci_col = S["column"][fe] - 1
    gi = S["gridcell"][fe] - 1
    ivt = S["itype"][fe]
    pbot = S["forc_pbot"][ci_col]
    cair = S["forc_pco2"][gi]
    c3 = np.rint(S["pft_c3psn"][ivt]) == 1
    # ac..an hold what the LAST ci_func evaluation of the shaded phase left (the phases share the arrays)
    ac, aj, ap, ag, an = (S[k][0, fe] for k in ("ac", "aj", "ap", "ag", "an"))
    lmr = S["lmrsha_z"][0, fe]
    day = S0["parsha_z"][0, fe] > 0.0
    assert day.sum() > 100 and (~day).sum() > 100
    assert np.all(S["nrad"][fe] == 1)
    np.testing.assert_array_equal(an, ag - lmr)
    # night :1781-1815
    assert np.all(ag[~day] == 0.0) and np.all(S["psnsha_z"][0, fe][~day] == 0.0) and np.all(S["cisha_z"][0, fe][~day] == 0.0)
    cf = pbot / (RGAS * 1.0e-3 * S["thm"][fe]) * 1.0e06
    btran = S["btran"][fe]
    if mtd == 1:
        floor = np.maximum(np.where(c3, 10000.0, 40000.0) * btran, 1.0)
    else:
        floor = S["pft_medlynintercept"][ivt]
    np.testing.assert_allclose(S["rssha_z"][0, fe][~day], np.minimum(2.0e4, 1.0 / floor[~day] * cf[~day]), rtol=1e-14)
    # day: co-limited gross rate from the three limiting rates (:2612-2626)
    ai = _smooth_min(S["pft_theta_cj"][ivt], ac, aj)
    ag_np = np.maximum(0.0, _smooth_min(prm.theta_ip, ai, ap))
    np.testing.assert_allclose(ag[day], ag_np[day], rtol=1e-9, atol=1e-12)
    # C4 rates do not depend on ci (:2598-2606); C3 TPU limit (:2595)
    c4d = day & ~c3
    np.testing.assert_array_equal(ac[c4d], S["vcmax_z"][0, fe][c4d])
    np.testing.assert_allclose(aj[c4d], 0.05 * S0["parsha_z"][0, fe][c4d] * 4.6, rtol=1e-15)
    np.testing.assert_allclose(ap[day & c3], 3.0 * S["tpu_z"][0, fe][day & c3], rtol=1e-15)
    # conductance floor when the leaf respires more than it fixes (:1880-1886), and the diffusion equation for ci (:1915-1919)
    gs, gb = S["gs_mol_sha"][0, fe], S["gb_mol"][fe]
    neg = day & (an < 0.0)
    np.testing.assert_array_equal(gs[neg], floor[neg])
    assert np.all(gs[day] >= np.minimum(floor[day], 1.0) * 0.999)
    ci = np.maximum(cair - an * pbot * (1.4 * gs + 1.6 * gb) / (gb * gs), 1.0e-06)
    np.testing.assert_allclose(S["cisha_z"][0, fe][day], ci[day], rtol=1e-13)
    np.testing.assert_allclose(S["rssha_z"][0, fe][day], np.minimum(cf[day] / gs[day], 2.0e4) / S["o3coefgsha"][fe][day], rtol=1e-13)
    np.testing.assert_allclose(S["psnsha_z"][0, fe][day], (ag * S["o3coefvsha"][fe])[day], rtol=1e-15)
    # the root finder converged: the ci the last evaluation used (recovered from ac for C3 leaves limited by Rubisco) is the
    # ci the diffusion equation returns, within the solver's 1 % step tolerance (:2251-2400)
    kc, ko, cp_, vc = S["kc"][fe], S["ko"][fe], S["cp"][fe], S["vcmax_z"][0, fe]
    m = day & c3 & (an > 0.0) & (vc > 0.0) & (ac > 0.0)
    oair = S["forc_po2"][gi]
    ci_used = (ac * kc * (1.0 + oair / ko) + vc * cp_)[m] / (vc - ac)[m]
    rel = np.abs(ci_used - ci[m]) / ci[m]
    assert np.quantile(rel, 0.99) < 0.05, float(np.quantile(rel, 0.99))
    # transpiration follows the potential evaporation (:1231-1248): zero where the soil-moisture stress closes the stomata
    assert np.all(S["qflx_tran_veg"][fe][btran <= 0.0] == 0.0)
    assert np.all(S["qflx_tran_veg"][fe] >= 0.0)
    # btran of this configuration is the root-weighted soil-moisture stress (SoilMoistStressMod.F90:377-431): rootr sums to 1
    rs = S["rootr"][:, fe].sum(axis=0)
    np.testing.assert_allclose(rs[btran > 0.0], 1.0, rtol=1e-12)


def test_default_sink_matches_numpy(oracle_lib):
    prm = abi.default_params()
    prm.use_hydrstress = 0
    sg, S = synthetic_canopy.make_full_case(300, seed=92)
    synthetic_canopy.balance_state(sg, S, np.random.Generator(np.random.PCG64(93)), 1.0e-11)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(94)))
    assert _run(oracle_lib, prm, sg, S)[0] == 0
    # patches without exposed vegetation: what BareGroundFluxes leaves (BareGroundFluxesMod.F90:294, :468)
    bare = np.ones(sg.npatch, dtype=bool); bare[sg.filters["exposedvegp"] - 1] = False
    S["rootr"][:, bare] = 0.0
    S["qflx_tran_veg"][bare] = 0.0
    rng = np.random.Generator(np.random.PCG64(95))
    S["qflx_tran_veg_col"] = rng.uniform(0.0, 3.0e-5, sg.ncol)
    fh = sg.filters["hydrologyc"]
    f = abi.make_struct("plantsinkdefault", S, sg.bounds)
    S["rootr_col"][...] = 7.0
    assert oracle_lib.oracle_vert_tran_sink_default(C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(f)) == 0
    # NumPy: the k-th patch of every column at a time, so each column's patches are summed in ascending order
    cols = fh - 1
    nlevsoi = S["qflx_rootsoi"].shape[0]
    acc = np.zeros((nlevsoi, len(cols)))
    tot = np.zeros(len(cols))
    for k in range(int(S["npatches"][cols].max())):
        has = S["npatches"][cols] > k
        p = np.where(has, S["patchi"][cols] - 1 + k, 0)
        w = np.where(has & (S["patch_active"][p] != 0), 1.0, 0.0)
        use = w > 0
        acc[:, use] = acc[:, use] + (S["rootr"][:nlevsoi, p] * S["qflx_tran_veg"][p] * S["wtcol"][p])[:, use]
        tot[use] = tot[use] + (S["qflx_tran_veg"][p] * S["wtcol"][p])[use]
    rc_np = np.where(tot != 0.0, acc / np.where(tot != 0.0, tot, 1.0), acc)
    np.testing.assert_array_equal(S["rootr_col"][:nlevsoi, cols], rc_np)
    np.testing.assert_array_equal(S["qflx_rootsoi"][:, cols], rc_np * S["qflx_tran_veg_col"][cols])
    # untouched: levels below nlevsoi and columns outside the filter
    assert np.all(S["rootr_col"][nlevsoi:, :] == 7.0)
    rest = np.ones(sg.ncol, dtype=bool); rest[cols] = False
    assert np.all(S["rootr_col"][:, rest] == 7.0)
    # with transpiration, the effective root fractions of a column sum to one
    s = S["rootr_col"][:nlevsoi, cols].sum(axis=0)
    np.testing.assert_allclose(s[tot > 0.0], 1.0, rtol=1e-12)
