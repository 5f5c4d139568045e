"""CPU checks of the CanopyFluxes/PHS oracle that do not depend on the CUDA path: invariants the
reference itself enforces and decomposition independence (the reference's
ERP/PEM system tests)."""
import ctypes as C

import numpy as np

from ctsm_b200 import abi, synthetic_canopy
from oracle import oracle
from tests.util import copy_state


def _run(OL, prm, sg, S, bounds=None, filt=None):
    st = abi.Status()
    f = abi.make_struct("canopyfluxes", S, sg.bounds)
    fe = sg.filters["exposedvegp"] if filt is None else filt
    b = sg.bounds if bounds is None else bounds
    fe = np.ascontiguousarray(fe) if len(fe) else np.zeros(1, dtype=np.int32)
    n = len(sg.filters["exposedvegp"]) if filt is None else len(filt)
    rc = OL.oracle_canopyfluxes(C.byref(prm), C.byref(b), n, abi.i32p(fe), C.byref(f), C.byref(st))
    return rc, st


def test_canopy_energy_closure_and_ranges(oracle_lib):
    prm = abi.default_params()
    sg, S = synthetic_canopy.make_full_case(400, seed=21)
    rc, st = _run(oracle_lib, prm, sg, S)
    assert rc == 0
    # CanopyFluxesMod.F90:1746-1760: |err| > 0.1 W/m2 is reported; none expected on plausible inputs
    assert st.n_warnings == 0
    fe = sg.filters["exposedvegp"] - 1
    it = S["num_iter"][fe]
    assert it.min() >= 3 and it.max() <= prm.itmax_canopy_fluxes + 1          # Appendix E.1
    for k in ("btran", "bsun", "bsha"):
        assert np.all((S[k][fe] >= 0.0) & (S[k][fe] <= 1.0 + 1e-12)), k
    assert np.all(S["qflx_tran_veg"][fe] >= 0.0)
    v = S["vegwp"][:, fe]                                                   # PhotosynthesisMod.F90:4643-4645 ordering
    assert np.all(v[3] >= v[2] - 1e-6) and np.all(v[2] >= v[0] - 1e-6) and np.all(v[2] >= v[1] - 1e-6)
    assert np.all(np.abs(S["t_veg"][fe] - S["thm"][fe]) < 40.0)
    night = S["parsun_z"][0, fe] <= 0.0
    assert np.all(S["fpsn"][fe][night] == 0.0) and S["fpsn"][fe][~night].max() > 1.0
    # untouched outside the filter (except the TimeStepInit / rb1 resets)
    out = np.ones(sg.npatch, dtype=bool); out[fe] = False
    assert np.all(S["t_ref2m"][out] == 1.0e36) and np.all(S["rb1"][out] == 0.0)


def test_canopy_water_update_is_consistent(oracle_lib):
    """:1615-1632: canopy water changes by (tran - evap)*dtime unless a pool is exhausted."""
    prm = abi.default_params()
    sg, S = synthetic_canopy.make_full_case(200, seed=22)
    S0 = copy_state(S)
    rc, _ = _run(oracle_lib, prm, sg, S)
    assert rc == 0
    fe = sg.filters["exposedvegp"] - 1
    d = (S["liqcan"] + S["snocan"] - S0["liqcan"] - S0["snocan"])[fe]
    want = ((S["qflx_tran_veg"] - S["qflx_evap_veg"]) * prm.dtime)[fe]
    total0 = (S0["liqcan"] + S0["snocan"])[fe]
    ok = np.abs(d - want) <= 1e-9 * np.maximum(1.0, np.abs(want))
    clipped = (total0 + want) < 1e-9
    assert np.all(ok | clipped)
    assert np.all(S["liqcan"][fe] >= 0.0) and np.all(S["snocan"][fe] >= 0.0)


def test_canopy_decomposition_independence(oracle_lib):
    """Results are bit-identical for any partition into clumps (columns/patches are independent)."""
    prm = abi.default_params()
    sg, S = synthetic_canopy.make_full_case(120, seed=23)
    whole, parts = copy_state(S), copy_state(S)
    assert _run(oracle_lib, prm, sg, whole)[0] == 0
    clumps, keep = oracle.make_clumps(sg, 7)
    st = abi.Status()
    f = abi.make_struct("canopyfluxes", parts, sg.bounds)
    for k in clumps:
        assert oracle_lib.oracle_canopyfluxes(C.byref(prm), C.byref(k.bounds), k.num_exposedvegp, k.filter_exposedvegp,
                                              C.byref(f), C.byref(st)) == 0
    for fs in abi.FIELDS["canopyfluxes"]:
        assert np.array_equal(whole[fs.name], parts[fs.name]), fs.name


def test_qsat_and_moninobuk_sanity(oracle_lib):
    qs, es, dq = C.c_double(), C.c_double(), C.c_double()
    oracle_lib.oracle_qsat(293.15, 101325.0, C.byref(qs), C.byref(es), C.byref(dq))
    assert abs(es.value - 2339.0) < 5.0 and abs(qs.value - 0.01448) < 2e-4 and dq.value > 0
    oracle_lib.oracle_qsat(253.15, 101325.0, C.byref(qs), C.byref(es), None)
    assert abs(es.value - 103.2) < 1.0            # over ice
    um, obu = C.c_double(), C.c_double()
    oracle_lib.oracle_moninobukini(2.0, 3.0, 290.0, 1.5, 30.0, 0.5, C.byref(um), C.byref(obu))
    assert um.value == 3.0 and obu.value > 0      # stable
    oracle_lib.oracle_moninobukini(2.0, 3.0, 290.0, -1.5, 30.0, 0.5, C.byref(um), C.byref(obu))
    assert abs(um.value - np.sqrt(9.25)) < 1e-15 and obu.value < 0


def test_fast_human_stress_indices_and_luna_accumulators_match_numpy(oracle_lib):
    """the outputs CanopyFluxes writes under the default namelist besides the fluxes - HumanIndexMod's
    FAST indices (CanopyFluxesMod.F90:1550-1583) and LUNA's daily accumulators (Acc24_Climate_LUNA, LunaMod.F90:695-724) -
    restated independently in numpy from the oracle's own 2 m diagnostics."""
    import ctypes as C
    import numpy as np
    from ctsm_b200 import abi, synthetic_canopy
    from tests.util import copy_state
    sg, S = synthetic_canopy.make_full_case(300, seed=17)
    prm = abi.default_params()
    R = copy_state(S)
    st = abi.Status()
    f = abi.make_struct("canopyfluxes", R, sg.bounds)
    fe = sg.filters["exposedvegp"]
    assert oracle_lib.oracle_canopyfluxes(C.byref(prm), C.byref(sg.bounds), len(fe), abi.i32p(fe), C.byref(f), C.byref(st)) == 0
    p = fe - 1
    tc = R["t_ref2m"][p] - 273.15
    rh = R["rh_ref2m"][p]
    es = R["vpd_ref2m"][p] / np.where(rh < 100.0, 1.0 - rh / 100.0, np.nan)           # e_ref2m from the vpd the routine wrote
    ok = np.isfinite(es)
    vap = R["vap_ref2m"][p]
    assert np.allclose(vap[ok], rh[ok] / 100.0 * es[ok], rtol=1e-9)
    assert np.array_equal(R["tc_ref2m"][p], tc)
    wbt = tc * np.arctan(0.151977 * np.sqrt(rh + 8.313659)) + np.arctan(tc + rh) - np.arctan(rh - 1.676331) \
        + 0.00391838 * rh ** 1.5 * np.arctan(0.023101 * rh) - 4.686035
    assert np.allclose(R["wbt_ref2m"][p], wbt, rtol=1e-12, atol=1e-12)
    tf = tc * 9.0 / 5.0 + 32.0
    hi = np.where(tf < 68.0, tf, -42.379 + 2.04901523 * tf + 10.14333127 * rh - 0.22475541 * tf * rh - 6.83783e-3 * tf ** 2
                  - 5.481717e-2 * rh ** 2 + 1.22874e-3 * tf ** 2 * rh + 8.5282e-4 * tf * rh ** 2 - 1.99e-6 * tf ** 2 * rh ** 2)
    assert np.allclose(R["nws_hi_ref2m"][p], (hi - 32.0) * 5.0 / 9.0, rtol=1e-12, atol=1e-11)
    assert np.allclose(R["appar_temp_ref2m"][p], tc + 3.3 * vap / 1000.0 - 0.7 * R["u10_clm"][p] - 4.0, rtol=1e-12, atol=1e-12)
    assert np.allclose(R["swbgt_ref2m"][p], 0.567 * tc + 0.393 * vap / 100.0 + 3.94, rtol=1e-12, atol=1e-12)
    assert np.allclose(R["humidex_ref2m"][p], tc + 5.0 / 9.0 * (vap / 100.0 - 10.0), rtol=1e-12, atol=1e-12)
    Tc = np.minimum(tc, 50.0); rhl = np.clip(rh, 5.0, 99.0)
    dc = np.where((Tc < -20.0) | (rhl < Tc * -2.27 + 27.7), Tc, 0.5 * wbt + 0.5 * Tc)
    assert np.allclose(R["discomf_index_ref2mS"][p], dc, rtol=1e-12, atol=1e-12)
    for k in ("wbt_ref2m", "nws_hi_ref2m", "appar_temp_ref2m", "swbgt_ref2m", "humidex_ref2m", "discomf_index_ref2mS"):
        assert np.array_equal(R[k + "_r"][p], R[k][p])
    # LUNA accumulators
    live = S["t_veg_day"][p] != 1.0e36
    day = S["sabv"][p] > 0
    assert live.any() and (~live).any() and day.any() and (~day).any()
    tv = R["t_veg"][p]
    assert np.array_equal(R["t_veg_day"][p], np.where(live & day, S["t_veg_day"][p] + tv, S["t_veg_day"][p]))
    assert np.array_equal(R["t_veg_night"][p], np.where(live & ~day, S["t_veg_night"][p] + tv, S["t_veg_night"][p]))
    assert np.array_equal(R["ndaysteps"][p], S["ndaysteps"][p] + (live & day))
    assert np.array_equal(R["nnightsteps"][p], S["nnightsteps"][p] + (live & ~day))
    assert np.array_equal(R["fpsn24"][p], np.where(live, S["fpsn24"][p] + 1800.0 * R["fpsn"][p], S["fpsn24"][p]))
    lai = (S["laisun_z"][0, p] + S["laisha_z"][0, p] > 0) & (S["nrad"][p] >= 1) & live
    par = S["parsun_z"][0, p]
    assert np.array_equal(R["par24d_z"][0, p], np.where(lai, S["par24d_z"][0, p] + 1800.0 * par, S["par24d_z"][0, p]))
    assert np.array_equal(R["par24x_z"][0, p], np.where(lai & (par > S["par24x_z"][0, p]), par, S["par24x_z"][0, p]))
    # patches outside the filter keep their fill / state
    out = np.setdiff1d(np.arange(sg.npatch), p)
    assert np.all(R["tc_ref2m"][out] == 1.0e36) and np.array_equal(R["fpsn24"][out], S["fpsn24"][out])
    # NONE switches the indices off
    prm.calc_human_stress_indices = 0
    R0 = copy_state(S)
    f0 = abi.make_struct("canopyfluxes", R0, sg.bounds)
    assert oracle_lib.oracle_canopyfluxes(C.byref(prm), C.byref(sg.bounds), len(fe), abi.i32p(fe), C.byref(f0), C.byref(st)) == 0
    assert np.all(R0["wbt_ref2m"] == 1.0e36) and np.array_equal(R0["t_veg"], R["t_veg"])
