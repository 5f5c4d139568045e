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
