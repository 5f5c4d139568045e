"""Perturbed-parameter ensemble through the PFT tables (BASELINE config 5, PFT-parameter part): member m's patches carry
itype = m*(mxpft+1) + pft and the tables hold one copy per member (ctsm_params_t.npft_table).  The ensemble run must
agree with the oracle run of the same ensemble (1e-10, identical iteration counts), and member 0 — whose table copy is
unperturbed — must reproduce the single-parameter-set run bit for bit."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from tests.util import copy_state
from tests.test_gpu_canopy import run_oracle, run_gpu, compare

pytestmark = pytest.mark.gpu


def test_pft_table_ensemble(oracle_lib):
    L = abi.lib()
    nmem = 4
    sg, S = synthetic_canopy.make_full_case(400, seed=91)
    base = copy_state(S)
    member = synthetic_canopy.make_ensemble(sg, S, nmem, np.random.Generator(np.random.PCG64(92)))
    prm = abi.default_params()
    prm.npft_table = nmem * (abi.MXPFT + 1)
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    ref, got = copy_state(S), copy_state(S)
    rc_ref, _ = run_oracle(oracle_lib, prm, sg, ref)
    rc, st = run_gpu(L, ctx, sg, got, abi.MEM_DEVICE)
    assert rc == rc_ref == 0, st.msg
    # strongly perturbed hydraulic parameters make more patches take discontinuous paths through the ci solve (bracket /
    # brent switches) on intermediate passes: the lagging outputs are held to 1e-7 here, everything else to 1e-10
    compare(sg, got, ref, S, lag_rtol=1e-7)
    L.ctsm_b200_finalize(ctx)
    # member 0 == the run with the base tables; the perturbed members differ from it
    prm1 = abi.default_params()
    ctx1 = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm1), C.byref(ctx1)) == 0
    single = copy_state(base)
    assert run_gpu(L, ctx1, sg, single, abi.MEM_DEVICE)[0] == 0
    L.ctsm_b200_finalize(ctx1)
    fe = sg.filters["exposedvegp"] - 1
    m0 = fe[member[fe] == 0]
    mk = fe[member[fe] > 0]
    for k in ("t_veg", "qflx_tran_veg", "fpsn", "num_iter", "btran"):
        assert np.array_equal(got[k][m0], single[k][m0]), k
    assert np.mean(got["qflx_tran_veg"][mk] != single["qflx_tran_veg"][mk]) > 0.5


def test_npft_table_must_be_whole_parameter_sets():
    L = abi.lib()
    prm = abi.default_params()
    prm.npft_table = abi.MXPFT + 5
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) != 0


def test_scalar_parameter_ensemble_matches_per_member_oracle_runs(oracle_lib):
    """BASELINE config 5, scalar part: e_ice, csoilc, cv, a_coef, z_dl perturbed per member
    (ctsm_b200_set_member_params).  ONE batched CUDA call over all members must equal, member by member, the oracle run
    with that member's scalars in ctsm_params_t over that member's gridcell range (members are contiguous gridcell
    ranges, i.e. clump bounds) - CanopyFluxes (csoilc, cv, a_coef, z_dl) and SoilWater (e_ice)."""
    from oracle import oracle
    from tests.test_gpu_soil import RTOL
    from tests.util import relerr, to_device, group_arrays
    L = abi.lib()
    nmem = 3
    sg, S = synthetic_canopy.make_full_case(300, seed=95)
    rng = np.random.Generator(np.random.PCG64(96))
    member_p = synthetic_canopy.make_ensemble(sg, S, nmem, rng, spread=0.0)          # PFT tables replicated, not perturbed
    member_g = np.minimum((np.arange(sg.ngrc) * nmem) // sg.ngrc, nmem - 1)
    col_member = member_g[sg.col_gridcell - 1].astype(np.int32)
    base = abi.default_params()
    scal = {k: np.ascontiguousarray(getattr(base, k) * rng.uniform(0.6, 1.5, nmem)) for k in ("e_ice", "csoilc", "cv", "a_coef", "z_dl")}
    prm = abi.default_params()
    prm.npft_table = nmem * (abi.MXPFT + 1)
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        assert L.ctsm_b200_set_member_params(ctx, nmem, abi.f64p(scal["e_ice"]), abi.f64p(scal["csoilc"]), abi.f64p(scal["cv"]),
                                             abi.f64p(scal["a_coef"]), abi.f64p(scal["z_dl"]), abi.i32p(col_member),
                                             sg.bounds.begc, sg.bounds.endc) == 0
        got = copy_state(S)
        rc, st = run_gpu(L, ctx, sg, got, abi.MEM_DEVICE)
        assert rc == 0, st.msg
        # SoilWater through the same context
        fh = sg.filters["hydrologyc"]
        D = to_device(group_arrays(got, "soilwater"))
        dfh = to_device({"h": fh})["h"]
        fw = abi.make_struct("soilwater", D, sg.bounds)
        stw = abi.Status()
        assert L.ctsm_b200_soilwater(ctx, C.byref(sg.bounds), len(fh), abi.i32p(dfh), C.byref(fw), abi.MEM_DEVICE, C.byref(stw)) == 0
        assert L.ctsm_b200_sync(ctx, C.byref(stw)) == 0, stw.msg
        gotw = {k: v.cpu().numpy() for k, v in D.items()}
        # wrong member count is refused
        assert L.ctsm_b200_set_member_params(ctx, nmem + 1, None, None, None, None, None, None, 0, -1) == 2
    finally:
        L.ctsm_b200_finalize(ctx)
    # oracle: member by member, each with its own scalars, over its own clump bounds
    ref = copy_state(S)
    clumps, keep = oracle.make_clumps(sg, nmem)                      # equal gridcell ranges = the members
    for m in range(nmem):
        k = clumps[m]
        assert np.all(member_g[k.bounds.begg - 1:k.bounds.endg] == m)
        pm = abi.default_params()
        pm.npft_table = nmem * (abi.MXPFT + 1)
        for name, v in scal.items():
            setattr(pm, name, float(v[m]))
        st = abi.Status()
        fc = abi.make_struct("canopyfluxes", ref, sg.bounds)
        assert oracle_lib.oracle_canopyfluxes(C.byref(pm), C.byref(k.bounds), k.num_exposedvegp, k.filter_exposedvegp, C.byref(fc),
                                              C.byref(st)) == 0
    compare(sg, got, ref, S)
    refw = copy_state(ref)                                           # SoilWater continues from the post-CanopyFluxes state
    for m in range(nmem):
        k = clumps[m]
        pm = abi.default_params()
        pm.e_ice = float(scal["e_ice"][m])
        st = abi.Status()
        fwr = abi.make_struct("soilwater", refw, sg.bounds)
        assert oracle_lib.oracle_soilwater(C.byref(pm), C.byref(k.bounds), k.num_hydrologyc, k.filter_hydrologyc, C.byref(fwr),
                                           C.byref(st)) == 0
    hc = fh - 1
    assert np.array_equal(gotw["num_substeps"][hc], refw["num_substeps"][hc])
    for name in ("h2osoi_liq", "smp_l", "hk_l", "qin", "qout"):
        assert relerr(gotw[name][..., hc], refw[name][..., hc]) <= RTOL, name
    pe = abi.default_params()                                        # e_ice is felt: the base value gives other conductivities
    plainw = copy_state(ref)
    fwp = abi.make_struct("soilwater", plainw, sg.bounds)
    assert oracle_lib.oracle_soilwater(C.byref(pe), C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(fwp), C.byref(abi.Status())) == 0
    assert np.mean(plainw["hk_l"][:5, hc] != refw["hk_l"][:5, hc]) > 0.2
    # the perturbation is felt: a run with the base scalars differs
    plain = copy_state(S)
    assert run_oracle(oracle_lib, prm, sg, plain)[0] == 0
    fe = sg.filters["exposedvegp"] - 1
    assert np.mean(plain["t_veg"][fe] != ref["t_veg"][fe]) > 0.5
