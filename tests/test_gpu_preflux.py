"""GPU parity of the routines clm_drv runs just before CanopyFluxes - BiogeophysPreFluxCalcs,
CalculateSurfaceHumidity, BareGroundFluxes - through the C ABI against the CPU oracle, chained as in clm_drv
(clm_driver.F90:680-718).  Tolerance: 1e-10 relative (north_star); fields without transcendentals come out bit-identical."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi
from tests.util import copy_state, to_device, group_arrays
from tests.test_oracle_preflux import case, run_preflux, run_humidity, run_bare

pytestmark = pytest.mark.gpu
RTOL = 1e-10
# results of +, -, *, /, min, max and copies only: identical bits expected
# (compared after the whole chain: BareGroundFluxes overwrites some of BiogeophysPreFluxCalcs' patch outputs)
EXACT = {"preflux": ("t_ssbef", "t_h2osfc_bef", "t_grnd", "emg", "htvp", "beta", "zii", "thv"),
         "surfacehumidity": (), "baregroundfluxes": ("btran", "t_veg", "rootr", "rresis", "displa", "z0mv", "num_iter", "qflx_tran_veg")}


CHAIN_OUTPUTS = {fs.name for g in ("preflux", "surfacehumidity", "baregroundfluxes") for fs in abi.FIELDS[g] if fs.intent != "IN"}


def gpu_call(L, ctx, group, sg, S, mem, call):
    st = abi.Status()
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(S, group))
        f = abi.make_struct(group, D, sg.bounds)
        flt = {k: to_device({"f": v})["f"] for k, v in sg.filters.items() if len(v)}
        rc = call(f, flt, st)
        assert rc == 0, st.msg
        rc = L.ctsm_b200_sync(ctx, C.byref(st))
        for k, v in D.items():
            S[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct(group, S, sg.bounds)
        rc = call(f, sg.filters, st)
    return rc, st


def compare(group, got, ref, worst):
    for fs in abi.FIELDS[group]:
        a, b = got[fs.name], ref[fs.name]
        if fs.intent == "IN":
            if fs.name not in CHAIN_OUTPUTS:               # (an input produced earlier in the chain carries that routine's error)
                assert np.array_equal(a, b, equal_nan=True), "input %s was modified" % fs.name
            continue
        if fs.ctype == "int" or fs.name in EXACT[group]:
            assert np.array_equal(a, b, equal_nan=True), "%s differs" % fs.name
            continue
        fin = np.abs(b) < 1e30
        assert np.array_equal(fin, np.abs(a) < 1e30), "%s: fill pattern differs" % fs.name
        if not fin.any():
            continue
        scale = float(np.max(np.abs(b[fin])))
        e = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-6 * scale + 1e-300)))
        worst[group + "." + fs.name] = e
        assert e <= RTOL, (group, fs.name, e)


@pytest.mark.parametrize("mem", [abi.MEM_DEVICE, abi.MEM_HOST])
@pytest.mark.parametrize("method,resis", [(2, 1), (1, 0)], ids=["meier2022_sl14", "zengwang2007_leepielke"])
def test_preflux_chain_matches_oracle(oracle_lib, mem, method, resis):
    L = abi.lib()
    sg, S = case(6000, 501, wet_every=3)
    S["htop"][sg.filters["nolakep"][::17] - 1] = 0.0
    prm = abi.default_params()
    prm.z0param_method, prm.soil_resis_method = method, resis
    ref, got = copy_state(S), copy_state(S)
    assert run_preflux(oracle_lib, prm, sg, ref) == 0
    assert run_humidity(oracle_lib, sg, ref) == 0
    assert run_bare(oracle_lib, prm, sg, ref) == 0
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    worst = {}
    try:
        b = C.byref(sg.bounds)
        n = {k: len(v) for k, v in sg.filters.items()}
        rc, st = gpu_call(L, ctx, "preflux", sg, got, mem, lambda f, fl, st: L.ctsm_b200_biogeophys_pre_flux_calcs(
            ctx, b, n["nolakec"], abi.i32p(fl["nolakec"]), n["nolakep"], abi.i32p(fl["nolakep"]), 0, None, 0, C.byref(f), mem, C.byref(st)))
        assert rc == 0, st.msg
        rc, st = gpu_call(L, ctx, "surfacehumidity", sg, got, mem, lambda f, fl, st: L.ctsm_b200_calculate_surface_humidity(
            ctx, b, n["nolakec"], abi.i32p(fl["nolakec"]), C.byref(f), mem, C.byref(st)))
        assert rc == 0, st.msg
        rc, st = gpu_call(L, ctx, "baregroundfluxes", sg, got, mem, lambda f, fl, st: L.ctsm_b200_bare_ground_fluxes(
            ctx, b, n["noexposedvegp"], abi.i32p(fl["noexposedvegp"]), C.byref(f), mem, C.byref(st)))
        assert rc == 0, st.msg
    finally:
        L.ctsm_b200_finalize(ctx)
    for g in ("preflux", "surfacehumidity", "baregroundfluxes"):
        compare(g, got, ref, worst)
    print("f2 chain worst:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    assert len(sg.filters["noexposedvegp"]) > 5000


def test_preflux_refuses_urban_and_first_steps(oracle_lib):
    L = abi.lib()
    sg, S = case(300, 511)
    prm = abi.default_params()
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        b = C.byref(sg.bounds)
        fc, fp = sg.filters["nolakec"], sg.filters["nolakep"]
        ref, got = copy_state(S), copy_state(S)
        assert run_preflux(oracle_lib, prm, sg, ref, flags=1) == 0
        st = abi.Status()
        f = abi.make_struct("preflux", got, sg.bounds)
        assert L.ctsm_b200_biogeophys_pre_flux_calcs(ctx, b, len(fc), abi.i32p(fc), len(fp), abi.i32p(fp), 0, None, 1, C.byref(f),
                                                     abi.MEM_HOST, C.byref(st)) == 0
        np.testing.assert_array_equal(got["z0m"], ref["z0m"])
        np.testing.assert_array_equal(got["forc_hgt_u_patch"], ref["forc_hgt_u_patch"])
        # an urban column filter, or a column of an urban landunit in filter_nolakec, is refused
        assert L.ctsm_b200_biogeophys_pre_flux_calcs(ctx, b, len(fc), abi.i32p(fc), len(fp), abi.i32p(fp), 1, abi.i32p(fc), 0,
                                                     C.byref(f), abi.MEM_HOST, C.byref(st)) == 16
        got["lun_itype"][fc[5] - 1] = 8
        rc = L.ctsm_b200_biogeophys_pre_flux_calcs(ctx, b, len(fc), abi.i32p(fc), len(fp), abi.i32p(fp), 0, None, 0, C.byref(f),
                                                   abi.MEM_HOST, C.byref(st))
        assert rc == 16 and st.subgrid_index == fc[5]
    finally:
        L.ctsm_b200_finalize(ctx)


def test_step_from_preflux_to_balancecheck(oracle_lib):
    """The ten-routine step in clm_drv order (clm_driver.F90:680-1422): BiogeophysPreFluxCalcs -> CalculateSurfaceHumidity ->
    BareGroundFluxes -> CanopyFluxes -> SoilTemperature -> SoilFluxes -> patch2col -> root-water sink -> SoilWater ->
    BalanceCheck, device-resident through driver.HotPath, against the same chain of the oracle."""
    from oracle import oracle
    import torch
    from ctsm_b200 import driver
    from tests.util import compare_step_fields
    from tests.test_gpu_canopy import compare as compare_canopy
    sg, S = case(4000, 521, wet_every=3)
    prm = abi.default_params()
    ref = copy_state(S)
    assert run_preflux(oracle_lib, prm, sg, ref) == 0
    assert run_humidity(oracle_lib, sg, ref) == 0
    assert run_bare(oracle_lib, prm, sg, ref) == 0
    before_canopy = copy_state(ref)
    oprm = abi.default_params()
    oprm.balance_skip_steps = int(oracle_lib.oracle_balancecheck_skip_steps(oprm.dtime))
    clumps, keep = oracle.make_clumps(sg, 8)
    order = ("soiltemperature", "soilwater", "canopyfluxes", "plantsink", "balancecheck", "soilfluxes", "patch2col")
    ref_c = copy_state(ref)
    fcs = abi.make_struct("canopyfluxes", ref_c, sg.bounds)
    assert oracle_lib.oracle_step_clumps(C.byref(oprm), len(clumps), clumps, None, None, C.byref(fcs), 4) == 0
    structs = [abi.make_struct(g, ref, sg.bounds) for g in order]
    assert oracle_lib.oracle_fullstep_clumps(C.byref(oprm), len(clumps), clumps, *[C.byref(x) for x in structs], 1, 127) == 0
    ctx = driver.Context(prm)
    try:
        routines = driver.PRE_ROUTINES + driver.ROUTINES
        names = sorted({fs.name for g in routines for fs in abi.FIELDS[g]})
        D = {k: torch.from_numpy(np.ascontiguousarray(S[k])).cuda() for k in names}
        driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, driver.PRE_ROUTINES + ("canopyfluxes",)).step()
        ctx.sync()
        got_c = {k: (D[k].cpu().numpy() if k in D else S[k]) for k in S}
        driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, driver.ROUTINES[1:]).step()
        ctx.sync()
        got = {k: (D[k].cpu().numpy() if k in D else S[k]) for k in S}
    finally:
        ctx.close()
    worst_c, ntie = compare_canopy(sg, got_c, ref_c, before_canopy, check_inputs=False)
    fe = sg.filters["exposedvegp"] - 1
    loose_p = np.zeros(sg.npatch, dtype=bool)
    loose_p[fe[(got_c["num_iter"][fe] != ref_c["num_iter"][fe]) | (ref_c["num_iter"][fe] >= 41)]] = True
    worst = compare_step_fields(sg, S, got, ref, loose_p, driver.ROUTINES[1:])
    print("ten-routine step: canopy worst", sorted(worst_c.items(), key=lambda kv: -kv[1])[:3], "ties", ntie,
          "rest worst", sorted(worst.items(), key=lambda kv: -kv[1])[:4])


def test_preflux_routines_with_empty_filters_and_clump_bounds(oracle_lib):
    """Empty filters are legal (a clump without non-lake columns) and write nothing; a call with one clump's bounds on
    proc-sized arrays touches that clump only and gives the undecomposed result there."""
    L = abi.lib()
    sg, S = case(1200, 531, wet_every=3)
    prm = abi.default_params()
    ref = copy_state(S)
    assert run_preflux(oracle_lib, prm, sg, ref) == 0
    assert run_humidity(oracle_lib, sg, ref) == 0
    assert run_bare(oracle_lib, prm, sg, ref) == 0
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        got = copy_state(S)
        st = abi.Status()
        b = C.byref(sg.bounds)
        z = np.zeros(1, dtype=np.int32)
        fpre = abi.make_struct("preflux", got, sg.bounds)
        fhum = abi.make_struct("surfacehumidity", got, sg.bounds)
        fbar = abi.make_struct("baregroundfluxes", got, sg.bounds)
        assert L.ctsm_b200_biogeophys_pre_flux_calcs(ctx, b, 0, abi.i32p(z), 0, abi.i32p(z), 0, None, 0, C.byref(fpre), abi.MEM_HOST, C.byref(st)) == 0
        assert L.ctsm_b200_calculate_surface_humidity(ctx, b, 0, abi.i32p(z), C.byref(fhum), abi.MEM_HOST, C.byref(st)) == 0
        assert L.ctsm_b200_bare_ground_fluxes(ctx, b, 0, abi.i32p(z), C.byref(fbar), abi.MEM_HOST, C.byref(st)) == 0
        for k in S:
            assert np.array_equal(got[k], S[k], equal_nan=True), k
        # clump by clump (bounds != alloc), host arrays
        from ctsm_b200 import driver
        for kb, fl in driver.make_slabs(sg, 5):
            cb = C.byref(kb)
            fc, fp, fb = fl["nolakec"], fl["nolakep"], fl["noexposedvegp"]
            assert L.ctsm_b200_biogeophys_pre_flux_calcs(ctx, cb, len(fc), abi.i32p(fc), len(fp), abi.i32p(fp), 0, None, 0,
                                                         C.byref(fpre), abi.MEM_HOST, C.byref(st)) == 0
            assert L.ctsm_b200_calculate_surface_humidity(ctx, cb, len(fc), abi.i32p(fc), C.byref(fhum), abi.MEM_HOST, C.byref(st)) == 0
            assert L.ctsm_b200_bare_ground_fluxes(ctx, cb, len(fb), abi.i32p(fb), C.byref(fbar), abi.MEM_HOST, C.byref(st)) == 0
        worst = {}
        for g in ("preflux", "surfacehumidity", "baregroundfluxes"):
            compare(g, got, ref, worst)
    finally:
        L.ctsm_b200_finalize(ctx)
