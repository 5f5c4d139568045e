"""GPU parity of HydrologyNoDrainage's routines around SoilWater through the C ABI against the CPU
oracle.  Tolerance: 1e-10 relative (north_star); everything that is +, -, *, /, min, max comes out bit-identical."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, driver
from tests.util import copy_state, to_device, group_arrays
from tests.test_oracle_hydrology import case, run_infiltration

pytestmark = pytest.mark.gpu
RTOL = 1e-10
# the chain's transcendentals: exp (fsat), pow (qinmax, frac_infclust), sin (k_wet); fields downstream of none of them are exact
LIBM = {"fsat", "fcov", "qflx_sat_excess_surf", "qflx_in_soil", "qflx_top_soil_to_h2osfc", "qinmax", "qflx_infl_excess",
        "qflx_in_soil_limited", "qflx_in_h2osfc", "qflx_infl_excess_surf", "qflx_h2osfc_surf", "qflx_h2osfc_drain", "h2osfc",
        "qflx_infl", "qflx_surf"}


def compare(got, ref, S, worst):
    for fs in abi.FIELDS["infiltration"]:
        a, b = got[fs.name], ref[fs.name]
        if fs.intent == "IN":
            assert np.array_equal(a, S[fs.name], equal_nan=True), "input %s was modified" % fs.name
            continue
        if fs.name not in LIBM:
            assert np.array_equal(a, b, equal_nan=True), "%s differs" % fs.name
            continue
        fin = np.abs(b) < 1e30
        assert np.array_equal(fin, np.abs(a) < 1e30), "%s: fill pattern differs" % fs.name
        scale = float(np.max(np.abs(b[fin])))
        e = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-6 * scale + 1e-300)))
        worst[fs.name] = e
        assert e <= RTOL, (fs.name, e)


def gpu_infiltration(L, ctx, sg, S, mem, bounds=None, fl=None):
    st = abi.Status()
    fl = sg.filters if fl is None else fl
    fn, fh = fl["nolakec"], fl["hydrologyc"]
    b = C.byref(bounds if bounds is not None else sg.bounds)
    z = np.zeros(1, dtype=np.int32)
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(S, "infiltration"))
        f = abi.make_struct("infiltration", D, sg.bounds)
        dn, dh = to_device({"a": fn if len(fn) else z, "b": fh if len(fh) else z}).values()
        rc = L.ctsm_b200_hydrology_infiltration(ctx, b, len(fn), abi.i32p(dn), len(fh), abi.i32p(dh), 0, None, C.byref(f), mem, C.byref(st))
        if rc == 0:
            rc = L.ctsm_b200_sync(ctx, C.byref(st))
        for k, v in D.items():
            S[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct("infiltration", S, sg.bounds)
        rc = L.ctsm_b200_hydrology_infiltration(ctx, b, len(fn), abi.i32p(fn if len(fn) else z), len(fh), abi.i32p(fh if len(fh) else z),
                                                0, None, C.byref(f), mem, C.byref(st))
    return rc, st


@pytest.mark.parametrize("mem", [abi.MEM_DEVICE, abi.MEM_HOST])
@pytest.mark.parametrize("h2osfcflag,crop0", [(1, 0), (0, 1)], ids=["default", "noh2osfc_cropfsat0"])
def test_infiltration_matches_oracle(oracle_lib, mem, h2osfcflag, crop0):
    L = abi.lib()
    sg, S = case(6000, 701)
    prm = abi.default_params()
    prm.h2osfcflag, prm.crop_fsat_equals_zero = h2osfcflag, crop0
    ref, got = copy_state(S), copy_state(S)
    assert run_infiltration(oracle_lib, prm, sg, ref) == 0
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        rc, st = gpu_infiltration(L, ctx, sg, got, mem)
        assert rc == 0, st.msg
    finally:
        L.ctsm_b200_finalize(ctx)
    worst = {}
    compare(got, ref, S, worst)
    ck = sg.filters["hydrologyc"][5] - 1                         # SurfaceWaterMod.F90:499's REAL(4) 1.0e-8: this column keeps its runoff
    if h2osfcflag:
        assert 9.99999994e-9 < got["qflx_h2osfc_surf"][ck] < 1.0e-8
    print("infiltration chain worst:", sorted(worst.items(), key=lambda kv: -kv[1])[:5])


def test_infiltration_empty_filters_clump_bounds_and_urban(oracle_lib):
    L = abi.lib()
    sg, S = case(1500, 711)
    prm = abi.default_params()
    ref = copy_state(S)
    assert run_infiltration(oracle_lib, prm, sg, ref) == 0
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        got = copy_state(S)
        z = np.zeros(0, dtype=np.int32)
        rc, st = gpu_infiltration(L, ctx, sg, got, abi.MEM_HOST, fl={"nolakec": z, "hydrologyc": z})
        assert rc == 0
        for k in S:
            assert np.array_equal(got[k], S[k], equal_nan=True), k
        for kb, fl in driver.make_slabs(sg, 5):                      # clump by clump (bounds != alloc), host arrays
            rc, st = gpu_infiltration(L, ctx, sg, got, abi.MEM_HOST, bounds=kb, fl=fl)
            assert rc == 0, st.msg
        compare(got, ref, S, {})
        st = abi.Status()
        f = abi.make_struct("infiltration", got, sg.bounds)
        fn, fh = sg.filters["nolakec"], sg.filters["hydrologyc"]
        assert L.ctsm_b200_hydrology_infiltration(ctx, C.byref(sg.bounds), len(fn), abi.i32p(fn), len(fh), abi.i32p(fh), 1, abi.i32p(fn),
                                                  C.byref(f), abi.MEM_HOST, C.byref(st)) == 16
        got["lun_itype"][fh[7] - 1] = 8
        rc = L.ctsm_b200_hydrology_infiltration(ctx, C.byref(sg.bounds), len(fn), abi.i32p(fn), len(fh), abi.i32p(fh), 0, None,
                                                C.byref(f), abi.MEM_HOST, C.byref(st))
        assert rc == 16 and st.subgrid_index == fh[7]
    finally:
        L.ctsm_b200_finalize(ctx)


def test_step_with_infiltration_feeds_soilwater(oracle_lib):
    """patch2col -> infiltration chain -> root-water sink -> SoilWater device-resident through driver.HotPath: SoilWater consumes the
    icefrac / eff_porosity / qflx_infl the chain produced on the device; against the same sequence of the oracle."""
    import torch
    from tests.util import compare_step_fields
    sg, S = case(3000, 721)
    prm = abi.default_params()
    ref = copy_state(S)
    assert run_infiltration(oracle_lib, prm, sg, ref) == 0
    st = abi.Status()
    fh = sg.filters["hydrologyc"]
    fs = abi.make_struct("plantsink", ref, sg.bounds)
    assert oracle_lib.oracle_vert_tran_sink_hydstress(C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(fs)) == 0
    fw = abi.make_struct("soilwater", ref, sg.bounds)
    assert oracle_lib.oracle_soilwater(C.byref(prm), C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(fw), C.byref(st)) == 0
    ctx = driver.Context(prm)
    try:
        routines = ("infiltration", "plantsink", "soilwater")
        names = sorted({f.name for g in routines for f in abi.FIELDS[g]})
        D = {k: torch.from_numpy(np.ascontiguousarray(S[k])).cuda() for k in names}
        driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, routines).step()
        ctx.sync()
        got = {k: (D[k].cpu().numpy() if k in D else S[k]) for k in S}
    finally:
        ctx.close()
    worst = compare_step_fields(sg, S, got, ref, np.zeros(sg.npatch, dtype=bool), routines)
    assert np.array_equal(got["num_substeps"], ref["num_substeps"])
    print("infiltration -> sink -> SoilWater worst:", sorted(worst.items(), key=lambda kv: -kv[1])[:4])


# ----------------------------------------------------------------------------------------------------------------------
# PerchedWaterTable / ThetaBasedWaterTable / RenewCondensation, the closing diagnostics, and HydrologyNoDrainage as a whole
from tests.test_oracle_hydrology import wt_case, run_water_table, run_diagnostics   # noqa: E402


def _gpu(L, ctx, group, sg, S, mem, filters, call):
    st = abi.Status()
    z = np.zeros(1, dtype=np.int32)
    filters = [f if len(f) else z for f in filters]
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(S, group))
        f = abi.make_struct(group, D, sg.bounds)
        rc = call(f, [to_device({"f": v})["f"] for v in filters], st)
        if rc == 0:
            rc = L.ctsm_b200_sync(ctx, C.byref(st))
        for k, v in D.items():
            S[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct(group, S, sg.bounds)
        rc = call(f, filters, st)
    return rc, st


def gpu_water_table(L, ctx, sg, S, mem, fh, bounds=None):
    b = C.byref(bounds if bounds is not None else sg.bounds)
    return _gpu(L, ctx, "watertable", sg, S, mem, [fh], lambda f, fl, st: L.ctsm_b200_water_table(
        ctx, b, len(fh), abi.i32p(fl[0]), 0, None, C.byref(f), mem, C.byref(st)))


def gpu_diagnostics(L, ctx, sg, S, mem, fn, fs, fns, fh, bounds=None):
    b = C.byref(bounds if bounds is not None else sg.bounds)
    return _gpu(L, ctx, "hydrodiag", sg, S, mem, [fn, fs, fns, fh], lambda f, fl, st: L.ctsm_b200_hydrology_diagnostics(
        ctx, b, len(fn), abi.i32p(fl[0]), len(fs), abi.i32p(fl[1]), len(fns), abi.i32p(fl[2]), len(fh), abi.i32p(fl[3]), 0, None,
        C.byref(f), mem, C.byref(st)))


@pytest.mark.parametrize("mem", [abi.MEM_DEVICE, abi.MEM_HOST])
def test_water_table_bit_exact(oracle_lib, mem):
    """no transcendentals in the three routines: every field identical to the oracle's, bit for bit"""
    L = abi.lib()
    sg, S = wt_case(6000, 731)
    prm = abi.default_params()
    ref, got = copy_state(S), copy_state(S)
    rc, st = run_water_table(oracle_lib, prm, sg, ref)
    assert rc == 0
    fh = sg.filters["hydrologyc"]
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        if mem == abi.MEM_HOST:                                  # clump by clump, bounds != alloc
            for kb, fl in driver.make_slabs(sg, 3):
                rc, st = gpu_water_table(L, ctx, sg, got, mem, fl["hydrologyc"], bounds=kb)
                assert rc == 0, st.msg
        else:
            rc, st = gpu_water_table(L, ctx, sg, got, mem, fh)
            assert rc == 0, st.msg
        for f in abi.FIELDS["watertable"]:
            assert np.array_equal(got[f.name], ref[f.name], equal_nan=True), f.name
        bad = copy_state(S)
        bare = fh[S["snl"][fh - 1] == 0]
        bad["qflx_solidevap_from_top_layer"][bare[5] - 1] = 10.0
        rc, st = gpu_water_table(L, ctx, sg, bad, mem, fh)
        assert rc == 18 and st.subgrid_index == bare[5] and b"RenewCondensation" in st.msg
        rc, st = gpu_water_table(L, ctx, sg, copy_state(S), mem, fh[:0])
        assert rc == 0
    finally:
        L.ctsm_b200_finalize(ctx)
    assert (ref["zwt_perched"][fh - 1] != ref["frost_table"][fh - 1]).sum() > 100


@pytest.mark.parametrize("mem", [abi.MEM_DEVICE, abi.MEM_HOST])
def test_hydrology_diagnostics_match_oracle(oracle_lib, mem):
    from tests.test_oracle_snow import snow_filters
    L = abi.lib()
    sg, S = wt_case(6000, 741)
    S["dz"][12, ::7] = 0.1
    prm = abi.default_params()
    fs, fns = snow_filters(oracle_lib, sg, S)
    fn, fh = sg.filters["nolakec"], sg.filters["hydrologyc"]
    ref, got = copy_state(S), copy_state(S)
    rc, st = run_diagnostics(oracle_lib, prm, sg, ref, fs, fns)
    assert rc == 0
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        if mem == abi.MEM_HOST:
            for kb, fl in driver.make_slabs(sg, 3):
                cut = lambda a: a[(a >= kb.begc) & (a <= kb.endc)]
                rc, st = gpu_diagnostics(L, ctx, sg, got, mem, fl["nolakec"], cut(fs), cut(fns), fl["hydrologyc"], bounds=kb)
                assert rc == 0, st.msg
        else:
            rc, st = gpu_diagnostics(L, ctx, sg, got, mem, fn, fs, fns, fh)
            assert rc == 0, st.msg
    finally:
        L.ctsm_b200_finalize(ctx)
    worst = {}
    for f in abi.FIELDS["hydrodiag"]:
        a, b = got[f.name], ref[f.name]
        if f.intent == "IN":
            assert np.array_equal(a, S[f.name], equal_nan=True), f.name
        elif f.name in ("soilpsi", "smp_l", "wf", "wf2"):        # pow
            fin = np.abs(b) < 1e30
            assert np.array_equal(fin, np.abs(a) < 1e30), f.name
            worst[f.name] = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300)))
            assert worst[f.name] <= RTOL, (f.name, worst[f.name])
        else:
            assert np.array_equal(a, b, equal_nan=True), f.name
    print("hydrology diagnostics worst:", worst)


def test_hydrology_no_drainage_device_resident(oracle_lib):
    """HydrologyNoDrainage as far as it is built, in the reference's order (HydrologyNoDrainageMod.F90:279-757): BuildSnowFilter,
    SnowWater, infiltration chain, root-water sink, SoilWater, water tables + RenewCondensation, snow-layer update, BuildSnowFilter,
    diagnostics - device-resident through driver.HotPath against the same sequence of the oracle."""
    import torch
    from tests.test_oracle_snow import snow_filters, run_snow_water, run_snow_layers, run_snow_capping
    sg, S = wt_case(3000, 751, saturate=False)
    rng = np.random.Generator(np.random.PCG64(752))
    S["topo"] = rng.uniform(0.0, 3000.0, sg.ncol)
    heavy = np.nonzero(S["snl"] < 0)[0][::9]                    # a few packs above h2osno_max
    S["h2osoi_ice"][11, heavy] = 11000.0
    S["dz"][11, heavy] = 11000.0 / 500.0
    for k in ("qflx_snwcp_ice", "qflx_snwcp_liq", "qflx_snwcp_discarded_ice", "qflx_snwcp_discarded_liq"):
        S[k] = np.full(sg.ncol, 1.0e36)
    prm = abi.default_params()
    fh = sg.filters["hydrologyc"]
    ref = copy_state(S)
    fs, fns = snow_filters(oracle_lib, sg, ref)
    st = abi.Status()
    assert run_snow_water(oracle_lib, prm, sg, ref, fs, fns)[0] == 0
    assert run_infiltration(oracle_lib, prm, sg, ref) == 0
    fsk = abi.make_struct("plantsink", ref, sg.bounds)
    assert oracle_lib.oracle_vert_tran_sink_hydstress(C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(fsk)) == 0
    fw = abi.make_struct("soilwater", ref, sg.bounds)
    assert oracle_lib.oracle_soilwater(C.byref(prm), C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(fw), C.byref(st)) == 0
    assert run_water_table(oracle_lib, prm, sg, ref)[0] == 0
    assert run_snow_capping(oracle_lib, prm, sg, ref, sg.filters["nolakec"], fs, 1000)[0] == 0
    assert run_snow_layers(oracle_lib, prm, sg, ref, fs)[0] == 0
    fs2, fns2 = snow_filters(oracle_lib, sg, ref)
    assert run_diagnostics(oracle_lib, prm, sg, ref, fs2, fns2)[0] == 0
    routines = ("snowwater", "infiltration", "plantsink", "soilwater", "watertable", "snowcapping", "snowlayers", "hydrodiag")
    ctx = driver.Context(prm)
    try:
        names = sorted({f.name for g in routines for f in abi.FIELDS[g]})
        D = {k: torch.from_numpy(np.ascontiguousarray(S[k])).cuda() for k in names}
        driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, routines).step()
        ctx.sync()
        got = {k: (D[k].cpu().numpy() if k in D else S[k]) for k in S}
    finally:
        ctx.close()
    assert np.array_equal(got["snl"], ref["snl"]) and np.array_equal(got["num_substeps"], ref["num_substeps"])
    worst = {}
    for g in routines:
        for f in abi.FIELDS[g]:
            if f.intent == "IN" or f.ctype == "int":
                continue
            a, b = got[f.name], ref[f.name]
            fin = np.abs(b) < 1e30
            assert np.array_equal(fin, np.abs(a) < 1e30), f.name
            if not fin.any():
                continue
            scale = float(np.max(np.abs(b[fin])))
            e = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-6 * scale + 1e-300)))
            worst[f.name] = max(worst.get(f.name, 0.0), e)
    bad = {k: v for k, v in worst.items() if not v <= RTOL}
    print("HydrologyNoDrainage worst:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    assert not bad, bad


def test_hydrology_no_drainage_two_steps_device_resident(oracle_lib):
    """Two consecutive HydrologyNoDrainage steps without the host in between: the second step consumes the layer structure (snl, dz, zi),
    water tables and surface store the first one left on the device.  Against the oracle run twice.
    Two SoilWater steps under an unchanged synthetic forcing are ill-conditioned on a few columns (the oracle's own result moves by up
    to 1e-7 when its inputs are nudged by one ulp), so conditioning is MEASURED as in the canopy tests: a column whose oracle result
    moves by more than 1e-12 under the nudge is held to 1e3 x its measured sensitivity, every other column to 1e-10."""
    import torch
    from tests.test_oracle_snow import snow_filters, run_snow_water, run_snow_layers, run_snow_capping
    sg, S = wt_case(2500, 761, saturate=False)
    # the forcing stays the same for both steps (a column whose pack vanishes in step 1 has no ice left to sublimate in step 2)
    S["qflx_solidevap_from_top_layer"][:] = 0.0
    S["qflx_liqevap_from_top_layer"] = np.minimum(S["qflx_liqevap_from_top_layer"], 1.0e-6)
    S["topo"] = np.random.Generator(np.random.PCG64(762)).uniform(0.0, 3000.0, sg.ncol)
    for k in ("qflx_snwcp_ice", "qflx_snwcp_liq", "qflx_snwcp_discarded_ice", "qflx_snwcp_discarded_liq"):
        S[k] = np.full(sg.ncol, 1.0e36)
    prm = abi.default_params()
    fh, fn = sg.filters["hydrologyc"], sg.filters["nolakec"]

    def oracle_two_steps(S0):
        ref = copy_state(S0)
        st = abi.Status()
        hist = []
        for step in range(2):
            fs, fns = snow_filters(oracle_lib, sg, ref)
            assert run_snow_water(oracle_lib, prm, sg, ref, fs, fns)[0] == 0
            assert run_infiltration(oracle_lib, prm, sg, ref) == 0
            fsk = abi.make_struct("plantsink", ref, sg.bounds)
            assert oracle_lib.oracle_vert_tran_sink_hydstress(C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(fsk)) == 0
            fw = abi.make_struct("soilwater", ref, sg.bounds)
            assert oracle_lib.oracle_soilwater(C.byref(prm), C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(fw), C.byref(st)) == 0
            assert run_water_table(oracle_lib, prm, sg, ref)[0] == 0
            assert run_snow_capping(oracle_lib, prm, sg, ref, fn, fs, 1000)[0] == 0
            assert run_snow_layers(oracle_lib, prm, sg, ref, fs)[0] == 0
            fs2, fns2 = snow_filters(oracle_lib, sg, ref)
            assert run_diagnostics(oracle_lib, prm, sg, ref, fs2, fns2)[0] == 0
            hist.append(ref["snl"].copy())
        return ref, hist

    ref, snl_hist = oracle_two_steps(S)
    assert (snl_hist[0] != S["snl"]).sum() > 300 and (snl_hist[1] != snl_hist[0]).sum() > 20      # the layer structure keeps moving
    nudged = copy_state(S)
    rng = np.random.Generator(np.random.PCG64(763))
    for k in ("h2osoi_liq", "dz", "hksat"):
        nudged[k] = nudged[k] * (1.0 + rng.choice([-1.0, 1.0], nudged[k].shape) * 2.2e-16)
    twin, _ = oracle_two_steps(nudged)
    routines = ("snowwater", "infiltration", "plantsink", "soilwater", "watertable", "snowcapping", "snowlayers", "hydrodiag")
    ctx = driver.Context(prm)
    try:
        names = sorted({f.name for g in routines for f in abi.FIELDS[g]})
        D = {k: torch.from_numpy(np.ascontiguousarray(S[k])).cuda() for k in names}
        hp = driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, routines)
        hp.step()
        hp.step()
        ctx.sync()
        got = {k: (D[k].cpu().numpy() if k in D else S[k]) for k in S}
    finally:
        ctx.close()
    assert np.array_equal(got["snl"], ref["snl"]) and np.array_equal(got["num_substeps"], ref["num_substeps"])

    def colerr(a, b):                       # worst relative error per column of a COL field (any level shape)
        fin = np.abs(b) < 1e30
        scale = float(np.max(np.abs(b[fin]))) if fin.any() else 1.0
        e = np.where(fin, np.abs(np.where(fin, a, 0.0) - np.where(fin, b, 0.0)) / np.maximum(np.abs(np.where(fin, b, 0.0)), 1e-6 * scale + 1e-300), 0.0)
        return e.reshape(-1, e.shape[-1]).max(axis=0)

    sens = np.zeros(sg.ncol)
    fields = [f for g in routines for f in abi.FIELDS[g] if f.intent != "IN" and f.ctype != "int" and f.sub == "COL"]
    for f in fields:
        sens = np.maximum(sens, colerr(twin[f.name], ref[f.name]))
    ill = sens > 1e-12
    worst_ok, worst_ill = {}, 0.0
    for f in fields:
        a, b = got[f.name], ref[f.name]
        assert np.array_equal(np.abs(b) < 1e30, np.abs(a) < 1e30), f.name
        e = colerr(a, b)
        worst_ok[f.name] = float(e[~ill].max())
        assert worst_ok[f.name] <= RTOL, (f.name, worst_ok[f.name])
        if ill.any():
            worst_ill = max(worst_ill, float((e[ill] / np.maximum(1e3 * sens[ill], RTOL)).max()))
    assert worst_ill <= 1.0, worst_ill
    assert ill.sum() < 0.05 * sg.ncol
    print("two HydrologyNoDrainage steps: %d ill-conditioned columns of %d (oracle sensitivity up to %.2g), worst elsewhere:" %
          (int(ill.sum()), sg.ncol, float(sens.max())), sorted(worst_ok.items(), key=lambda kv: -kv[1])[:4])


@pytest.mark.parametrize("nslab", [1, 3])
def test_hydrology_no_drainage_in_a_resident_window(nslab):
    """HydrologyNoDrainage with host-owned arrays inside a resident window (asynchronous staging, clump by clump): the snow filters
    are built from the DEVICE copy of col%snl where that is the current one (after the snow-layer update, whose download is still in
    flight), and the host arrays end up bit-identical to the self-contained CTSM_MEM_HOST calls."""
    sg, S = wt_case(2000, 771, saturate=False)
    S["topo"] = np.random.Generator(np.random.PCG64(772)).uniform(0.0, 3000.0, sg.ncol)
    for k in ("qflx_snwcp_ice", "qflx_snwcp_liq", "qflx_snwcp_discarded_ice", "qflx_snwcp_discarded_liq"):
        S[k] = np.full(sg.ncol, 1.0e36)
    routines = ("snowwater", "infiltration", "plantsink", "soilwater", "watertable", "snowcapping", "snowlayers", "hydrodiag")
    ctx = driver.Context(abi.default_params())
    try:
        plain, win = copy_state(S), copy_state(S)
        driver.HotPath(ctx, sg, plain, abi.MEM_HOST, routines).step()
        hp = driver.HotPath(ctx, sg, win, abi.MEM_HOST, routines, nslab=nslab, window=True)
        hp.step()
        names = sorted({f.name for g in routines for f in abi.FIELDS[g]})
        for k in names:
            assert np.array_equal(plain[k], win[k], equal_nan=True), k
        assert (plain["snl"] != S["snl"]).sum() > 200
    finally:
        ctx.close()
