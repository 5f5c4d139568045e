"""GPU parity of the snow routines of HydrologyNoDrainage through the C ABI against the CPU oracle:
BuildSnowFilter (bit-exact), SnowWater (no transcendentals: bit-exact), SnowCompaction + CombineSnowLayers + DivideSnowLayers +
ZeroEmptySnowLayers (identical layer counts, reals within 1e-10: compaction's exp / pow / acos differ by an ulp between libm and
CUDA, and the subdivision proportions inherit it)."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, driver
from tests.util import copy_state, to_device, group_arrays
from tests.test_oracle_snow import case, snow_filters, run_snow_water, run_snow_layers

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def gpu_call(L, ctx, group, sg, S, mem, filters, call):
    """filters: list of int32 arrays; call(f, device-or-host filter pointers..., st) -> rc"""
    st = abi.Status()
    z = np.zeros(1, dtype=np.int32)
    filters = [f if len(f) else z for f in filters]
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(S, group))
        f = abi.make_struct(group, D, sg.bounds)
        dfl = [to_device({"f": v})["f"] for v in filters]
        rc = call(f, dfl, st)
        if rc == 0:
            rc = L.ctsm_b200_sync(ctx, C.byref(st))
        for k, v in D.items():
            S[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct(group, S, sg.bounds)
        rc = call(f, filters, st)
    return rc, st


def gpu_snow_water(L, ctx, sg, S, mem, fs, fns, bounds=None):
    b = C.byref(bounds if bounds is not None else sg.bounds)
    return gpu_call(L, ctx, "snowwater", sg, S, mem, [fs, fns], lambda f, fl, st: L.ctsm_b200_snow_water(
        ctx, b, len(fs), abi.i32p(fl[0]), len(fns), abi.i32p(fl[1]), C.byref(f), mem, C.byref(st)))


def gpu_snow_layers(L, ctx, sg, S, mem, fs, bounds=None):
    b = C.byref(bounds if bounds is not None else sg.bounds)
    return gpu_call(L, ctx, "snowlayers", sg, S, mem, [fs], lambda f, fl, st: L.ctsm_b200_snow_layers(
        ctx, b, len(fs), abi.i32p(fl[0]), C.byref(f), mem, C.byref(st)))


@pytest.mark.parametrize("mem", [abi.MEM_DEVICE, abi.MEM_HOST])
def test_build_snow_filter_bit_exact(oracle_lib, mem):
    L = abi.lib()
    sg, S = case(5000, 901)
    fs, fns = snow_filters(oracle_lib, sg, S)
    fn = sg.filters["nolakec"]
    prm = abi.default_params()
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        for flt in (fn, fn[:1], fn[:0], fn[:2049]):
            a, b = np.zeros(max(len(flt), 1), np.int32), np.zeros(max(len(flt), 1), np.int32)
            na, nb = C.c_int32(-1), C.c_int32(-1)
            if mem == abi.MEM_DEVICE:
                D = to_device({"f": flt if len(flt) else np.zeros(1, np.int32), "snl": S["snl"], "a": a, "b": b})
                rc = L.ctsm_b200_build_snow_filter(ctx, C.byref(sg.bounds), len(flt), abi.i32p(D["f"]), abi.i32p(D["snl"]), sg.bounds.begc,
                                                   sg.bounds.endc, abi.i32p(D["a"]), C.byref(na), abi.i32p(D["b"]), C.byref(nb), mem)
                a, b = D["a"].cpu().numpy(), D["b"].cpu().numpy()
            else:
                rc = L.ctsm_b200_build_snow_filter(ctx, C.byref(sg.bounds), len(flt), abi.i32p(flt if len(flt) else a), abi.i32p(S["snl"]),
                                                   sg.bounds.begc, sg.bounds.endc, abi.i32p(a), C.byref(na), abi.i32p(b), C.byref(nb), mem)
            assert rc == 0
            snow = S["snl"][flt - 1] < 0
            assert na.value == snow.sum() and nb.value == (~snow).sum()
            assert np.array_equal(a[:na.value], flt[snow]) and np.array_equal(b[:nb.value], flt[~snow])
    finally:
        L.ctsm_b200_finalize(ctx)
    assert len(fs) > 1000 and len(fns) > 1000


@pytest.mark.parametrize("mem", [abi.MEM_DEVICE, abi.MEM_HOST])
@pytest.mark.parametrize("aerosol", [1, 0])
def test_snow_water_bit_exact(oracle_lib, mem, aerosol):
    L = abi.lib()
    sg, S = case(6000, 911)
    prm = abi.default_params()
    prm.snicar_use_aerosol = aerosol
    fs, fns = snow_filters(oracle_lib, sg, S)
    ref, got = copy_state(S), copy_state(S)
    rc, st = run_snow_water(oracle_lib, prm, sg, ref, fs, fns)
    assert rc == 0, st.msg
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        rc, st = gpu_snow_water(L, ctx, sg, got, mem, fs, fns)
        assert rc == 0, st.msg
    finally:
        L.ctsm_b200_finalize(ctx)
    for fsd in abi.FIELDS["snowwater"]:
        assert np.array_equal(got[fsd.name], ref[fsd.name], equal_nan=True), fsd.name
    assert (ref["qflx_snow_percolation"][:, fs - 1] > 0).sum() > 500


def test_snow_water_failure_empty_filters_and_clump_bounds(oracle_lib):
    L = abi.lib()
    sg, S = case(1500, 921)
    prm = abi.default_params()
    fs, fns = snow_filters(oracle_lib, sg, S)
    ref = copy_state(S)
    rc, st = run_snow_water(oracle_lib, prm, sg, ref, fs, fns)
    assert rc == 0
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        got = copy_state(S)
        for kb, fl in driver.make_slabs(sg, 4):                      # clump by clump (bounds != alloc), host arrays
            ks = fs[(fs >= kb.begc) & (fs <= kb.endc)]
            kn = fns[(fns >= kb.begc) & (fns <= kb.endc)]
            rc, st = gpu_snow_water(L, ctx, sg, got, abi.MEM_HOST, ks, kn, bounds=kb)
            assert rc == 0, st.msg
        for fsd in abi.FIELDS["snowwater"]:
            assert np.array_equal(got[fsd.name], ref[fsd.name], equal_nan=True), fsd.name
        # the reference's endrun: top-layer ice driven significantly negative
        bad = copy_state(S)
        bad["qflx_solidevap_from_top_layer"][fs[17] - 1] = 1.0
        bad["qflx_liqevap_from_top_layer"][fs[40] - 1] = 1.0
        rc, st = gpu_snow_water(L, ctx, sg, bad, abi.MEM_HOST, fs, fns)
        assert rc == 18 and st.subgrid_index == fs[17] and b"h2osoi_ice has gone significantly negative" in st.msg
        bad = copy_state(S)
        bad["qflx_liqevap_from_top_layer"][fs[40] - 1] = 1.0
        rc, st = gpu_snow_water(L, ctx, sg, bad, abi.MEM_HOST, fs, fns)
        assert rc == 18 and st.subgrid_index == fs[40] and b"h2osoi_liq has gone significantly negative" in st.msg
    finally:
        L.ctsm_b200_finalize(ctx)


def compare_layers(got, ref, S, fs, worst):
    c = fs - 1
    assert np.array_equal(got["snl"], ref["snl"]), "layer counts differ on %d columns" % int((got["snl"] != ref["snl"]).sum())
    for fsd in abi.FIELDS["snowlayers"]:
        a, b = got[fsd.name], ref[fsd.name]
        if fsd.intent == "IN":
            assert np.array_equal(a, S[fsd.name], equal_nan=True), "input %s was modified" % fsd.name
            continue
        if fsd.ctype == "int":
            assert np.array_equal(a, b), fsd.name
            continue
        fin = np.abs(b) < 1e30
        assert np.array_equal(fin, np.abs(a) < 1e30), "%s: fill pattern differs" % fsd.name
        assert np.array_equal(a == 0.0, b == 0.0), "%s: zero pattern differs" % fsd.name
        if fsd.name == "t_soisno":                                  # (a temperature's error is judged on the Kelvin scale)
            e = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1.0)))
        else:
            scale = float(np.max(np.abs(b[fin][np.abs(b[fin]) > 0]))) if (np.abs(b[fin]) > 0).any() else 1.0
            e = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-9 * scale)))
        worst[fsd.name] = e
        assert e <= RTOL, (fsd.name, e)
    untouched = np.ones(len(S["snl"]), dtype=bool)
    untouched[c] = False
    for k in ("dz", "h2osoi_ice", "h2osoi_liq", "t_soisno", "snw_rds", "mss_dst3", "zi"):
        assert np.array_equal(got[k][..., untouched], S[k][..., untouched]), k


@pytest.mark.parametrize("mem", [abi.MEM_DEVICE, abi.MEM_HOST])
@pytest.mark.parametrize("method,wind,subgrid", [(2, 1, 1), (1, 0, 0)], ids=["vionnet_wind_subgrid", "anderson_nowind_iceold"])
def test_snow_layers_match_oracle(oracle_lib, mem, method, wind, subgrid):
    L = abi.lib()
    sg, S = case(6000, 931)
    prm = abi.default_params()
    prm.snow_overburden_compaction_method, prm.wind_dependent_snow_density, prm.use_subgrid_fluxes = method, wind, subgrid
    fs, _ = snow_filters(oracle_lib, sg, S)
    ref, got = copy_state(S), copy_state(S)
    rc, st = run_snow_layers(oracle_lib, prm, sg, ref, fs)
    assert rc == 0
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        rc, st = gpu_snow_layers(L, ctx, sg, got, mem, fs)
        assert rc == 0, st.msg
    finally:
        L.ctsm_b200_finalize(ctx)
    worst = {}
    compare_layers(got, ref, S, fs, worst)
    c = fs - 1
    print("snow layers worst:", sorted(worst.items(), key=lambda kv: -kv[1])[:5], "merged", int((ref["snl"][c] > S["snl"][c]).sum()),
          "split", int((ref["snl"][c] < S["snl"][c]).sum()), "gone", int((ref["snl"][c] == 0).sum()))


def test_snow_layers_clump_bounds_empty_and_refusals(oracle_lib):
    L = abi.lib()
    sg, S = case(1500, 941)
    prm = abi.default_params()
    fs, _ = snow_filters(oracle_lib, sg, S)
    ref = copy_state(S)
    rc, st = run_snow_layers(oracle_lib, prm, sg, ref, fs)
    assert rc == 0
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        got = copy_state(S)
        rc, st = gpu_snow_layers(L, ctx, sg, got, abi.MEM_HOST, fs[:0])
        assert rc == 0
        for k in S:
            assert np.array_equal(got[k], S[k], equal_nan=True), k
        for kb, fl in driver.make_slabs(sg, 4):
            rc, st = gpu_snow_layers(L, ctx, sg, got, abi.MEM_HOST, fs[(fs >= kb.begc) & (fs <= kb.endc)], bounds=kb)
            assert rc == 0, st.msg
        compare_layers(got, ref, S, fs, {})
        bad = copy_state(S)
        bad["lun_itype"][fs[9] - 1] = 8
        rc, st = gpu_snow_layers(L, ctx, sg, bad, abi.MEM_HOST, fs)
        assert rc == 16 and st.subgrid_index == fs[9]
        bad = copy_state(S)
        bad["lun_itype"][fs[9] - 1] = 5
        rc, st = gpu_snow_layers(L, ctx, sg, bad, abi.MEM_HOST, fs)
        assert rc == 2 and st.subgrid_index == fs[9]
    finally:
        L.ctsm_b200_finalize(ctx)


def test_snow_sequence_of_hydrology_no_drainage(oracle_lib):
    """BuildSnowFilter -> SnowWater -> infiltration chain -> ... -> snow-layer update -> BuildSnowFilter, device-resident through
    driver.HotPath (HydrologyNoDrainageMod.F90:279-402 without the soil routines between), against the same sequence of the oracle:
    the second snow filter (built on the device from the new snl) is identical."""
    import torch
    from tests.test_oracle_hydrology import run_infiltration
    sg, S = case(3000, 951)
    prm = abi.default_params()
    fs, fns = snow_filters(oracle_lib, sg, S)
    ref = copy_state(S)
    assert run_snow_water(oracle_lib, prm, sg, ref, fs, fns)[0] == 0
    assert run_infiltration(oracle_lib, prm, sg, ref) == 0
    assert run_snow_layers(oracle_lib, prm, sg, ref, fs)[0] == 0
    fs2, fns2 = snow_filters(oracle_lib, sg, ref)
    ctx = driver.Context(prm)
    try:
        routines = ("snowwater", "infiltration", "snowlayers")
        names = sorted({f.name for g in routines for f in abi.FIELDS[g]})
        D = {k: torch.from_numpy(np.ascontiguousarray(S[k])).cuda() for k in names}
        hp = driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, routines)
        hp.step()
        ctx.sync()
        g2, gn2 = hp.BuildSnowFilter()
        got = {k: (D[k].cpu().numpy() if k in D else S[k]) for k in S}
    finally:
        ctx.close()
    assert np.array_equal(g2, fs2) and np.array_equal(gn2, fns2)
    assert np.array_equal(got["snl"], ref["snl"])
    for k in ("qflx_rain_plus_snomelt", "qflx_snow_percolation", "qflx_snow_drain", "qflx_top_soil"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    for k in ("h2osoi_liq", "h2osoi_ice", "dz", "t_soisno", "qflx_infl", "h2osfc", "mss_bcphi", "int_snow"):
        fin = np.abs(ref[k]) < 1e30
        e = np.max(np.abs(got[k][fin] - ref[k][fin]) / np.maximum(np.abs(ref[k][fin]), 1e-6 * np.max(np.abs(ref[k][fin]))))
        assert e <= RTOL, (k, e)


@pytest.mark.parametrize("mem", [abi.MEM_DEVICE, abi.MEM_HOST])
@pytest.mark.parametrize("reset,reset_glc,nstep", [(0, 0, 100), (1, 1, 10)], ids=["capping", "reset_active"])
def test_snow_capping_bit_exact(oracle_lib, mem, reset, reset_glc, nstep):
    """SnowCapping has no transcendentals: identical bits in every field"""
    from tests.test_oracle_snow import capping_case, run_snow_capping
    L = abi.lib()
    sg, S = capping_case(5000, 961)
    prm = abi.default_params()
    prm.reset_snow, prm.reset_snow_glc, prm.reset_snow_glc_ela = reset, reset_glc, 1500.0
    fs, _ = snow_filters(oracle_lib, sg, S)
    fi = sg.filters["nolakec"]
    ref, got = copy_state(S), copy_state(S)
    rc, st = run_snow_capping(oracle_lib, prm, sg, ref, fi, fs, nstep)
    assert rc == 0
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        def call(S_, fi_, fs_, bounds=None):
            b = C.byref(bounds if bounds is not None else sg.bounds)
            return gpu_call(L, ctx, "snowcapping", sg, S_, mem, [fi_, fs_], lambda f, fl, st: L.ctsm_b200_snow_capping(
                ctx, b, len(fi_), abi.i32p(fl[0]), len(fs_), abi.i32p(fl[1]), C.byref(f), nstep, mem, C.byref(st)))
        if mem == abi.MEM_HOST:
            for kb, fl in driver.make_slabs(sg, 3):
                rc, st = call(got, fl["nolakec"], fs[(fs >= kb.begc) & (fs <= kb.endc)], kb)
                assert rc == 0, st.msg
        else:
            rc, st = call(got, fi, fs)
            assert rc == 0, st.msg
        for f in abi.FIELDS["snowcapping"]:
            assert np.array_equal(got[f.name], ref[f.name], equal_nan=True), f.name
        capped = ((ref["qflx_snwcp_ice"] > 0) & (ref["qflx_snwcp_ice"] < 1e30)) | ((ref["qflx_snwcp_discarded_ice"] > 0) & (ref["qflx_snwcp_discarded_ice"] < 1e30))
        assert capped.sum() > 50
        bad = copy_state(S)                                     # negative mass remaining: a bottom layer whose liquid is negative
        bad["h2osoi_liq"][11, np.nonzero(capped)[0][3]] = -1.0
        rc, st = call(bad, fi, fs)
        assert rc == 18 and st.subgrid_index == np.nonzero(capped)[0][3] + 1 and b"capping procedure failed" in st.msg
    finally:
        L.ctsm_b200_finalize(ctx)
