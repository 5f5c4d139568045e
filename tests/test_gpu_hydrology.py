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
