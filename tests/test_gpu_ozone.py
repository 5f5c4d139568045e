"""GPU parity of CalcOzoneUptake / CalcOzoneStress through the C ABI: identical bits (no transcendentals)."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, driver
from tests.util import copy_state, to_device, group_arrays
from tests.test_oracle_ozone import case, run_uptake, run_stress

pytestmark = pytest.mark.gpu


def _gpu(L, ctx, sg, S, mem, filters, call):
    st = abi.Status()
    z = np.zeros(1, dtype=np.int32)
    filters = [f if len(f) else z for f in filters]
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(S, "ozone"))
        f = abi.make_struct("ozone", D, sg.bounds)
        rc = call(f, [to_device({"f": v})["f"] for v in filters], st)
        if rc == 0:
            rc = L.ctsm_b200_sync(ctx, C.byref(st))
        for k, v in D.items():
            S[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct("ozone", S, sg.bounds)
        rc = call(f, filters, st)
    return rc, st


@pytest.mark.parametrize("mem", [abi.MEM_DEVICE, abi.MEM_HOST])
def test_ozone_uptake_and_stress_bit_exact(oracle_lib, mem):
    L = abi.lib()
    sg, S = case(6000, 1021)
    prm = abi.default_params()
    fe, fn = sg.filters["exposedvegp"], sg.filters["noexposedvegp"]
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    try:
        ref, got = copy_state(S), copy_state(S)
        assert run_uptake(oracle_lib, prm, sg, ref) == 0
        if mem == abi.MEM_HOST:                                   # clump by clump, bounds != alloc
            for kb, fl in driver.make_slabs(sg, 3):
                ke = fl["exposedvegp"]
                rc, st = _gpu(L, ctx, sg, got, mem, [ke], lambda f, x, st: L.ctsm_b200_calc_ozone_uptake(
                    ctx, C.byref(kb), len(ke), abi.i32p(x[0]), C.byref(f), mem, C.byref(st)))
                assert rc == 0, st.msg
        else:
            rc, st = _gpu(L, ctx, sg, got, mem, [fe], lambda f, x, st: L.ctsm_b200_calc_ozone_uptake(
                ctx, C.byref(sg.bounds), len(fe), abi.i32p(x[0]), C.byref(f), mem, C.byref(st)))
            assert rc == 0, st.msg
        for f in abi.FIELDS["ozone"]:
            assert np.array_equal(got[f.name], ref[f.name], equal_nan=True), f.name
        for method in (1, 2):
            assert run_stress(oracle_lib, sg, ref, method) == 0
            rc, st = _gpu(L, ctx, sg, got, mem, [fe, fn], lambda f, x, st: L.ctsm_b200_calc_ozone_stress(
                ctx, C.byref(sg.bounds), len(fe), abi.i32p(x[0]), len(fn), abi.i32p(x[1]), method, 1, C.byref(f), mem, C.byref(st)))
            assert rc == 0, st.msg
            for f in abi.FIELDS["ozone"]:
                assert np.array_equal(got[f.name], ref[f.name], equal_nan=True), (method, f.name)
        keep = copy_state(got)                                    # Falk outside a LUNA step, empty filters: nothing changes
        rc, st = _gpu(L, ctx, sg, got, mem, [fe, fn], lambda f, x, st: L.ctsm_b200_calc_ozone_stress(
            ctx, C.byref(sg.bounds), len(fe), abi.i32p(x[0]), len(fn), abi.i32p(x[1]), 2, 0, C.byref(f), mem, C.byref(st)))
        assert rc == 0
        rc, st = _gpu(L, ctx, sg, got, mem, [fe[:0]], lambda f, x, st: L.ctsm_b200_calc_ozone_uptake(
            ctx, C.byref(sg.bounds), 0, abi.i32p(x[0]), C.byref(f), mem, C.byref(st)))
        assert rc == 0
        for k in keep:
            assert np.array_equal(got[k], keep[k], equal_nan=True), k
        st = abi.Status()
        f = abi.make_struct("ozone", got, sg.bounds)
        assert L.ctsm_b200_calc_ozone_stress(ctx, C.byref(sg.bounds), len(fe), abi.i32p(fe), len(fn), abi.i32p(fn), 3, 1, C.byref(f),
                                             abi.MEM_HOST, C.byref(st)) == 2
    finally:
        L.ctsm_b200_finalize(ctx)
    assert ((ref["o3coefvsun"][fe - 1] < 1.0) & (ref["o3coefvsun"][fe - 1] > 0.0)).sum() > 1000
