"""BeginWaterColumnBalance (BalanceCheckMod.F90:171; ComputeWaterMassNonLake, TotalWaterAndHeatMod.F90:92): the oracle is
checked on the CPU against an independent numpy statement of the column water mass; the CUDA kernel must then agree with
the oracle bit for bit (sums in the reference's order)."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from oracle import oracle
from tests.util import copy_state, group_arrays, to_device

BASELINE = 5000.0      # aquifer_water_baseline (WaterStateType: aquifer_water_baseline = 5000 mm)


def _case(n, seed):
    sg, S = synthetic_canopy.make_full_case(n, seed=seed)
    synthetic_canopy.waterbalance_state(sg, S, np.random.Generator(np.random.PCG64(seed + 3)))
    return sg, S


def _oracle(sg, S):
    OL = oracle.lib()
    st = abi.Status()
    f = abi.make_struct("waterbalance", S, sg.bounds)
    fc = sg.filters["nolakec"]
    return OL.oracle_begin_water_column_balance(C.byref(sg.bounds), len(fc), abi.i32p(fc), C.byref(f), BASELINE, C.byref(st))


def test_oracle_water_mass_matches_numpy():
    sg, S = _case(300, 81)
    before = S["begwb"].copy()
    assert _oracle(sg, S) == 0
    c = sg.filters["nolakec"] - 1
    lo = -11
    lev = np.arange(lo, 26)[:, None]
    snow = (lev >= S["snl"][None, :] + 1) & (lev <= 0)
    soil = lev >= 1
    liq = (S["h2osoi_liq"] * (snow | soil)).sum(0)
    ice = (S["h2osoi_ice"] * (snow | soil)).sum(0) + S["excess_ice"].sum(0)
    w = S["wtcol"] * S["patch_active"]
    can = np.zeros(sg.ncol); np.add.at(can, S["column"] - 1, (S["liqcan"] + S["snocan"]) * w)
    want = liq + ice + can + S["h2osno_no_layers"] + S["h2osfc"] + np.where(S["col_hydrologically_active"] != 0, S["wa"] - BASELINE, 0.0)
    assert np.max(np.abs(S["begwb"][c] - want[c]) / np.maximum(np.abs(want[c]), 1.0)) < 1e-13
    sno = S["h2osno_no_layers"] + ((S["h2osoi_liq"] + S["h2osoi_ice"]) * snow).sum(0)
    assert np.max(np.abs(S["h2osno_old"][c] - sno[c])) < 1e-10
    other = np.ones(sg.ncol, bool); other[c] = False
    assert np.array_equal(S["begwb"][other], before[other])          # lake columns untouched


@pytest.mark.gpu
@pytest.mark.parametrize("size,mem", [(64, abi.MEM_HOST), (3000, abi.MEM_DEVICE)])
def test_begin_water_column_balance_bitwise(gpu_ctx, size, mem):
    L, ctx, prm = gpu_ctx
    sg, S = _case(size, 83)
    ref, got = copy_state(S), copy_state(S)
    assert _oracle(sg, ref) == 0
    st = abi.Status()
    fc = sg.filters["nolakec"]
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(got, "waterbalance"))
        dfc = to_device({"f": fc})["f"]
        f = abi.make_struct("waterbalance", D, sg.bounds)
        assert L.ctsm_b200_begin_water_column_balance(ctx, C.byref(sg.bounds), len(fc), abi.i32p(dfc), C.byref(f), BASELINE, mem,
                                                      C.byref(st)) == 0
        assert L.ctsm_b200_sync(ctx, C.byref(st)) == 0
        for k, v in D.items():
            got[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct("waterbalance", got, sg.bounds)
        assert L.ctsm_b200_begin_water_column_balance(ctx, C.byref(sg.bounds), len(fc), abi.i32p(fc), C.byref(f), BASELINE, mem,
                                                      C.byref(st)) == 0
    for name in ("begwb", "h2osno_old"):
        assert np.array_equal(got[name], ref[name], equal_nan=True), name
    assert np.all(np.abs(ref["begwb"][fc - 1]) < 1e30)
