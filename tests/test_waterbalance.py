"""BeginWaterColumnBalance (BalanceCheckMod.F90:171) and WaterGridcellBalance (:132) - ComputeWaterMassNonLake /
ComputeWaterMassLake (TotalWaterAndHeatMod.F90:92,144) and c2g: the oracle is checked on the CPU against an independent
numpy statement of the column and gridcell water mass; the CUDA kernels must then agree with the oracle bit for bit
(sums in the reference's order)."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from oracle import oracle
from tests.util import copy_state, group_arrays, to_device

BASELINE = 5000.0      # aquifer_water_baseline (WaterStateType: aquifer_water_baseline = 5000 mm)


def _case(n, seed):
    sg, S = synthetic_canopy.make_full_case(n, seed=seed)
    synthetic_canopy.watergrid_state(sg, S, np.random.Generator(np.random.PCG64(seed + 3)))
    return sg, S


def _oracle(sg, S):
    OL = oracle.lib()
    st = abi.Status()
    f = abi.make_struct("waterbalance", S, sg.bounds)
    fc, fl = sg.filters["nolakec"], sg.filters["lakec"]
    return OL.oracle_begin_water_column_balance(C.byref(sg.bounds), len(fc), abi.i32p(fc), len(fl), abi.i32p(fl), C.byref(f),
                                                BASELINE, C.byref(st))


def _oracle_grid(sg, S, flag):
    OL = oracle.lib()
    st = abi.Status()
    f = abi.make_struct("watergridbalance", S, sg.bounds)
    fc, fl = sg.filters["nolakec"], sg.filters["lakec"]
    return OL.oracle_water_gridcell_balance(C.byref(sg.bounds), len(fc), abi.i32p(fc), len(fl), abi.i32p(fl), C.byref(f),
                                            BASELINE, flag, C.byref(st))


def _numpy_masses(sg, S):
    """column water mass, independently: (non-lake mass, lake layers mass, snow water, lake water)"""
    lev = np.arange(-11, 26)[:, None]
    snow = (lev >= S["snl"][None, :] + 1) & (lev <= 0)
    soil = lev >= 1
    liq = (S["h2osoi_liq"] * (snow | soil)).sum(0)
    ice = (S["h2osoi_ice"] * (snow | soil)).sum(0)
    w = S["wtcol"] * S["patch_active"]
    can = np.zeros(sg.ncol); np.add.at(can, S["column"] - 1, (S["liqcan"] + S["snocan"]) * w)
    nonlake = (liq + ice + S["excess_ice"].sum(0) + can + S["h2osno_no_layers"] + S["h2osfc"]
               + np.where(S["col_hydrologically_active"] != 0, S["wa"] - BASELINE, 0.0))
    lake_layers = liq + ice + S["h2osno_no_layers"]
    sno = S["h2osno_no_layers"] + ((S["h2osoi_liq"] + S["h2osoi_ice"]) * snow).sum(0)
    lake_water = (S["dz_lake"] * 1000.0).sum(0)
    return nonlake, lake_layers, sno, lake_water


def test_oracle_water_mass_matches_numpy():
    sg, S = _case(300, 81)
    before = S["begwb"].copy()
    assert _oracle(sg, S) == 0
    c, lk = sg.filters["nolakec"] - 1, sg.filters["lakec"] - 1
    assert len(lk) > 0
    nonlake, lake_layers, sno, _ = _numpy_masses(sg, S)
    assert np.max(np.abs(S["begwb"][c] - nonlake[c]) / np.maximum(np.abs(nonlake[c]), 1.0)) < 1e-13
    assert np.max(np.abs(S["begwb"][lk] - lake_layers[lk]) / np.maximum(np.abs(lake_layers[lk]), 1.0)) < 1e-13
    both = np.concatenate([c, lk])
    assert np.max(np.abs(S["h2osno_old"][both] - sno[both])) < 1e-10
    other = np.ones(sg.ncol, bool); other[both] = False
    assert np.array_equal(S["begwb"][other], before[other])          # inactive columns untouched


@pytest.mark.parametrize("flag", [0, 1])
def test_oracle_water_gridcell_balance_matches_numpy(flag):
    sg, S = _case(300, 85)
    assert _oracle_grid(sg, S, flag) == 0
    nonlake, lake_layers, _, lake_water = _numpy_masses(sg, S)
    wb_col = np.zeros(sg.ncol)
    c, lk = sg.filters["nolakec"] - 1, sg.filters["lakec"] - 1
    base = S["dynbal_baseline_liq"] + S["dynbal_baseline_ice"]
    wb_col[c] = nonlake[c] - base[c]
    wb_col[lk] = lake_layers[lk] + lake_water[lk] - base[lk]
    act = (S["col_active"] != 0) & (S["wtgcell"] != 0)
    num = np.zeros(sg.ngrc); den = np.zeros(sg.ngrc)
    np.add.at(num, sg.col_gridcell[act] - 1, (wb_col * S["wtgcell"])[act])
    np.add.at(den, sg.col_gridcell[act] - 1, S["wtgcell"][act])
    want = num / den - S["qflx_liq_dynbal_left_to_dribble"] - S["qflx_ice_dynbal_left_to_dribble"]
    got = S["endwb_grc"] if flag else S["begwb_grc"]
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) < 1e-13
    assert np.all((S["begwb_grc"] if flag else S["endwb_grc"]) == 1.0e36)     # only the flagged field is written


def test_oracle_c2g_reports_overweight_gridcell():
    sg, S = _case(64, 87)
    g = 5
    cols = np.nonzero(sg.col_gridcell == g)[0]
    S["wtgcell"][cols] = 0.8                                              # two or three columns: sum of weights > 1
    if len(cols) < 2:
        S["wtgcell"][cols] = 1.5
    OL = oracle.lib()
    st = abi.Status()
    f = abi.make_struct("watergridbalance", S, sg.bounds)
    fc, fl = sg.filters["nolakec"], sg.filters["lakec"]
    rc = OL.oracle_water_gridcell_balance(C.byref(sg.bounds), len(fc), abi.i32p(fc), len(fl), abi.i32p(fl), C.byref(f), BASELINE, 0,
                                          C.byref(st))
    assert rc == 20 and st.subgrid_index == g and st.subgrid_level == 1


@pytest.mark.gpu
@pytest.mark.parametrize("size,mem", [(64, abi.MEM_HOST), (3000, abi.MEM_DEVICE)])
def test_begin_water_column_balance_bitwise(gpu_ctx, size, mem):
    L, ctx, prm = gpu_ctx
    sg, S = _case(size, 83)
    ref, got = copy_state(S), copy_state(S)
    assert _oracle(sg, ref) == 0
    st = abi.Status()
    fc, fl = sg.filters["nolakec"], sg.filters["lakec"]
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(got, "waterbalance"))
        dfc, dfl = to_device({"f": fc})["f"], to_device({"f": fl})["f"]
        f = abi.make_struct("waterbalance", D, sg.bounds)
        assert L.ctsm_b200_begin_water_column_balance(ctx, C.byref(sg.bounds), len(fc), abi.i32p(dfc), len(fl), abi.i32p(dfl),
                                                      C.byref(f), BASELINE, mem, C.byref(st)) == 0
        assert L.ctsm_b200_sync(ctx, C.byref(st)) == 0
        for k, v in D.items():
            got[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct("waterbalance", got, sg.bounds)
        assert L.ctsm_b200_begin_water_column_balance(ctx, C.byref(sg.bounds), len(fc), abi.i32p(fc), len(fl), abi.i32p(fl),
                                                      C.byref(f), BASELINE, mem, C.byref(st)) == 0
    for name in ("begwb", "h2osno_old"):
        assert np.array_equal(got[name], ref[name], equal_nan=True), name
    assert np.all(np.abs(ref["begwb"][fc - 1]) < 1e30) and np.all(np.abs(ref["begwb"][fl - 1]) < 1e30)


@pytest.mark.gpu
@pytest.mark.parametrize("size,mem,flag", [(64, abi.MEM_HOST, 0), (3000, abi.MEM_DEVICE, 1), (3000, abi.MEM_HOST, 1)])
def test_water_gridcell_balance_bitwise(gpu_ctx, size, mem, flag):
    L, ctx, prm = gpu_ctx
    sg, S = _case(size, 89)
    ref, got = copy_state(S), copy_state(S)
    assert _oracle_grid(sg, ref, flag) == 0
    st = abi.Status()
    fc, fl = sg.filters["nolakec"], sg.filters["lakec"]
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(got, "watergridbalance"))
        dfc, dfl = to_device({"f": fc})["f"], to_device({"f": fl})["f"]
        f = abi.make_struct("watergridbalance", D, sg.bounds)
        assert L.ctsm_b200_water_gridcell_balance(ctx, C.byref(sg.bounds), len(fc), abi.i32p(dfc), len(fl), abi.i32p(dfl),
                                                  C.byref(f), BASELINE, flag, mem, C.byref(st)) == 0
        assert L.ctsm_b200_sync(ctx, C.byref(st)) == 0
        for k, v in D.items():
            got[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct("watergridbalance", got, sg.bounds)
        assert L.ctsm_b200_water_gridcell_balance(ctx, C.byref(sg.bounds), len(fc), abi.i32p(fc), len(fl), abi.i32p(fl),
                                                  C.byref(f), BASELINE, flag, mem, C.byref(st)) == 0
    for name in ("begwb_grc", "endwb_grc"):
        assert np.array_equal(got[name], ref[name]), name
    assert np.all(np.abs((ref["endwb_grc"] if flag else ref["begwb_grc"])) < 1e30)


@pytest.mark.gpu
def test_water_gridcell_balance_overweight_is_reported(gpu_ctx):
    L, ctx, prm = gpu_ctx
    sg, S = _case(64, 87)
    g = 5
    cols = np.nonzero(sg.col_gridcell == g)[0]
    S["wtgcell"][cols] = 1.5
    st = abi.Status()
    fc, fl = sg.filters["nolakec"], sg.filters["lakec"]
    f = abi.make_struct("watergridbalance", S, sg.bounds)
    rc = L.ctsm_b200_water_gridcell_balance(ctx, C.byref(sg.bounds), len(fc), abi.i32p(fc), len(fl), abi.i32p(fl), C.byref(f),
                                            BASELINE, 0, abi.MEM_HOST, C.byref(st))
    assert rc == 20 and st.subgrid_index == g and st.subgrid_level == 1
