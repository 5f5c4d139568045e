"""GPU parity of the soil-moisture-stress configuration (use_hydrstress = .false.): the default
root-water sink alone (bit for bit) and the whole step - CanopyFluxes with Photosynthesis for sunlit / shaded leaves ->
SoilTemperature -> SoilFluxes -> patch2col -> Compute_EffecRootFrac_And_VertTranSink_Default -> SoilWater -> BalanceCheck -
against the oracle, through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, driver, synthetic_canopy
from tests.util import copy_state, to_device, group_arrays
from tests.test_gpu_canopy import compare as compare_canopy

pytestmark = pytest.mark.gpu
RTOL = 1e-10
STEP_GROUPS = ("soiltemperature", "soilwater", "canopyfluxes", "plantsink", "balancecheck", "soilfluxes", "patch2col")


def _case(n, seed):
    sg, S = synthetic_canopy.make_full_case(n, seed=seed)
    synthetic_canopy.balance_state(sg, S, np.random.Generator(np.random.PCG64(seed + 1)), 1.0e-11)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(seed + 2)))
    synthetic_canopy.bare_ground_state(sg, S)
    return sg, S


@pytest.mark.parametrize("mem", [abi.MEM_DEVICE, abi.MEM_HOST])
def test_default_sink_bit_exact(oracle_lib, mem):
    L = abi.lib()
    prm = abi.default_params()
    prm.use_hydrstress = 0
    sg, S = _case(3000, 301)
    S["qflx_tran_veg_col"] = np.random.Generator(np.random.PCG64(5)).uniform(0.0, 3.0e-5, sg.ncol)
    ex = sg.filters["exposedvegp"] - 1
    S["qflx_tran_veg"][ex] = np.random.Generator(np.random.PCG64(6)).uniform(0.0, 4.0e-5, len(ex))
    S["rootr"][:, ex] = np.random.Generator(np.random.PCG64(7)).uniform(0.0, 0.2, (S["rootr"].shape[0], len(ex)))
    S["qflx_tran_veg"][ex[::7]] = 0.0                     # some columns without any transpiration (temp == 0)
    ref, got = copy_state(S), copy_state(S)
    fh = sg.filters["hydrologyc"]
    f = abi.make_struct("plantsinkdefault", ref, sg.bounds)
    assert oracle_lib.oracle_vert_tran_sink_default(C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(f)) == 0
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    st = abi.Status()
    try:
        if mem == abi.MEM_DEVICE:
            D = to_device(group_arrays(got, "plantsinkdefault"))
            dfh = to_device({"f": fh})["f"]
            f = abi.make_struct("plantsinkdefault", D, sg.bounds)
            assert L.ctsm_b200_vert_tran_sink_default(ctx, C.byref(sg.bounds), len(fh), abi.i32p(dfh), C.byref(f), mem, C.byref(st)) == 0
            assert L.ctsm_b200_sync(ctx, C.byref(st)) == 0
            for k, v in D.items():
                got[k][...] = v.cpu().numpy()
        else:
            f = abi.make_struct("plantsinkdefault", got, sg.bounds)
            assert L.ctsm_b200_vert_tran_sink_default(ctx, C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(f), mem, C.byref(st)) == 0
    finally:
        L.ctsm_b200_finalize(ctx)
    for fs in abi.FIELDS["plantsinkdefault"]:
        assert np.array_equal(got[fs.name], ref[fs.name], equal_nan=True), fs.name
    assert np.any(ref["qflx_rootsoi"][:, fh - 1] > 0.0)


@pytest.mark.parametrize("mtd", [2, 1], ids=["medlyn", "ballberry"])
def test_full_step_without_hydraulic_stress(oracle_lib, mtd):
    from oracle import oracle
    import torch
    sg, S = _case(4000, 311 + mtd)
    prm = abi.default_params()
    prm.use_hydrstress, prm.stomatalcond_mtd = 0, mtd
    ref = copy_state(S)
    oprm = abi.default_params()
    oprm.use_hydrstress, oprm.stomatalcond_mtd = 0, mtd
    oprm.balance_skip_steps = int(oracle_lib.oracle_balancecheck_skip_steps(oprm.dtime))
    clumps, keep = oracle.make_clumps(sg, 8)
    structs = [abi.make_struct(g, ref, sg.bounds) for g in STEP_GROUPS]
    fd = abi.make_struct("plantsinkdefault", ref, sg.bounds)
    oracle_lib.oracle_set_plantsink_default(C.byref(fd))
    try:
        assert oracle_lib.oracle_fullstep_clumps(C.byref(oprm), len(clumps), clumps, *[C.byref(x) for x in structs], 1, 127) == 0
    finally:
        oracle_lib.oracle_set_plantsink_default(None)
    ref_c = copy_state(S)                                    # CanopyFluxes alone, for the per-patch comparison rule
    fc = abi.make_struct("canopyfluxes", ref_c, sg.bounds)
    assert oracle_lib.oracle_step_clumps(C.byref(oprm), len(clumps), clumps, None, None, C.byref(fc), 4) == 0
    ctx = driver.Context(prm)
    try:
        names = sorted({fs.name for g in driver.ROUTINES for fs in abi.FIELDS[g]} | {fs.name for fs in abi.FIELDS["plantsinkdefault"]})
        D = {k: torch.from_numpy(np.ascontiguousarray(S[k])).cuda() for k in names}
        driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, ("canopyfluxes",)).step()
        ctx.sync()
        got_c = {k: (D[k].cpu().numpy() if k in D else S[k]) for k in S}
        hp = driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, driver.ROUTINES[1:])
        assert hp.group_of["plantsink"] == "plantsinkdefault"
        hp.step()
        ctx.sync()
        got = {k: (D[k].cpu().numpy() if k in D else S[k]) for k in S}
    finally:
        ctx.close()
    worst, ntie = compare_canopy(sg, got_c, ref_c, S, check_inputs=False)
    fe = sg.filters["exposedvegp"] - 1
    # the non-PHS outputs really were produced
    day = S["parsha_z"][0, fe] > 0.0
    assert np.all(got_c["an"][0, fe] < 1e30) and np.any(got_c["ag"][0, fe][day] > 0.0)
    loose_p = np.zeros(sg.npatch, dtype=bool)
    loose_p[fe[(got_c["num_iter"][fe] != ref_c["num_iter"][fe]) | (ref_c["num_iter"][fe] >= 41)]] = True
    loose_c = np.zeros(sg.ncol, dtype=bool)
    loose_c[S["column"][loose_p] - 1] = True
    loose_g = np.zeros(sg.ngrc, dtype=bool)
    loose_g[sg.col_gridcell[loose_c] - 1] = True
    skip_of = {"PATCH": loose_c[S["column"] - 1], "COL": loose_c, "GRC": loose_g}
    step_worst = {}
    for g in ("soiltemperature", "soilfluxes", "patch2col", "plantsinkdefault", "soilwater", "balancecheck"):
        for fs in abi.FIELDS[g]:
            if fs.intent == "IN" or fs.sub not in skip_of or fs.name.startswith("err"):
                continue
            a, b = got[fs.name], ref[fs.name]
            keepm = ~skip_of[fs.sub]
            if fs.ctype == "int":
                assert np.array_equal(a[..., keepm], b[..., keepm]), fs.name
                continue
            fin = np.abs(b) < 1e30
            assert np.array_equal(fin, np.abs(a) < 1e30), fs.name
            bb, aa = np.where(fin, b, 0.0)[..., keepm], np.where(fin, a, 0.0)[..., keepm]
            scale = float(np.max(np.abs(bb))) if bb.size else 0.0
            e = float(np.max(np.abs(aa - bb) / np.maximum(np.abs(bb), 1e-2 * scale + 1e-300))) if bb.size else 0.0
            step_worst[fs.name] = max(step_worst.get(fs.name, 0.0), e)
    bad = {k: v for k, v in step_worst.items() if not v <= RTOL}
    print("no-PHS step: canopy worst", sorted(worst.items(), key=lambda kv: -kv[1])[:4], "ties", ntie,
          "rest worst", sorted(step_worst.items(), key=lambda kv: -kv[1])[:4])
    assert not bad, bad
    assert np.any(ref["qflx_rootsoi"] > 0.0)
