"""GPU parity: SoilFluxes (+ p2c) through the C ABI vs the CPU oracle.  The routine is sums, products and one sqrt(sqrt())
in a fixed order, without transcendentals, and the kernel keeps the reference's order of operations: every output must
agree BIT FOR BIT (integer/index work aside, the strictest bar of the suite)."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from tests.util import copy_state, group_arrays, to_device

pytestmark = pytest.mark.gpu


def _case(oracle_lib, prm, n, seed):
    """State as SoilFluxes finds it: CanopyFluxes and SoilTemperature (oracle) have run on the synthetic state."""
    from tests.test_gpu_canopy import run_oracle
    sg, S = synthetic_canopy.make_full_case(n, seed=seed)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(seed + 2)))
    assert run_oracle(oracle_lib, prm, sg, S)[0] == 0
    st = abi.Status()
    ft = abi.make_struct("soiltemperature", S, sg.bounds)
    fp, fc = sg.filters["nolakep"], sg.filters["nolakec"]
    assert oracle_lib.oracle_soiltemperature(C.byref(prm), C.byref(sg.bounds), len(fp), abi.i32p(fp), len(fc), abi.i32p(fc),
                                             C.byref(ft), C.byref(st)) == 0
    return sg, S


def _run_oracle(OL, prm, sg, S):
    st = abi.Status()
    f = abi.make_struct("soilfluxes", S, sg.bounds)
    fp, fc = sg.filters["nolakep"], sg.filters["nolakec"]
    return OL.oracle_soilfluxes(C.byref(prm), C.byref(sg.bounds), len(fc), abi.i32p(fc), len(fp), abi.i32p(fp), C.byref(f), C.byref(st)), st


def _run_gpu(L, ctx, sg, S, mem):
    st = abi.Status()
    fp, fc = sg.filters["nolakep"], sg.filters["nolakec"]
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(S, "soilfluxes"))
        dflt = to_device({"p": fp, "c": fc})
        f = abi.make_struct("soilfluxes", D, sg.bounds)
        rc = L.ctsm_b200_soilfluxes(ctx, C.byref(sg.bounds), len(fc), abi.i32p(dflt["c"]), len(fp), abi.i32p(dflt["p"]),
                                    C.byref(f), mem, C.byref(st))
        assert rc == 0
        rc = L.ctsm_b200_sync(ctx, C.byref(st))
        for k, v in D.items():
            S[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct("soilfluxes", S, sg.bounds)
        rc = L.ctsm_b200_soilfluxes(ctx, C.byref(sg.bounds), len(fc), abi.i32p(fc), len(fp), abi.i32p(fp), C.byref(f), mem,
                                    C.byref(st))
    return rc, st


@pytest.mark.parametrize("size,mem,seed", [(64, abi.MEM_HOST, 51), (64, abi.MEM_DEVICE, 52), (3000, abi.MEM_DEVICE, 53)])
def test_soilfluxes_matches_oracle_bitwise(gpu_ctx, oracle_lib, size, mem, seed):
    L, ctx, prm = gpu_ctx
    sg, S = _case(oracle_lib, prm, size, seed)
    ref, got = copy_state(S), copy_state(S)
    rc_ref, _ = _run_oracle(oracle_lib, prm, sg, ref)
    rc, st = _run_gpu(L, ctx, sg, got, mem)
    assert rc == rc_ref == 0, st.msg
    for fs in abi.FIELDS["soilfluxes"]:
        a, b = got[fs.name], ref[fs.name]
        assert np.array_equal(a, b, equal_nan=True), "%s differs (%s)" % (fs.name, fs.intent)
    # the routine really ran: the flux correction moved the ground fluxes and the partition closes
    fp = sg.filters["nolakep"] - 1
    assert np.any(got["eflx_sh_grnd"][fp] != S["eflx_sh_grnd"][fp])
    part = (got["qflx_liqevap_from_top_layer_patch"] + got["qflx_solidevap_from_top_layer_patch"]
            - got["qflx_liqdew_to_top_layer_patch"] - got["qflx_soliddew_to_top_layer_patch"])[fp]
    ok = np.abs(part - got["qflx_ev_snow"][fp]) <= 1e-12 * np.maximum(np.abs(got["qflx_ev_snow"][fp]), 1e-12)
    assert ok.mean() > 0.9        # except where the snow-evaporation limit of :283-292 redistributes the fluxes


def test_soilfluxes_clump_bounds_and_urban_refusal(gpu_ctx, oracle_lib):
    """Called with clump bounds on proc-sized arrays, and an urban column in the filter is refused like the other
    routines of the path refuse it (CTSM_ERR_URBAN, column index reported)."""
    L, ctx, prm = gpu_ctx
    sg, S = _case(oracle_lib, prm, 200, 61)
    ref, got = copy_state(S), copy_state(S)
    # second half of the gridcells as one clump
    from oracle import oracle
    clumps, keep = oracle.make_clumps(sg, 2)
    k = clumps[1]
    st = abi.Status()
    fr = abi.make_struct("soilfluxes", ref, sg.bounds)
    assert oracle_lib.oracle_soilfluxes(C.byref(prm), C.byref(k.bounds), k.num_nolakec, k.filter_nolakec, k.num_nolakep,
                                        k.filter_nolakep, C.byref(fr), C.byref(st)) == 0
    fg = abi.make_struct("soilfluxes", got, sg.bounds)
    assert L.ctsm_b200_soilfluxes(ctx, C.byref(k.bounds), k.num_nolakec, k.filter_nolakec, k.num_nolakep, k.filter_nolakep,
                                  C.byref(fg), abi.MEM_HOST, C.byref(st)) == 0
    for fs in abi.FIELDS["soilfluxes"]:
        assert np.array_equal(got[fs.name], ref[fs.name], equal_nan=True), fs.name
    # untouched outside the clump
    p0 = k.bounds.begp - 1
    assert np.array_equal(got["eflx_sh_grnd"][:p0], S["eflx_sh_grnd"][:p0])
    # urban refusal
    bad = copy_state(S)
    c = int(sg.filters["nolakec"][3])
    bad["lun_itype"][c - 1] = 8
    fb = abi.make_struct("soilfluxes", bad, sg.bounds)
    fp, fc = sg.filters["nolakep"], sg.filters["nolakec"]
    rc = L.ctsm_b200_soilfluxes(ctx, C.byref(sg.bounds), len(fc), abi.i32p(fc), len(fp), abi.i32p(fp), C.byref(fb), abi.MEM_HOST,
                                C.byref(st))
    assert rc == 16 and st.subgrid_index == c


@pytest.mark.parametrize("mem,warp", [(abi.MEM_HOST, 1), (abi.MEM_DEVICE, 1), (abi.MEM_DEVICE, 0)])
def test_patch2col_matches_oracle_bitwise(gpu_ctx, oracle_lib, mem, warp):
    """clm_drv_patch2col (clm_driver.F90:1655): eleven p2c averages, same accumulation order, bit-identical - the warp-per-column
    kernels (default) and the thread-per-column ones (ctsm_b200_set_sink_tuning)."""
    L, ctx, prm = gpu_ctx
    assert L.ctsm_b200_set_sink_tuning(ctx, warp) == 0
    sg, S = _case(oracle_lib, prm, 1500, 57)
    assert _run_oracle(oracle_lib, prm, sg, S)[0] == 0            # SoilFluxes first: its outputs are what gets averaged
    ref, got = copy_state(S), copy_state(S)
    allc = np.arange(sg.bounds.begc, sg.bounds.endc + 1, dtype=np.int32)
    fc = sg.filters["nolakec"]
    fr = abi.make_struct("patch2col", ref, sg.bounds)
    assert oracle_lib.oracle_patch2col(C.byref(sg.bounds), len(allc), abi.i32p(allc), len(fc), abi.i32p(fc), C.byref(fr)) == 0
    st = abi.Status()
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(got, "patch2col"))
        dflt = to_device({"a": allc, "c": fc})
        f = abi.make_struct("patch2col", D, sg.bounds)
        assert L.ctsm_b200_patch2col(ctx, C.byref(sg.bounds), len(allc), abi.i32p(dflt["a"]), len(fc), abi.i32p(dflt["c"]),
                                     C.byref(f), mem, C.byref(st)) == 0
        assert L.ctsm_b200_sync(ctx, C.byref(st)) == 0
        for k, v in D.items():
            got[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct("patch2col", got, sg.bounds)
        assert L.ctsm_b200_patch2col(ctx, C.byref(sg.bounds), len(allc), abi.i32p(allc), len(fc), abi.i32p(fc), C.byref(f), mem,
                                     C.byref(st)) == 0
    for fs in abi.FIELDS["patch2col"]:
        assert np.array_equal(got[fs.name], ref[fs.name], equal_nan=True), fs.name
    assert np.abs(ref["qflx_evap_soi_col"][fc - 1]).max() > 0
    assert L.ctsm_b200_set_sink_tuning(ctx, 1) == 0
