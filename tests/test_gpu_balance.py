"""GPU parity: BalanceCheck/EnergyBalanceCheck and the PHS root-water sink vs the CPU oracle.
The residuals are differences of O(1e2..1e3) operands: agreement is required to 1e-10 of the operand
scale (|begwb|, |fluxes|), the clump maxima and maxloc indices, warnings and the abort decision must
be identical."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from tests.util import copy_state, group_arrays, to_device

pytestmark = pytest.mark.gpu


def _case(n, seed, noise=1e-11):
    sg, S = synthetic_canopy.make_full_case(n, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    synthetic_canopy.balance_state(sg, S, rng, noise)
    return sg, S


def _balance(fn, handle, prm, sg, S, danstep, mem=None):
    rep, st = abi.BalanceReport(), abi.Status()
    allc = np.arange(1, sg.ncol + 1, dtype=np.int32)
    f = abi.make_struct("balancecheck", S, sg.bounds)
    if mem is None:
        rc = fn(C.byref(prm), C.byref(sg.bounds), len(allc), abi.i32p(allc), C.byref(f), danstep, C.byref(rep), C.byref(st))
    else:
        rc = fn(handle, C.byref(sg.bounds), len(allc), abi.i32p(allc), C.byref(f), danstep, mem, C.byref(rep), C.byref(st))
    return rc, rep, st


@pytest.mark.parametrize("noise,danstep", [(1e-11, 100), (1e-8, 100), (1e-3, 2), (1e-3, 100)])
def test_balancecheck_matches_oracle(gpu_ctx, oracle_lib, noise, danstep):
    L, ctx, prm = gpu_ctx
    skip = L.ctsm_b200_balancecheck_init(ctx)
    assert skip == oracle_lib.oracle_balancecheck_skip_steps(prm.dtime) == 3
    prm2 = abi.default_params(); prm2.balance_skip_steps = skip
    sg, S = _case(300, 31, noise)
    ref, got = copy_state(S), copy_state(S)
    rc_ref, rep_ref, st_ref = _balance(oracle_lib.oracle_balancecheck, None, prm2, sg, ref, danstep)
    rc, rep, st = _balance(L.ctsm_b200_balancecheck, ctx, None, sg, got, danstep, abi.MEM_HOST)
    assert rc == rc_ref and rep.abort_kind == rep_ref.abort_kind
    assert list(rep.warn) == list(rep_ref.warn) and list(rep.index) == list(rep_ref.index)
    for a, b in zip(rep.max_abs, rep_ref.max_abs):
        assert abs(a - b) <= 1e-9 * max(abs(b), 1e-12) + 1e-13
    if rc:
        assert rc == 20 and (st.subgrid_index, st.subgrid_level) == (st_ref.subgrid_index, st_ref.subgrid_level)
        assert noise == 1e-3 and danstep > skip
    scale = {"errh2o": 3000.0, "errh2osno": 300.0, "errh2o_grc": 3000.0, "errsol": 800.0, "errlon": 500.0, "errseb": 1000.0}
    for name in ("errh2o", "errh2osno", "errh2o_grc", "errsol", "errlon", "errseb", "netrad", "snow_sources", "snow_sinks"):
        a, b = got[name], ref[name]
        assert np.all(np.abs(a - b) <= 1e-10 * np.maximum(np.abs(b), scale.get(name, 0.0))), name


def test_balancecheck_before_init_is_an_error(oracle_lib):
    L = abi.lib()
    prm = abi.default_params()
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    sg, S = _case(64, 3)
    rc, rep, st = _balance(L.ctsm_b200_balancecheck, ctx, None, sg, S, 10, abi.MEM_HOST)
    assert rc == 2       # the reference aborts in GetBalanceCheckSkipSteps (BalanceCheckMod.F90:117-128)
    L.ctsm_b200_finalize(ctx)


@pytest.mark.parametrize("mem,sink_warp", [(abi.MEM_HOST, 1), (abi.MEM_DEVICE, 1), (abi.MEM_HOST, 0)])
def test_vert_tran_sink_matches_oracle(gpu_ctx, oracle_lib, mem, sink_warp):
    """both kernels (one warp per column: default; one thread per column) against the oracle, bit for bit; inactive patches,
    zero-weight patches and patches without exposed vegetation are present"""
    L, ctx, prm = gpu_ctx
    assert L.ctsm_b200_set_sink_tuning(ctx, sink_warp) == 0
    sg, S = _case(500, 41)
    S["wtcol"][::19] = 0.0
    # k_soil_root / vegwp as CanopyFluxes leaves them
    from tests.test_gpu_canopy import run_oracle
    assert run_oracle(oracle_lib, prm, sg, S)[0] == 0
    ref, got = copy_state(S), copy_state(S)
    fh = sg.filters["hydrologyc"]
    f = abi.make_struct("plantsink", ref, sg.bounds)
    assert oracle_lib.oracle_vert_tran_sink_hydstress(C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(f)) == 0
    st = abi.Status()
    if mem == abi.MEM_DEVICE:
        D = to_device(group_arrays(got, "plantsink"))
        dfh = to_device({"f": fh})["f"]
        f = abi.make_struct("plantsink", D, sg.bounds)
        assert L.ctsm_b200_vert_tran_sink_hydstress(ctx, C.byref(sg.bounds), len(fh), abi.i32p(dfh), C.byref(f), mem, C.byref(st)) == 0
        assert L.ctsm_b200_sync(ctx, C.byref(st)) == 0
        for k, v in D.items():
            got[k][...] = v.cpu().numpy()
    else:
        f = abi.make_struct("plantsink", got, sg.bounds)
        assert L.ctsm_b200_vert_tran_sink_hydstress(ctx, C.byref(sg.bounds), len(fh), abi.i32p(fh), C.byref(f), mem, C.byref(st)) == 0
    for name in ("qflx_rootsoi", "qflx_phs_neg", "qflx_hydr_redist"):
        assert np.array_equal(got[name], ref[name]), name     # same operations in the same order: bit-identical
    assert np.abs(ref["qflx_rootsoi"][:, fh - 1]).max() > 0
    assert L.ctsm_b200_set_sink_tuning(ctx, 1) == 0
