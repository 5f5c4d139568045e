"""Resident window (ctsm_b200_host_window_begin/_end): the seven-routine step with host-owned arrays, issued slab after
slab with asynchronous staging, must leave the host arrays bit-identical to the self-contained CTSM_MEM_HOST calls and to
the device-resident step; bytes moved are what the field table predicts (each IN/INOUT field up once, no OUT uploads)."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, driver, synthetic_canopy
from tests.util import copy_state

pytestmark = pytest.mark.gpu


def _case(n=600, seed=7):
    sg, S = synthetic_canopy.make_full_case(n, seed=seed)
    synthetic_canopy.balance_state(sg, S, np.random.Generator(np.random.PCG64(seed + 1)), 1.0e-11)
    synthetic_canopy.soilfluxes_state(sg, S, np.random.Generator(np.random.PCG64(seed + 2)))
    return sg, S


@pytest.mark.parametrize("nslab", [1, 3])
def test_window_step_matches_plain_host_step(nslab):
    sg, S = _case()
    ctx = driver.Context(abi.default_params())
    try:
        plain, win = copy_state(S), copy_state(S)
        driver.HotPath(ctx, sg, plain, abi.MEM_HOST).step()
        hp = driver.HotPath(ctx, sg, win, abi.MEM_HOST, nslab=nslab, window=True)
        hp.step()
        names = sorted({fs.name for g in driver.ROUTINES for fs in abi.FIELDS[g]})
        for k in names:
            assert np.array_equal(plain[k], win[k], equal_nan=True), k
        h2d_first, _ = hp.window_bytes()
        # second window on the same arrays: mirrors exist, only IN/INOUT first touches are uploaded
        for k in names:
            win[k][...] = S[k]
        hp.step()
        for k in names:
            assert np.array_equal(plain[k], win[k], equal_nan=True), k
        h2d, d2h = hp.window_bytes()
        want_up = 0
        seen = set()
        for g in driver.ROUTINES:
            for fs in abi.FIELDS[g]:
                n = sg.bounds.extent(fs.sub) * fs.nlev * (8 if fs.ctype == "double" else 4)
                if fs.sub == "PFT":
                    n = ctx.prm.npft_table * fs.nlev * (8 if fs.ctype == "double" else 4)
                if fs.intent in ("IN", "INOUT") and fs.name not in seen:
                    want_up += n
                seen.add(fs.name)
        filt = h2d - want_up
        assert 0 <= filt <= 4 * 8 * (sg.ncol + sg.npatch) * 2, (h2d, want_up)     # the rest are the filters
        assert h2d < h2d_first
        naive_up, naive_down = driver.staged_bytes(sg, driver.ROUTINES, preserve_out=True)
        assert h2d < 0.4 * naive_up and d2h <= naive_down
    finally:
        ctx.close()


def test_window_misuse_is_refused():
    ctx = driver.Context(abi.default_params())
    try:
        L = ctx.L
        st = abi.Status()
        assert L.ctsm_b200_host_window_end(ctx.h, C.byref(st)) == 2          # no window open
        assert L.ctsm_b200_host_window_begin(ctx.h) == 0
        assert L.ctsm_b200_host_window_begin(ctx.h) == 2                     # already open
        assert L.ctsm_b200_host_invalidate(ctx.h, None) == 2                 # not inside a window
        assert L.ctsm_b200_host_window_end(ctx.h, C.byref(st)) == 0
    finally:
        ctx.close()
