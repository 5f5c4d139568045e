"""Perturbed-parameter ensemble through the PFT tables (BASELINE config 5, PFT-parameter part): member m's patches carry
itype = m*(mxpft+1) + pft and the tables hold one copy per member (ctsm_params_t.npft_table).  The ensemble run must
agree with the oracle run of the same ensemble (1e-10, identical iteration counts), and member 0 — whose table copy is
unperturbed — must reproduce the single-parameter-set run bit for bit."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi, synthetic_canopy
from tests.util import copy_state
from tests.test_gpu_canopy import run_oracle, run_gpu, compare

pytestmark = pytest.mark.gpu


def test_pft_table_ensemble(oracle_lib):
    L = abi.lib()
    nmem = 4
    sg, S = synthetic_canopy.make_full_case(400, seed=91)
    base = copy_state(S)
    member = synthetic_canopy.make_ensemble(sg, S, nmem, np.random.Generator(np.random.PCG64(92)))
    prm = abi.default_params()
    prm.npft_table = nmem * (abi.MXPFT + 1)
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) == 0
    ref, got = copy_state(S), copy_state(S)
    rc_ref, _ = run_oracle(oracle_lib, prm, sg, ref)
    rc, st = run_gpu(L, ctx, sg, got, abi.MEM_DEVICE)
    assert rc == rc_ref == 0, st.msg
    # strongly perturbed hydraulic parameters make more patches take discontinuous paths through the ci solve (bracket /
    # brent switches) on intermediate passes: the lagging outputs are held to 1e-7 here, everything else to 1e-10
    compare(sg, got, ref, S, lag_rtol=1e-7)
    L.ctsm_b200_finalize(ctx)
    # member 0 == the run with the base tables; the perturbed members differ from it
    prm1 = abi.default_params()
    ctx1 = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm1), C.byref(ctx1)) == 0
    single = copy_state(base)
    assert run_gpu(L, ctx1, sg, single, abi.MEM_DEVICE)[0] == 0
    L.ctsm_b200_finalize(ctx1)
    fe = sg.filters["exposedvegp"] - 1
    m0 = fe[member[fe] == 0]
    mk = fe[member[fe] > 0]
    for k in ("t_veg", "qflx_tran_veg", "fpsn", "num_iter", "btran"):
        assert np.array_equal(got[k][m0], single[k][m0]), k
    assert np.mean(got["qflx_tran_veg"][mk] != single["qflx_tran_veg"][mk]) > 0.5


def test_npft_table_must_be_whole_parameter_sets():
    L = abi.lib()
    prm = abi.default_params()
    prm.npft_table = abi.MXPFT + 5
    ctx = C.c_void_p()
    assert L.ctsm_b200_init(C.byref(prm), C.byref(ctx)) != 0
