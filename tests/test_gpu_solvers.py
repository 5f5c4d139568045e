"""GPU parity: batched solvers through the C ABI vs the CPU oracle (bit-level tolerances)."""
import ctypes as C

import numpy as np
import pytest

from ctsm_b200 import abi
from tests.util import relerr

pytestmark = pytest.mark.gpu


def _bounds(begc, endc):
    b = abi.Bounds()
    b.begc, b.endc = begc, endc
    b.begg, b.endg, b.begl, b.endl, b.begp, b.endp = 1, 1, 1, 1, 1, 1
    b.begCohort, b.endCohort, b.level, b.clump_index = 1, 0, 1, -1
    return b


def _tridiag_case(nc, seed, begc=1, frac=0.9):
    """lbj=-11..ubj=25, jtop = snl+1, 90% ascending filter, diagonally dominant."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lbj, ubj = -11, 25
    nl = ubj - lbj + 1
    snl = rng.integers(-12, 1, size=nc)
    jtop = (snl + 1).astype(np.int32)
    filt = (np.nonzero(rng.random(nc) < frac)[0] + begc).astype(np.int32)
    a = -rng.uniform(0, 2, (nl, nc)); c = -rng.uniform(0, 2, (nl, nc))
    b = 1 + np.abs(a) + np.abs(c) + rng.uniform(0, 1, (nl, nc))
    r = rng.normal(280, 10, (nl, nc))
    u = np.full((nl, nc), -777.0)
    return lbj, ubj, jtop, filt, a, b, c, r, u


@pytest.mark.parametrize("nc,begc,mem", [(4096, 1, abi.MEM_HOST), (4096, 1, abi.MEM_DEVICE), (1000, 37, abi.MEM_DEVICE),
                                         (1, 1, abi.MEM_HOST), (262144, 1, abi.MEM_DEVICE)])
def test_tridiagonal_matches_oracle(gpu_ctx, oracle_lib, nc, begc, mem):
    import torch
    L, ctx, _ = gpu_ctx
    lbj, ubj, jtop, filt, a, b, c, r, u = _tridiag_case(nc, 20260101 + 1, begc)
    bd = _bounds(begc, begc + nc - 1)
    u_ref = u.copy()
    oracle_lib.oracle_tridiagonal(C.byref(bd), lbj, ubj, abi.i32p(jtop), len(filt), abi.i32p(filt), abi.f64p(a),
                                  abi.f64p(b), abi.f64p(c), abi.f64p(r), abi.f64p(u_ref))
    if mem == abi.MEM_HOST:
        rc = L.ctsm_b200_tridiagonal(ctx, C.byref(bd), lbj, ubj, abi.i32p(jtop), len(filt), abi.i32p(filt),
                                     abi.f64p(a), abi.f64p(b), abi.f64p(c), abi.f64p(r), abi.f64p(u), mem)
        assert rc == 0
        got = u
    else:
        d = {k: torch.from_numpy(v).cuda() for k, v in dict(jtop=jtop, filt=filt, a=a, b=b, c=c, r=r, u=u).items()}
        rc = L.ctsm_b200_tridiagonal(ctx, C.byref(bd), lbj, ubj, abi.i32p(d["jtop"]), len(filt), abi.i32p(d["filt"]),
                                     abi.f64p(d["a"]), abi.f64p(d["b"]), abi.f64p(d["c"]), abi.f64p(d["r"]),
                                     abi.f64p(d["u"]), mem)
        assert rc == 0
        st = abi.Status()
        assert L.ctsm_b200_sync(ctx, C.byref(st)) == 0
        got = d["u"].cpu().numpy()
    # same operations in the same order, no FMA on either side: bit-exact
    assert np.array_equal(got, u_ref)
    # untouched entries (outside filter / above jtop) keep their previous value
    assert (got == -777.0).sum() == (u_ref == -777.0).sum() > 0 or nc == 1
    # analytic check on one column against a dense solve
    ci = filt[0] - begc
    j0 = jtop[ci] - lbj
    n = (ubj - lbj + 1) - j0
    A = np.diag(b[j0:, ci]) + np.diag(a[j0 + 1:, ci], -1) + np.diag(c[j0:-1, ci], 1)
    assert relerr(got[j0:, ci], np.linalg.solve(A, r[j0:, ci])) <= 1e-12


def _band_case(nc, seed, pivoting):
    rng = np.random.Generator(np.random.PCG64(seed))
    lbj, ubj = -12, 25
    nl = ubj - lbj + 1
    jtop = rng.integers(-12, 1, size=nc).astype(np.int32)
    jbot = np.full(nc, 25, dtype=np.int32)
    b = rng.normal(size=(nl, 5, nc))
    if not pivoting:
        b[:, 2, :] = 1.0 + np.abs(b).sum(axis=1)
    r = rng.normal(280, 10, (nl, nc))
    u = np.full((nl, nc), -777.0)
    filt = (np.nonzero(rng.random(nc) < 0.9)[0] + 1).astype(np.int32)
    return lbj, ubj, jtop, jbot, filt, b, r, u


@pytest.mark.parametrize("pivoting", [False, True])
@pytest.mark.parametrize("mem", [abi.MEM_HOST, abi.MEM_DEVICE])
def test_banddiagonal_matches_oracle(gpu_ctx, oracle_lib, pivoting, mem):
    import torch
    L, ctx, _ = gpu_ctx
    nc = 5000
    lbj, ubj, jtop, jbot, filt, b, r, u = _band_case(nc, 99 + pivoting, pivoting)
    bd = _bounds(1, nc)
    u_ref = u.copy()
    st = abi.Status()
    assert oracle_lib.oracle_banddiagonal(C.byref(bd), lbj, ubj, abi.i32p(jtop), abi.i32p(jbot), len(filt),
                                          abi.i32p(filt), 5, abi.f64p(b), abi.f64p(r), abi.f64p(u_ref), C.byref(st)) == 0
    if mem == abi.MEM_HOST:
        rc = L.ctsm_b200_banddiagonal(ctx, C.byref(bd), lbj, ubj, abi.i32p(jtop), abi.i32p(jbot), len(filt),
                                      abi.i32p(filt), 5, abi.f64p(b), abi.f64p(r), abi.f64p(u), mem, C.byref(st))
        got = u
    else:
        d = {k: torch.from_numpy(v).cuda() for k, v in dict(jtop=jtop, jbot=jbot, filt=filt, b=b, r=r, u=u).items()}
        rc = L.ctsm_b200_banddiagonal(ctx, C.byref(bd), lbj, ubj, abi.i32p(d["jtop"]), abi.i32p(d["jbot"]), len(filt),
                                      abi.i32p(d["filt"]), 5, abi.f64p(d["b"]), abi.f64p(d["r"]), abi.f64p(d["u"]), mem,
                                      C.byref(st))
        assert rc == 0
        rc = L.ctsm_b200_sync(ctx, C.byref(st))
        got = d["u"].cpu().numpy()
    assert rc == 0, st.msg
    assert np.array_equal(got, u_ref)       # identical pivot choices and operation order => identical bits


def test_banddiagonal_singular_reports_first_column(gpu_ctx, oracle_lib):
    L, ctx, _ = gpu_ctx
    nc = 300
    lbj, ubj, jtop, jbot, filt, b, r, u = _band_case(nc, 5, False)
    filt = np.arange(1, nc + 1, dtype=np.int32)
    for bad in (120, 77):                   # two singular columns: the lower index must be reported
        b[:, :, bad - 1] = 0.0
    bd = _bounds(1, nc)
    st, st2 = abi.Status(), abi.Status()
    u_ref = u.copy()
    rc_ref = oracle_lib.oracle_banddiagonal(C.byref(bd), lbj, ubj, abi.i32p(jtop), abi.i32p(jbot), nc, abi.i32p(filt), 5,
                                            abi.f64p(b), abi.f64p(r), abi.f64p(u_ref), C.byref(st2))
    rc = L.ctsm_b200_banddiagonal(ctx, C.byref(bd), lbj, ubj, abi.i32p(jtop), abi.i32p(jbot), nc, abi.i32p(filt), 5,
                                  abi.f64p(b), abi.f64p(r), abi.f64p(u), abi.MEM_HOST, C.byref(st))
    assert rc == rc_ref == 10
    assert (st.subgrid_index, st.info, st.subgrid_level) == (st2.subgrid_index, st2.info, 3) == (77, 1, 3)
    assert st.msg == st2.msg == b"BandDiagonal ERROR: dgbsv returned error code"


@pytest.mark.parametrize("dominant", [True, False])
def test_dgtsv_batch_matches_oracle(gpu_ctx, oracle_lib, dominant):
    L, ctx, _ = gpu_ctx
    rng = np.random.Generator(np.random.PCG64(7 + dominant))
    nc, nlev = 4000, 20
    nlayers = rng.integers(2, nlev + 1, size=nc).astype(np.int32)
    filt = (np.nonzero(rng.random(nc) < 0.8)[0] + 1).astype(np.int32)
    amx, cmx = rng.normal(size=(nlev, nc)), rng.normal(size=(nlev, nc))
    bmx = rng.normal(size=(nlev, nc))
    if dominant:
        bmx = -1.0 - np.abs(amx) - np.abs(cmx) - rng.random((nlev, nc))
    rmx = rng.normal(size=(nlev, nc))
    x, x_ref = np.full((nlev, nc), -777.0), np.full((nlev, nc), -777.0)
    bd = _bounds(1, nc)
    st = abi.Status()
    assert oracle_lib.oracle_dgtsv_batch(C.byref(bd), nlev, abi.i32p(nlayers), len(filt), abi.i32p(filt), abi.f64p(amx),
                                         abi.f64p(bmx), abi.f64p(cmx), abi.f64p(rmx), abi.f64p(x_ref), C.byref(st)) == 0
    rc = L.ctsm_b200_dgtsv_batch(ctx, C.byref(bd), nlev, abi.i32p(nlayers), len(filt), abi.i32p(filt), abi.f64p(amx),
                                 abi.f64p(bmx), abi.f64p(cmx), abi.f64p(rmx), abi.f64p(x), abi.MEM_HOST, C.byref(st))
    assert rc == 0, st.msg
    assert np.array_equal(x, x_ref)
