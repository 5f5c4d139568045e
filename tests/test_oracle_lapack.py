"""Pin the oracle's LAPACK restatements (oracle/oracle_lapack.c) against scipy's
LAPACK entry points — the same dgbsv / dgtsv the reference calls
(BandDiagonalMod.F90:197, SoilWaterMovementMod.F90:1287)."""
import ctypes as C

import numpy as np
import pytest
from scipy.linalg import lapack

from ctsm_b200 import abi


def _band_case(rng, n, dominant):
    kl = ku = 2
    A = np.zeros((n, n))
    for i in range(n):
        for j in range(max(0, i - kl), min(n, i + ku + 1)):
            A[i, j] = rng.normal()
        if dominant:
            A[i, i] = 1.0 + np.abs(A[i]).sum()
    ab = np.zeros((2 * kl + ku + 1, n), order="F")
    for j in range(n):
        for i in range(max(0, j - ku), min(n, j + kl + 1)):
            ab[kl + ku + i - j, j] = A[i, j]
    return A, ab


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 26, 38])
@pytest.mark.parametrize("dominant", [True, False])
def test_dgbsv_matches_scipy(oracle_lib, n, dominant):
    rng = np.random.default_rng(100 + n + (7 if dominant else 0))
    for _ in range(20):
        A, ab = _band_case(rng, n, dominant)
        b = rng.normal(size=n)
        lub, piv, x, info = lapack.dgbsv(2, 2, ab.copy(order="F"), b.copy())
        ab2 = np.asfortranarray(ab.copy())
        x2 = b.copy()
        ipiv = np.zeros(n, dtype=np.int32)
        info2 = C.c_int32(0)
        oracle_lib.oracle_dgbsv(n, 2, 2, 1, ab2.ctypes.data_as(C.POINTER(C.c_double)), 7, abi.i32p(ipiv),
                                abi.f64p(x2), n, C.byref(info2))
        assert info2.value == info == 0
        assert np.array_equal(ipiv - 1, piv)          # identical pivot sequence
        # scipy's OpenBLAS BLAS kernels may contract to FMA: agreement to a few ulp of the conditioning
        assert np.allclose(x2, x, rtol=1e-11, atol=1e-13 * np.abs(x).max())
        assert np.allclose(A @ x2, b, rtol=0, atol=1e-9 * max(1.0, np.abs(x2).max()) * np.abs(A).max())


def test_dgbsv_singular_info(oracle_lib):
    n = 6
    ab = np.zeros((7, n), order="F")
    ab[4, :] = 1.0
    ab[4, 3] = 0.0     # zero pivot, no sub-diagonals to pivot with
    b = np.ones(n)
    _, _, _, info = lapack.dgbsv(2, 2, ab.copy(order="F"), b.copy())
    ipiv = np.zeros(n, dtype=np.int32)
    info2 = C.c_int32(0)
    x = b.copy()
    oracle_lib.oracle_dgbsv(n, 2, 2, 1, ab.ctypes.data_as(C.POINTER(C.c_double)), 7, abi.i32p(ipiv), abi.f64p(x), n,
                            C.byref(info2))
    assert info2.value == info == 4
    assert np.array_equal(x, b)        # solve skipped: rhs untouched


@pytest.mark.parametrize("n", [2, 3, 5, 20])
@pytest.mark.parametrize("dominant", [True, False])
def test_dgtsv_matches_scipy(oracle_lib, n, dominant):
    rng = np.random.default_rng(200 + n + (7 if dominant else 0))
    for _ in range(50):
        dl, du = rng.normal(size=max(n - 1, 0)), rng.normal(size=max(n - 1, 0))
        d = rng.normal(size=n)
        if dominant:
            d = 1.0 + np.abs(d) + np.concatenate([[0], np.abs(dl)]) + np.concatenate([np.abs(du), [0]])
        b = rng.normal(size=n)
        _, _, _, x, info = lapack.dgtsv(dl.copy(), d.copy(), du.copy(), b.copy())
        dl2, d2, du2, x2 = (np.concatenate([dl, [0.0]]), d.copy(), np.concatenate([du, [0.0]]), b.copy())
        info2 = C.c_int32(0)
        oracle_lib.oracle_dgtsv(n, 1, abi.f64p(dl2), abi.f64p(d2), abi.f64p(du2), abi.f64p(x2), n, C.byref(info2))
        assert info2.value == info == 0
        # dgtsv has no BLAS calls: every build of the reference algorithm without FMA gives these bits
        assert np.allclose(x2, x, rtol=1e-12, atol=1e-14 * np.abs(x).max())


def test_dgtsv_zero_pivot(oracle_lib):
    dl, d, du, b = np.array([0.0, 1.0, 0.0]), np.array([0.0, 2.0, 3.0]), np.array([1.0, 1.0, 0.0]), np.ones(3)
    _, _, _, _, info = lapack.dgtsv(dl[:2].copy(), d.copy(), du[:2].copy(), b.copy())
    info2 = C.c_int32(0)
    oracle_lib.oracle_dgtsv(3, 1, abi.f64p(dl), abi.f64p(d), abi.f64p(du), abi.f64p(b), 3, C.byref(info2))
    assert info2.value == info == 1


def test_dgtsv_n1(oracle_lib):
    dl, d, du, b = np.zeros(1), np.array([4.0]), np.zeros(1), np.array([2.0])
    info = C.c_int32(0)
    oracle_lib.oracle_dgtsv(1, 1, abi.f64p(dl), abi.f64p(d), abi.f64p(du), abi.f64p(b), 1, C.byref(info))
    assert info.value == 0 and b[0] == 0.5
