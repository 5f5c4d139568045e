"""Pins the CPU oracle to every known answer the reference's own unit tests hold for the hot path
(SURVEY.md section 8c):
  plc / d1plc      src/biogeophys/test/Photosynthesis_test/test_Photosynthesis.pf:45,57-71
                   (params from PhotosynthesisMod.F90:929-934 setParamsForTesting)
  quadratic        src/utils/test/quadratic_test/test_quadratic.pf
  truncate_small_values   src/utils/test/numerics_test/
  Wet_BulbS        src/biogeophys/test/HumanStress_test/test_humanstress.pf:26-29 (fast human-stress indices)
  BalanceCheckInit skip steps   src/biogeophys/test/Balance_test/test_Balance.pf:39-105
  filter order     src/main/test/filter_test/test_filter_col.pf (stable ascending order)
The vectors are also stored in tests/golden/reference_unit_tests.json (written by
tests/golden/make_golden.py from the values quoted in those .pf files).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from ctsm_b200 import abi

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_unit_tests.json")))


def test_plc_d1plc_known_answers(oracle_lib):
    g = GOLD["plc"]
    x = -1000.0 * g["nlevgrnd"]
    assert abs(oracle_lib.oracle_plc(x, g["psi50"], g["ck"]) - g["plc"]) <= g["tol"]
    assert abs(oracle_lib.oracle_d1plc(x, g["psi50"], g["ck"]) - g["d1plc"]) <= g["tol"]


def test_plc_clamps_below_half_percent(oracle_lib):
    # PhotosynthesisMod.F90:5187
    assert oracle_lib.oracle_plc(-3.0 * 150000.0, -150000.0, 3.95) == 0.0
    assert oracle_lib.oracle_plc(-1.0, -150000.0, 3.95) > 0.999999


@pytest.mark.parametrize("case", GOLD["quadratic"])
def test_quadratic_known_answers(oracle_lib, case):
    r1, r2 = C.c_double(), C.c_double()
    rc = oracle_lib.oracle_quadratic(case["a"], case["b"], case["c"], C.byref(r1), C.byref(r2))
    if case.get("aborts"):
        assert rc == 15          # CTSM_ERR_QUADRATIC where the reference calls endrun
        return
    assert rc == 0
    tol = case.get("tol", 0.0)
    assert abs(r1.value - case["r1"]) <= tol and abs(r2.value - case["r2"]) <= tol


def test_quadratic_near_zero_discriminant(oracle_lib):
    # test_quadratic.pf: (1, 4, 4 + {0, eps/2, 2 eps}) -> r1 ~ r2 ~ -2 within 1e-6
    eps = np.finfo(np.float64).eps
    for d in (0.0, 0.5 * eps, 2.0 * eps):
        r1, r2 = C.c_double(), C.c_double()
        assert oracle_lib.oracle_quadratic(1.0, 4.0, 4.0 + d, C.byref(r1), C.byref(r2)) == 0
        assert abs(r1.value + 2.0) < 1e-6 and abs(r2.value + 2.0) < 1e-6


def test_wet_bulbs_known_answers(oracle_lib):
    oracle_lib.oracle_wet_bulbs.argtypes = [C.c_double, C.c_double]
    oracle_lib.oracle_wet_bulbs.restype = C.c_double
    for g in GOLD["wet_bulbs"]:
        assert abs(oracle_lib.oracle_wet_bulbs(g["tc"], g["rh"]) - g["wbt"]) <= g["tol"]
    tc = 100.0                                                     # the .pf's sweep: never NaN from 100 C down to -50 C at rh = 100
    while tc > -50.0:
        assert np.isfinite(oracle_lib.oracle_wet_bulbs(tc, 100.0))
        tc -= 0.1


def test_balancecheck_skip_steps(oracle_lib):
    oracle_lib.oracle_balancecheck_skip_steps.argtypes = [C.c_double]
    oracle_lib.oracle_balancecheck_skip_steps.restype = C.c_int
    for dtime, want in GOLD["balance_skip_steps"]:
        assert oracle_lib.oracle_balancecheck_skip_steps(float(dtime)) == want


def test_truncate_small_values(oracle_lib):
    oracle_lib.oracle_truncate_small_values.argtypes = [C.c_int, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_double),
                                                        C.POINTER(C.c_double), C.c_double]
    oracle_lib.oracle_truncate_small_values.restype = None
    for case in GOLD["truncate_small_values"]:
        base = np.array(case["baseline"], dtype=np.float64)
        data = np.array(case["data"], dtype=np.float64)
        filt = np.array(case["filter"], dtype=np.int32)
        oracle_lib.oracle_truncate_small_values(len(filt), abi.i32p(filt), 1, abi.f64p(base), abi.f64p(data), case["eps"])
        assert np.array_equal(data, np.array(case["expect"], dtype=np.float64)), case["name"]


def test_exposedveg_filter_is_stable_ascending(oracle_lib):
    b = abi.Bounds()
    b.begp, b.endp = 3, 12
    filt = np.array([3, 4, 6, 7, 9, 12], dtype=np.int32)
    fv = np.array([1, 0, 9, 1, 1, 9, 0, 9, 9, 1], dtype=np.int32)       # indexed begp..endp
    ey, en = np.zeros(6, dtype=np.int32), np.zeros(6, dtype=np.int32)
    ny, nn = C.c_int32(), C.c_int32()
    oracle_lib.oracle_set_exposedvegp_filter(C.byref(b), 6, abi.i32p(filt), abi.i32p(fv), abi.i32p(ey), C.byref(ny),
                                             abi.i32p(en), C.byref(nn))
    assert list(ey[:ny.value]) == [3, 6, 7, 12] and list(en[:nn.value]) == [4, 9]
