#!/usr/bin/env python
"""Writes tests/golden/reference_unit_tests.json: the known answers held by the reference's own
pFUnit tests for the hot path, transcribed from the .pf files (paths relative to the CTSM checkout):

  src/biogeophys/test/Photosynthesis_test/test_Photosynthesis.pf:45,57-71   plc, d1plc (tol 1e-13)
  src/biogeophys/PhotosynthesisMod.F90:929-934                              the parameters those tests set
  src/utils/test/quadratic_test/test_quadratic.pf                           quadratic
  src/utils/test/numerics_test/test_truncate_small_values.pf                truncate_small_values
  src/biogeophys/test/Balance_test/test_Balance.pf:39-105                   BalanceCheckInit skip steps

The reference itself cannot be executed here (no Fortran compiler), so these are
transcriptions of the expected values in its tests, not outputs of a run.
"""
import json
import os

gold = {
    "plc": {"nlevgrnd": 25, "psi50": -150000.0, "ck": 3.95, "plc": 0.999415208562283,
            "d1plc": 9.237002539040872e-08, "tol": 1e-13},
    "quadratic": [
        # zero_root: a = 1000, 5.12345678, -31.415927465859; b = c = 0 -> r1 = 0, r2 = 1e36 (tol 1e-13... tol of the .pf)
        {"a": 1000.0, "b": 0.0, "c": 0.0, "r1": 0.0, "r2": 1.0e36, "tol": 1e-13},
        {"a": 5.12345678, "b": 0.0, "c": 0.0, "r1": 0.0, "r2": 1.0e36, "tol": 1e-13},
        {"a": -31.415927465859, "b": 0.0, "c": 0.0, "r1": 0.0, "r2": 1.0e36, "tol": 1e-13},
        # simple_roots
        {"a": 1.0, "b": 3.0, "c": 2.0, "r1": -2.0, "r2": -1.0, "tol": 1e-13},
        {"a": 1.0, "b": 0.0, "c": -4.0, "r1": -2.0, "r2": 2.0, "tol": 1e-13},
        # check_errorcondsazero / imaginary / barelyimaginary (4 + 100*epsilon)
        {"a": 0.0, "b": 0.0, "c": 0.0, "aborts": True},
        {"a": 1.0, "b": 2.0, "c": 5.0, "aborts": True},
        {"a": 1.0, "b": 4.0, "c": 4.0 + 100.0 * 2.220446049250313e-16, "aborts": True},
    ],
    # HumanStress_test/test_humanstress.pf:26-29 (Wet_BulbS, tolerance 1e-8)
    "wet_bulbs": [{"tc": 0.0, "rh": 0.0, "wbt": -3.6531108341574, "tol": 1e-8},
                  {"tc": 0.0, "rh": 100.0, "wbt": -0.13165370616986, "tol": 1e-8}],
    "balance_skip_steps": [[1800, 3], [7200, 3], [300, 13], [36, 101]],
    "truncate_small_values": [
        {"name": "tsv_truncates_correct_points", "eps": 1e-13, "filter": [1, 2, 3],
         "baseline": [1.0, 1.0, 1.0], "data": [0.5, 1e-16, -1.0], "expect": [0.5, 0.0, -1.0]},
        {"name": "tsv_custom_tolerance_truncates_correct_points", "eps": 1e-12, "filter": [1, 2, 3],
         "baseline": [1.0, 1.0, 1.0], "data": [5e-12, 5e-13, 5e-12], "expect": [5e-12, 0.0, 5e-12]},
        {"name": "tsv_truncates_large_magnitude", "eps": 1e-13, "filter": [1],
         "baseline": [1e30], "data": [1e10], "expect": [0.0]},
        {"name": "tsv_does_not_truncate_small_magnitude", "eps": 1e-13, "filter": [1],
         "baseline": [1e-30], "data": [1e-31], "expect": [1e-31]},
    ],
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_unit_tests.json"), "w") as fh:
    json.dump(gold, fh, indent=1)
