"""CPU pin of the oracle's surface-layer helpers and of its CanopyFluxes (oracle/oracle_canopy.c) by tests/canopy_python.py, plain
Python written from FrictionVelocityMod.F90, QSatMod.F90 and CanopyFluxesMod.F90: identical bits."""
import ctypes as C

import numpy as np

from tests import canopy_python as cp


class FricVel(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("ustar", "temp1", "temp2", "temp12m", "temp22m", "fm", "vds", "u10_clm", "va", "u10", "fv")]


def test_friction_velocity_monin_obukhov_qsat_match_python(oracle_lib):
    OL = oracle_lib
    d = C.c_double
    OL.oracle_friction_velocity_point.argtypes = [d] * 8 + [C.c_int, d, d, C.POINTER(FricVel)]
    OL.oracle_friction_velocity_point.restype = None
    OL.oracle_moninobukini.argtypes = [d] * 6 + [C.POINTER(d)] * 2
    OL.oracle_moninobukini.restype = None
    OL.oracle_qsat.argtypes = [d, d] + [C.POINTER(d)] * 3
    OL.oracle_qsat.restype = None
    rng = np.random.Generator(np.random.PCG64(1301))
    regimes = {"very_unstable": 0, "unstable": 0, "stable": 0, "very_stable": 0, "tall": 0, "same_hgt": 0}
    for n in range(6000):
        hgt_u = float(rng.uniform(2.5, 60.0))
        same = n % 3 == 0
        hgt_t = hgt_u if same else float(rng.uniform(2.5, 60.0))
        hgt_q = hgt_t if n % 2 == 0 else float(rng.uniform(2.5, 60.0))
        displa = float(rng.uniform(0.0, 0.6 * min(hgt_u, hgt_t, hgt_q)))
        z0m = float(10 ** rng.uniform(-4, 0.3))
        z0h = z0m if n % 5 == 0 else float(z0m * 10 ** rng.uniform(-3, 0))
        z0q = z0h if n % 2 == 0 else float(z0m * 10 ** rng.uniform(-3, 0))
        zl = hgt_u - displa
        zeta = float(rng.choice([-1.0, 1.0]) * 10 ** rng.uniform(-2.5, 1.3))
        if n % 97 == 0:
            zeta = float(rng.choice([-1.574, -0.465, 1.0]))       # the regime boundaries themselves
        obu = zl / zeta
        it = int(rng.integers(1, 5))
        ur, um = float(rng.uniform(0.1, 25.0)), float(rng.uniform(0.1, 25.0))
        fm0 = float(rng.uniform(-5, 3))
        o = FricVel(fm=fm0)
        OL.oracle_friction_velocity_point(hgt_u, hgt_t, hgt_q, displa, z0m, z0h, z0q, obu, it, ur, um, C.byref(o))
        r = cp.friction_velocity(hgt_u, hgt_t, hgt_q, displa, z0m, z0h, z0q, obu, it, ur, um, fm0)
        for k, _ in FricVel._fields_:
            assert getattr(o, k) == getattr(r, k), (n, k, getattr(o, k), getattr(r, k))
        z = zl / obu
        regimes["very_unstable"] += z < -1.574
        regimes["unstable"] += -1.574 <= z < 0
        regimes["stable"] += 0 <= z <= 1
        regimes["very_stable"] += z > 1
        regimes["tall"] += zl - z0m > 10.0
        regimes["same_hgt"] += (hgt_q == hgt_t and z0q == z0h)
        # MoninObukIni
        thv, dthv = float(rng.uniform(240, 320)), float(rng.uniform(-12, 12) if n % 50 else 0.0)
        zms = float(rng.choice([0.5, 2.0]))
        um_o, obu_o = d(), d()
        OL.oracle_moninobukini(zms, ur, thv, dthv, zl, z0m, C.byref(um_o), C.byref(obu_o))
        assert (um_o.value, obu_o.value) == cp.monin_obuk_ini(zms, ur, thv, dthv, zl, z0m)
        # QSat
        T, p = float(rng.uniform(180, 390)), float(rng.uniform(5.0e4, 1.05e5))
        if n % 101 == 0:
            T = 273.15
        qs, es, qsdT = d(), d(), d()
        OL.oracle_qsat(T, p, C.byref(qs), C.byref(es), C.byref(qsdT))
        assert (qs.value, es.value, qsdT.value) == cp.qsat(T, p)[:3]
    assert min(regimes.values()) > 300, regimes
