"""Independent restatements, in plain Python and written from the Fortran (NOT from oracle/oracle_canopy.c), of the surface-layer
helpers CanopyFluxes and BareGroundFluxes share: FrictionVelocity, StabilityFunc1 / 2, MoninObukIni
(src/biogeophys/FrictionVelocityMod.F90:754-1209, the patch form without landunit_index) and QSat (src/biogeophys/QSatMod.F90:61-127);
and of the CanopyFluxes patch iteration itself (src/biogeophys/CanopyFluxesMod.F90:191-1765).  Test infrastructure: they pin the C
oracle (tests/test_oracle_canopy_pin.py).  One point / one patch at a time, same libm and operation order: identical bits."""
import math
from types import SimpleNamespace

VKC = 0.4
GRAV = 9.80616
PI = 3.14159265358979323846
TKFRZ = 273.15


def stability_func1(zeta):
    """FrictionVelocityMod.F90:1120-1139"""
    chik2 = math.sqrt(1.0 - 16.0 * zeta)
    chik = math.sqrt(chik2)
    return 2.0 * math.log((1.0 + chik) * 0.5) + math.log((1.0 + chik2) * 0.5) - 2.0 * math.atan(chik) + PI * 0.5


def stability_func2(zeta):
    """FrictionVelocityMod.F90:1142-1159"""
    chik2 = math.sqrt(1.0 - 16.0 * zeta)
    return 2.0 * math.log((1.0 + chik2) * 0.5)


def monin_obuk_ini(zetamaxstable, ur, thv, dthv, zldis, z0m):
    """FrictionVelocityMod.F90:1162-1209; returns (um, obu)"""
    wc = 0.5
    if dthv >= 0.0:
        um = max(ur, 0.1)
    else:
        um = math.sqrt(ur * ur + wc * wc)
    rib = GRAV * zldis * dthv / (thv * um * um)
    if rib >= 0.0:
        zeta = rib * math.log(zldis / z0m) / (1.0 - 5.0 * min(rib, 0.19))
        zeta = min(zetamaxstable, max(zeta, 0.01))
    else:
        zeta = rib * math.log(zldis / z0m)
        zeta = max(-100.0, min(zeta, -0.01))
    return um, zldis / zeta


def _temp_profile(zldis, obu, z0):
    """the four-regime scalar profile FrictionVelocity evaluates for temp1, temp2, temp12m, temp22m"""
    zetat = 0.465
    zeta = zldis / obu
    if zeta < -zetat:
        return VKC / (math.log(-zetat * obu / z0) - stability_func2(-zetat) + stability_func2(z0 / obu)
                      + 0.8 * ((zetat) ** (-0.333) - (-zeta) ** (-0.333)))
    if zeta < 0.0:
        return VKC / (math.log(zldis / z0) - stability_func2(zeta) + stability_func2(z0 / obu))
    if zeta <= 1.0:
        return VKC / (math.log(zldis / z0) + 5.0 * zeta - 5.0 * z0 / obu)
    return VKC / (math.log(obu / z0) + 5.0 - 5.0 * z0 / obu + (5.0 * math.log(zeta) + zeta - 1.0))


def friction_velocity(hgt_u, hgt_t, hgt_q, displa, z0m, z0h, z0q, obu, it, ur, um, fm):
    """FrictionVelocityMod.F90:754-1117 for one patch; fm is the previous iteration's value (used when it > 1)"""
    zetam = 1.574
    o = SimpleNamespace()
    zldis = hgt_u - displa
    zeta = zldis / obu
    if zeta < -zetam:
        o.ustar = VKC * um / (math.log(-zetam * obu / z0m) - stability_func1(-zetam) + stability_func1(z0m / obu)
                              + 1.14 * ((-zeta) ** 0.333 - (zetam) ** 0.333))
    elif zeta < 0.0:
        o.ustar = VKC * um / (math.log(zldis / z0m) - stability_func1(zeta) + stability_func1(z0m / obu))
    elif zeta <= 1.0:
        o.ustar = VKC * um / (math.log(zldis / z0m) + 5.0 * zeta - 5.0 * z0m / obu)
    else:
        o.ustar = VKC * um / (math.log(obu / z0m) + 5.0 - 5.0 * z0m / obu + (5.0 * math.log(zeta) + zeta - 1.0))
    if zeta < 0.0:
        o.vds = 2.e-3 * o.ustar * (1.0 + (300.0 / (-obu)) ** 0.666)
    else:
        o.vds = 2.e-3 * o.ustar
    if zldis - z0m <= 10.0:
        o.u10_clm = um
    else:
        if zeta < -zetam:
            o.u10_clm = um - (o.ustar / VKC * (math.log(-zetam * obu / (10.0 + z0m)) - stability_func1(-zetam)
                                               + stability_func1((10.0 + z0m) / obu) + 1.14 * ((-zeta) ** 0.333 - (zetam) ** 0.333)))
        elif zeta < 0.0:
            o.u10_clm = um - (o.ustar / VKC * (math.log(zldis / (10.0 + z0m)) - stability_func1(zeta)
                                               + stability_func1((10.0 + z0m) / obu)))
        elif zeta <= 1.0:
            o.u10_clm = um - (o.ustar / VKC * (math.log(zldis / (10.0 + z0m)) + 5.0 * zeta - 5.0 * (10.0 + z0m) / obu))
        else:
            o.u10_clm = um - (o.ustar / VKC * (math.log(obu / (10.0 + z0m)) + 5.0 - 5.0 * (10.0 + z0m) / obu
                                               + (5.0 * math.log(zeta) + zeta - 1.0)))
    o.va = um
    o.temp1 = _temp_profile(hgt_t - displa, obu, z0h)
    if hgt_q == hgt_t and z0q == z0h:
        o.temp2 = o.temp1
    else:
        o.temp2 = _temp_profile(hgt_q - displa, obu, z0q)
    o.temp12m = _temp_profile(2.0 + z0h, obu, z0h)
    if z0q == z0h:
        o.temp22m = o.temp12m
    else:
        o.temp22m = _temp_profile(2.0 + z0q, obu, z0q)
    zldis = hgt_u - displa
    zeta = zldis / obu
    if min(zeta, 1.0) < 0.0:
        tmp1 = (1.0 - 16.0 * min(zeta, 1.0)) ** 0.25
        tmp2 = math.log((1.0 + tmp1 * tmp1) / 2.0)
        tmp3 = math.log((1.0 + tmp1) / 2.0)
        fmnew = 2.0 * tmp3 + tmp2 - 2.0 * math.atan(tmp1) + 1.5707963
    else:
        fmnew = -5.0 * min(zeta, 1.0)
    if it == 1:
        o.fm = fmnew
    else:
        o.fm = 0.5 * (fm + fmnew)
    zeta10 = min(10.0 / obu, 1.0)
    if zeta == 0.0:
        zeta10 = 0.0
    if zeta10 < 0.0:
        tmp1 = (1.0 - 16.0 * zeta10) ** 0.25
        tmp2 = math.log((1.0 + tmp1 * tmp1) / 2.0)
        tmp3 = math.log((1.0 + tmp1) / 2.0)
        fm10 = 2.0 * tmp3 + tmp2 - 2.0 * math.atan(tmp1) + 1.5707963
    else:
        fm10 = -5.0 * zeta10
    tmp4 = math.log(max(1.0, hgt_u / 10.0))
    o.u10 = ur - o.ustar / VKC * (tmp4 - o.fm + fm10)
    o.fv = o.ustar
    return o


_A = (6.11213476, 0.444007856, 0.143064234e-01, 0.264461437e-03, 0.305903558e-05, 0.196237241e-07, 0.892344772e-10,
      -0.373208410e-12, 0.209339997e-15)
_B = (0.444017302, 0.286064092e-01, 0.794683137e-03, 0.121211669e-04, 0.103354611e-06, 0.404125005e-09, -0.788037859e-12,
      -0.114596802e-13, 0.381294516e-16)
_C = (6.11123516, 0.503109514, 0.188369801e-01, 0.420547422e-03, 0.614396778e-05, 0.602780717e-07, 0.387940929e-09,
      0.149436277e-11, 0.262655803e-14)
_D = (0.503277922, 0.377289173e-01, 0.126801703e-02, 0.249468427e-04, 0.313703411e-06, 0.257180651e-08, 0.133268878e-10,
      0.394116744e-13, 0.498070196e-16)


def _horner(k, td):
    return k[0] + td * (k[1] + td * (k[2] + td * (k[3] + td * (k[4] + td * (k[5] + td * (k[6] + td * (k[7] + td * k[8])))))))


def qsat(T, p):
    """QSatMod.F90:61-127; returns (qs, es, qsdT, esdT)"""
    td = min(100.0, max(-75.0, T - TKFRZ))
    es = _horner(_A, td) if td >= 0.0 else _horner(_C, td)
    es = es * 100.0
    vp = 1.0 / (p - 0.378 * es)
    vp1 = 0.622 * vp
    qs = es * vp1
    esdT = _horner(_B, td) if td >= 0.0 else _horner(_D, td)
    esdT = esdT * 100.0
    vp2 = vp1 * vp
    qsdT = esdT * vp2 * p
    return qs, es, qsdT, esdT
