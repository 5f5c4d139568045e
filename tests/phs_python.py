"""Independent restatement of PhotosynthesisHydraulicStress and everything it calls (src/biogeophys/PhotosynthesisMod.F90:2704-5228:
PhotosynthesisHydraulicStress, hybrid_PHS, brent_PHS, ci_func_PHS, calcstress, spacA, spacF, getvegwp, getqflx, plc, d1plc; quadratic
src/utils/quadraticMod.F90) in plain Python, written from the Fortran, NOT from oracle/oracle_phs.c.  Test infrastructure: it pins
the C oracle (tests/test_oracle_phs.py).  One patch at a time; `P` is a plain namespace of the patch's inputs (scalars and
Fortran-indexed lists), vegwp vectors are dicts keyed 1..4 (sun, sha, xyl, root) as in the reference."""
import math
from types import SimpleNamespace

SUN, SHA, XYL, ROOT = 1, 2, 3, 4
NLEVSOI = 20
TOL_LAI = 0.001
RGAS = 6.02214e26 * 1.38065e-23
TFRZ = 273.15
SPVAL = 1.0e36


def plc(x, P, level):
    """:5167-5195 (vegetation_weibull)"""
    v = 2.0 ** (-(x / P.psi50[level]) ** P.ck[level])
    if v < 0.005:
        v = 0.0
    return v


def d1plc(x, P, level):
    """:5199-5228"""
    return -P.ck[level] * math.log(2.0) * (2.0 ** (-(x / P.psi50[level]) ** P.ck[level])) * ((x / P.psi50[level]) ** P.ck[level]) / x


def getqflx(P, gb_mol, gs_mol_sun, gs_mol_sha, qflx_sun, qflx_sha, qsatl, qaf, havegs):
    """:5080-5164; returns (gs_mol_sun, gs_mol_sha, qflx_sun, qflx_sha)"""
    cf = P.forc_pbot / (RGAS * 1.e-3 * P.tgcm) * 1.e6
    wtl = (P.elai + P.esai) * gb_mol
    efpot = P.forc_rho * wtl * (qsatl - qaf)
    if havegs:
        if efpot > 0.0 and P.elai > 0.0:
            if gs_mol_sun > 0.0:
                rppdry_sun = P.fdry / gb_mol * (P.laisun / (1.0 / gb_mol + 1.0 / gs_mol_sun)) / P.elai
                qflx_sun = efpot * rppdry_sun / cf
            else:
                qflx_sun = 0.0
            if gs_mol_sha > 0.0:
                rppdry_sha = P.fdry / gb_mol * (P.laisha / (1.0 / gb_mol + 1.0 / gs_mol_sha)) / P.elai
                qflx_sha = efpot * rppdry_sha / cf
            else:
                qflx_sha = 0.0
        else:
            qflx_sun = 0.0
            qflx_sha = 0.0
    else:
        if qflx_sun > 0.0:
            gs_mol_sun = gb_mol * qflx_sun * cf * P.elai / (efpot * P.fdry * P.laisun - qflx_sun * cf * P.elai)
        else:
            gs_mol_sun = 0.0
        if qflx_sha > 0.0:
            gs_mol_sha = gb_mol * qflx_sha * cf * P.elai / (efpot * P.fdry * P.laisha - qflx_sha * cf * P.elai)
        else:
            gs_mol_sha = 0.0
    return gs_mol_sun, gs_mol_sha, qflx_sun, qflx_sha


def _fsum(values):
    """Fortran SUM of an array expression as gfortran inlines it: a running sum in index order"""
    s = 0.0
    for v in values:
        s = s + v
    return s


def getvegwp(P, x, gb_mol, gs_mol_sun, gs_mol_sha, qsatl, qaf):
    """:4979-5077; x is updated in place; returns soilflux"""
    grav1 = 1000.0 * P.htop
    grav2 = {j: 1000.0 * P.z[j] for j in range(1, NLEVSOI + 1)}
    _, _, qflx_sun, qflx_sha = getqflx(P, gb_mol, gs_mol_sun, gs_mol_sha, 0.0, 0.0, qsatl, qaf, True)
    ksum = _fsum(P.k[j] for j in range(1, NLEVSOI + 1))
    if abs(ksum) == 0.0:
        x[ROOT] = _fsum(P.smp[j] - grav2[j] for j in range(1, NLEVSOI + 1)) / NLEVSOI
    else:
        x[ROOT] = (_fsum(P.k[j] * (P.smp[j] - grav2[j]) for j in range(1, NLEVSOI + 1)) - qflx_sun - qflx_sha) / ksum
    fr = plc(x[ROOT], P, ROOT)
    if P.tsai > 0.0 and fr > 0.0:
        x[XYL] = x[ROOT] - grav1 - (qflx_sun + qflx_sha) / (fr * P.kmax[ROOT] / P.htop * P.tsai)
    else:
        x[XYL] = x[ROOT] - grav1
    fx = plc(x[XYL], P, XYL)
    if P.laisha > 0.0 and fx > 0.0:
        x[SHA] = x[XYL] - (qflx_sha / (fx * P.kmax[XYL] * P.laisha))
    else:
        x[SHA] = x[XYL]
    if P.laisun > 0.0 and fx > 0.0:
        x[SUN] = x[XYL] - (qflx_sun / (fx * P.kmax[XYL] * P.laisun))
    else:
        x[SUN] = x[XYL]
    soilflux = 0.0
    for j in range(1, NLEVSOI + 1):
        soilflux = soilflux + P.k[j] * (P.smp[j] - x[ROOT] - grav2[j])
    return soilflux


def spacF(P, x, qflx_sun, qflx_sha):
    """:4898-4976; returns f keyed 1..4"""
    grav1 = P.htop * 1000.0
    grav2 = {j: P.z[j] * 1000.0 for j in range(1, NLEVSOI + 1)}
    fsto1 = plc(x[SUN], P, SUN)
    fsto2 = plc(x[SHA], P, SHA)
    fx = plc(x[XYL], P, XYL)
    fr = plc(x[ROOT], P, ROOT)
    f = {}
    f[SUN] = qflx_sun * fsto1 - P.laisun * P.kmax[SUN] * fx * (x[XYL] - x[SUN])
    f[SHA] = qflx_sha * fsto2 - P.laisha * P.kmax[SHA] * fx * (x[XYL] - x[SHA])
    f[XYL] = (P.laisun * P.kmax[SUN] * fx * (x[XYL] - x[SUN]) + P.laisha * P.kmax[SHA] * fx * (x[XYL] - x[SHA])
              - P.tsai * P.kmax[XYL] / P.htop * fr * (x[ROOT] - x[XYL] - grav1))
    f[ROOT] = (P.tsai * P.kmax[XYL] / P.htop * fr * (x[ROOT] - x[XYL] - grav1)
               + _fsum(P.k[j] * (x[ROOT] + grav2[j]) for j in range(1, NLEVSOI + 1))
               - _fsum(P.k[j] * P.smp[j] for j in range(1, NLEVSOI + 1)))
    if P.laisha < TOL_LAI:
        f[SUN], f[SHA] = f[SHA], f[SUN]
    return f


def spacA(P, x, qflx_sun, qflx_sha):
    """:4715-4893; returns (invA, flag) with invA[i][k], i, k = 1..4"""
    A = {i: {k: 0.0 for k in range(1, 5)} for i in range(1, 5)}
    inv = {i: {k: 0.0 for k in range(1, 5)} for i in range(1, 5)}
    grav1 = P.htop * 1000.0
    fx = plc(x[XYL], P, XYL)
    fr = plc(x[ROOT], P, ROOT)
    dfsto1 = d1plc(x[SUN], P, SUN)
    dfsto2 = d1plc(x[SHA], P, SHA)
    dfx = d1plc(x[XYL], P, XYL)
    dfr = d1plc(x[ROOT], P, ROOT)
    ls, lh, kx = P.laisun, P.laisha, P.kmax
    A[1][1] = -ls * kx[SUN] * fx - qflx_sun * dfsto1
    A[1][3] = ls * kx[SUN] * dfx * (x[XYL] - x[SUN]) + ls * kx[SUN] * fx
    A[2][2] = -lh * kx[SHA] * fx - qflx_sha * dfsto2
    A[2][3] = lh * kx[SHA] * dfx * (x[XYL] - x[SHA]) + lh * kx[SHA] * fx
    A[3][1] = ls * kx[SUN] * fx
    A[3][2] = lh * kx[SHA] * fx
    A[3][3] = (-ls * kx[SUN] * dfx * (x[XYL] - x[SUN]) - ls * kx[SUN] * fx - lh * kx[SHA] * dfx * (x[XYL] - x[SHA]) - lh * kx[SHA] * fx
               - P.tsai * kx[XYL] / P.htop * fr)
    A[3][4] = P.tsai * kx[XYL] / P.htop * dfr * (x[ROOT] - x[XYL] - grav1) + P.tsai * kx[XYL] / P.htop * fr
    A[4][3] = P.tsai * kx[XYL] / P.htop * fr
    A[4][4] = (-P.tsai * kx[XYL] / P.htop * fr - P.tsai * kx[XYL] / P.htop * dfr * (x[ROOT] - x[XYL] - grav1)
               - _fsum(P.k[j] for j in range(1, NLEVSOI + 1)))
    invfactor = 1.0
    for i in range(1, 5):
        for k in range(1, 5):
            A[i][k] = invfactor * A[i][k]
    if ls > TOL_LAI and lh > TOL_LAI:
        determ = (A[4][4] * A[2][2] * A[3][3] * A[1][1] - A[4][4] * A[2][2] * A[3][1] * A[1][3]
                  - A[4][4] * A[3][2] * A[2][3] * A[1][1] - A[4][3] * A[1][1] * A[2][2] * A[3][4])
        if abs(determ) <= 1.e-50:
            return inv, True
        L = 1.0 / determ
        inv[1][1] = L * A[4][4] * A[2][2] * A[3][3] - L * A[4][4] * A[3][2] * A[2][3] - L * A[4][3] * A[2][2] * A[3][4]
        inv[2][1] = L * A[2][3] * A[4][4] * A[3][1]
        inv[3][1] = -L * A[4][4] * A[2][2] * A[3][1]
        inv[4][1] = L * A[4][3] * A[2][2] * A[3][1]
        inv[1][2] = L * A[1][3] * A[4][4] * A[3][2]
        inv[2][2] = L * A[4][4] * A[3][3] * A[1][1] - L * A[4][4] * A[3][1] * A[1][3] - L * A[4][3] * A[1][1] * A[3][4]
        inv[3][2] = -L * A[1][1] * A[4][4] * A[3][2]
        inv[4][2] = L * A[4][3] * A[1][1] * A[3][2]
        inv[1][3] = -L * A[1][3] * A[2][2] * A[4][4]
        inv[2][3] = -L * A[2][3] * A[1][1] * A[4][4]
        inv[3][3] = L * A[2][2] * A[1][1] * A[4][4]
        inv[4][3] = -L * A[4][3] * A[1][1] * A[2][2]
        inv[1][4] = L * A[1][3] * A[3][4] * A[2][2]
        inv[2][4] = L * A[2][3] * A[3][4] * A[1][1]
        inv[3][4] = -L * A[3][4] * A[1][1] * A[2][2]
        inv[4][4] = L * A[2][2] * A[3][3] * A[1][1] - L * A[2][2] * A[3][1] * A[1][3] - L * A[3][2] * A[2][3] * A[1][1]
        for i in range(1, 5):
            for k in range(1, 5):
                inv[i][k] = invfactor * inv[i][k]
    else:
        if lh <= TOL_LAI:
            A[2][2] = A[1][1]
            A[3][2] = A[3][1]
            A[2][3] = A[1][3]
        determ = A[2][2] * A[3][3] * A[4][4] - A[3][4] * A[2][2] * A[4][3] - A[2][3] * A[3][2] * A[4][4]
        if abs(determ) <= 1.e-50:
            return inv, True
        inv[2][2] = A[3][3] * A[4][4] - A[3][4] * A[4][3]
        inv[2][3] = -A[2][3] * A[4][4]
        inv[2][4] = A[3][4] * A[2][3]
        inv[3][2] = -A[3][2] * A[4][4]
        inv[3][3] = A[2][2] * A[4][4]
        inv[3][4] = -A[3][4] * A[2][2]
        inv[4][2] = A[3][2] * A[4][3]
        inv[4][3] = -A[2][2] * A[4][3]
        inv[4][4] = A[2][2] * A[3][3] - A[2][3] * A[3][2]
        r = 1.0 / determ
        for i in range(1, 5):
            for k in range(1, 5):
                inv[i][k] = r * inv[i][k]
    return inv, False


def calcstress(P, x, gb_mol, gs_mol_sun, gs_mol_sha, qsatl, qaf):
    """:4490-4710; x (dict 1..4) is updated in place; returns SimpleNamespace(bsun, bsha, night, tran, iters).
    tran is qflx_tran_veg when the routine sets it (night), else None; vegwp_pd is x at night before local noon, else spval."""
    itmax, tolf, toldx = 50, 1.e-6, 1.e-9
    if x[SUN] > 0.0:
        night = True
        x[SUN] = x[SHA]
    else:
        night = False
    gs0sun, gs0sha = gs_mol_sun, gs_mol_sha
    gs0sun, gs0sha, qflx_sun, qflx_sha = getqflx(P, gb_mol, gs0sun, gs0sha, 0.0, 0.0, qsatl, qaf, True)
    iters = 0
    if (P.laisun > TOL_LAI or P.laisha > TOL_LAI) and (qflx_sun > 0.0 or qflx_sha > 0.0):
        it = 0
        while True:
            it += 1
            f = spacF(P, x, qflx_sun, qflx_sha)
            if math.sqrt(_fsum(f[i] * f[i] for i in range(1, 5))) < tolf * (qflx_sun + qflx_sha):
                flag = False
                break
            if it > itmax:
                flag = False
                break
            A, flag = spacA(P, x, qflx_sun, qflx_sha)
            if flag:
                break
            dx = {}
            if P.laisun > TOL_LAI and P.laisha > TOL_LAI:
                for i in range(1, 5):
                    dx[i] = _fsum(A[i][k] * f[k] for k in range(1, 5))
            else:
                dx[SUN] = 0.0
                for i in range(SHA, ROOT + 1):
                    dx[i] = _fsum(A[i][k] * f[k] for k in range(SHA, ROOT + 1))
            mx = max(abs(dx[i]) for i in range(1, 5))
            if mx > 50000.0:
                for i in range(1, 5):
                    dx[i] = 50000.0 * dx[i] / mx
            if P.laisun > TOL_LAI and P.laisha > TOL_LAI:
                for i in range(1, 5):
                    x[i] = x[i] + dx[i]
            elif P.laisha > TOL_LAI:
                for i in range(1, 5):
                    x[i] = x[i] + dx[i]
                x[SUN] = x[XYL]                      # psi_sun = psi_xyl because laisun == 0 (:4628)
            else:
                x[XYL] = x[XYL] + dx[XYL]
                x[ROOT] = x[ROOT] + dx[ROOT]
                x[SUN] = x[SUN] + dx[SHA]            # dx(sun) and dx(sha) are flipped in the laisha == 0 case (:4631)
                x[SHA] = x[XYL]                      # psi_sha = psi_xyl because laisha == 0 (:4632)
            if math.sqrt(_fsum(dx[i] * dx[i] for i in range(1, 5))) < toldx:
                break
            if x[XYL] > x[ROOT]:
                x[XYL] = x[ROOT]
            if x[SUN] > x[XYL]:
                x[SUN] = x[XYL]
            if x[SHA] > x[XYL]:
                x[SHA] = x[XYL]
        iters = it
    else:
        flag = True
    if flag:
        getvegwp(P, x, gb_mol, gs0sun, gs0sha, qsatl, qaf)
        bsun = plc(x[SUN], P, SUN)
        bsha = plc(x[SHA], P, SHA)
    else:
        qsun = qflx_sun * plc(x[SUN], P, SUN)
        qsha = qflx_sha * plc(x[SHA], P, SHA)
        gs0sun, gs0sha, _, _ = getqflx(P, gb_mol, gs0sun, gs0sha, qsun, qsha, qsatl, qaf, False)
        bsun = gs0sun / gs_mol_sun if qflx_sun > 0.0 else plc(x[SUN], P, SUN)
        bsha = gs0sha / gs_mol_sha if qflx_sha > 0.0 else plc(x[SHA], P, SHA)
    if bsun < 0.01:
        bsun = 0.0
    if bsha < 0.01:
        bsha = 0.0
    tran = None
    if night:
        gs0sun = bsun * gs_mol_sun
        gs0sha = bsha * gs_mol_sha
        soilflux = getvegwp(P, x, gb_mol, gs0sun, gs0sha, qsatl, qaf)
        if soilflux < 0.0:
            soilflux = 0.0
        tran = soilflux
    pd = dict(x) if (night and P.local_time_lt_noon) else {i: SPVAL for i in range(1, 5)}
    return SimpleNamespace(bsun=bsun, bsha=bsha, night=night, tran=tran, iters=iters, vegwp_pd=pd)
