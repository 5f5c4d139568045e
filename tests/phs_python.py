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


# ------------------------------------------------------------------------------------------------------------------------------
# PhotosynthesisHydraulicStress and its root finder (PhotosynthesisMod.F90:2704-4486), nlevcan = 1, use_cn = .false.,
# lnc_opt = .false., vcmax_opt = 0, use_c13 = .false. (the configuration of the hot path)
# ------------------------------------------------------------------------------------------------------------------------------
RPI = 3.14159265358979323846
BBBOPT_C3, BBBOPT_C4 = 10000.0, 40000.0          # :85-86
MEDLYN_RH_CAN_MAX, MEDLYN_RH_CAN_FACT = 50.0, 0.001   # :87-88
MAX_CS = 1.e-06                                   # :89
BB1987, MEDLYN2011 = 1, 2
EPSILON = 2.220446049250313e-16


class EndRun(Exception):
    pass


def quadratic(a, b, c):
    """src/utils/quadraticMod.F90:17-68; returns (r1, r2)"""
    if a == 0.0:
        raise EndRun("quadratic: a = 0")
    root = b * b - 4.0 * a * c
    if root < 0.0:
        if -root < 3.0 * EPSILON:
            root = 0.0
        else:
            raise EndRun("quadratic: b^2 - 4ac is negative")
    if b >= 0.0:
        q = -0.5 * (b + math.sqrt(root))
    else:
        q = -0.5 * (b - math.sqrt(root))
    r1 = q / a
    if q != 0.0:
        r2 = c / q
    else:
        r2 = 1.e36
    return r1, r2


def ft(tl, ha):
    """:2488-2510"""
    return math.exp(ha / (RGAS * 1.e-3 * (TFRZ + 25.0)) * (1.0 - (TFRZ + 25.0) / tl))


def fth(tl, hd, se, scaleFactor):
    """:2513-2536"""
    return scaleFactor / (1.0 + math.exp((-hd + se * tl) / (RGAS * 1.e-3 * tl)))


def fth25(hd, se):
    """:2539-2561"""
    return 1.0 + math.exp((-hd + se * (TFRZ + 25.0)) / (RGAS * 1.e-3 * (TFRZ + 25.0)))


def _gsmin(P, M):
    if M.stomatalcond_mtd == MEDLYN2011:
        return P.medlynintercept
    if M.stomatalcond_mtd == BB1987:
        return P.bbb
    raise EndRun("must choose stomatalcond_mtd method")


def ci_func_PHS(P, M, W, x, cisun, cisha, bflag, gs0sun, gs0sha, fvalsun, fvalsha):
    """:4227-4486.  W carries what the routine keeps in photosyns_inst / its inout dummies: ac, aj, ap, ag, an (dicts keyed SUN,
    SHA), gs_mol (dict), bsun, bsha.  Returns (fvalsun, fvalsha); the incoming values survive where the routine does not assign."""
    if bflag:
        r = calcstress(P, x, P.gb_mol, gs0sun, gs0sha, P.qsatl, P.qaf)
        W.bsun, W.bsha = r.bsun, r.bsha
        W.vegwp_pd = r.vegwp_pd
        if r.tran is not None:
            W.qflx_tran_veg = r.tran
        W.calcstress_calls += 1
    bsun, bsha = W.bsun, W.bsha
    ac, aj, ap, ag, an = W.ac, W.aj, W.ap, W.ag, W.an
    if P.c3flag:
        ac[SUN] = bsun * P.vcmax_z[SUN] * max(cisun - P.cp, 0.0) / (cisun + P.kc * (1.0 + P.oair / P.ko))
        ac[SHA] = bsha * P.vcmax_z[SHA] * max(cisha - P.cp, 0.0) / (cisha + P.kc * (1.0 + P.oair / P.ko))
        aj[SUN] = P.je[SUN] * max(cisun - P.cp, 0.0) / (4.0 * cisun + 8.0 * P.cp)
        aj[SHA] = P.je[SHA] * max(cisha - P.cp, 0.0) / (4.0 * cisha + 8.0 * P.cp)
        ap[SUN] = 3.0 * P.tpu_z[SUN]
        ap[SHA] = 3.0 * P.tpu_z[SHA]
    else:
        ac[SUN] = bsun * P.vcmax_z[SUN]
        ac[SHA] = bsha * P.vcmax_z[SHA]
        aj[SUN] = P.qe * P.par_z[SUN] * 4.6
        aj[SHA] = P.qe * P.par_z[SHA] * 4.6
        ap[SUN] = P.kp_z[SUN] * max(cisun, 0.0) / P.forc_pbot
        ap[SHA] = P.kp_z[SHA] * max(cisha, 0.0) / P.forc_pbot
    for s in (SUN, SHA):
        r1, r2 = quadratic(P.theta_cj, -(ac[s] + aj[s]), ac[s] * aj[s])
        ai = min(r1, r2)
        r1, r2 = quadratic(M.theta_ip, -(ai + ap[s]), ai * ap[s])
        ag[s] = max(0.0, min(r1, r2))
    an[SUN] = ag[SUN] - bsun * P.lmr_z[SUN]
    an[SHA] = ag[SHA] - bsha * P.lmr_z[SHA]
    gs = W.gs_mol
    if an[SUN] < 0.0:
        gs[SUN] = max(bsun * _gsmin(P, M), 1.0)
        fvalsun = 0.0
    if an[SHA] < 0.0:
        gs[SHA] = max(bsha * _gsmin(P, M), 1.0)
        fvalsha = 0.0
    if an[SUN] < 0.0 and an[SHA] < 0.0:
        return fvalsun, fvalsha
    if an[SUN] >= 0.0:
        cs_sun = P.cair - 1.4 / P.gb_mol * an[SUN] * P.forc_pbot
        cs_sun = max(cs_sun, MAX_CS)
    if M.stomatalcond_mtd == MEDLYN2011:
        mi, ms = P.medlynintercept, P.medlynslope
        if an[SUN] >= 0.0:
            term = 1.6 * an[SUN] / (cs_sun / P.forc_pbot * 1.e06)
            aquad = 1.0
            bquad = -(2.0 * (mi * 1.e-06 + term) + (ms * term) * (ms * term) / (P.gb_mol * 1.e-06 * P.rh_can))
            cquad = mi * mi * 1.e-12 + (2.0 * mi * 1.e-06 + term * (1.0 - ms * ms / P.rh_can)) * term
            r1, r2 = quadratic(aquad, bquad, cquad)
            gs[SUN] = max(r1, r2) * 1.e06
        if an[SHA] >= 0.0:
            cs_sha = P.cair - 1.4 / P.gb_mol * an[SHA] * P.forc_pbot
            cs_sha = max(cs_sha, MAX_CS)
            term = 1.6 * an[SHA] / (cs_sha / P.forc_pbot * 1.e06)
            aquad = 1.0
            bquad = -(2.0 * (mi * 1.e-06 + term) + (ms * term) * (ms * term) / (P.gb_mol * 1.e-06 * P.rh_can))
            cquad = mi * mi * 1.e-12 + (2.0 * mi * 1.e-06 + term * (1.0 - ms * ms / P.rh_can)) * term
            r1, r2 = quadratic(aquad, bquad, cquad)
            gs[SHA] = max(r1, r2) * 1.e06
    elif M.stomatalcond_mtd == BB1987:
        if an[SUN] >= 0.0:
            aquad = cs_sun
            bquad = cs_sun * (P.gb_mol - max(bsun * P.bbb, 1.0)) - P.mbb * an[SUN] * P.forc_pbot
            cquad = -P.gb_mol * (cs_sun * max(bsun * P.bbb, 1.0) + P.mbb * an[SUN] * P.forc_pbot * P.rh_can)
            r1, r2 = quadratic(aquad, bquad, cquad)
            gs[SUN] = max(r1, r2)
        if an[SHA] >= 0.0:
            cs_sha = P.cair - 1.4 / P.gb_mol * an[SHA] * P.forc_pbot
            cs_sha = max(cs_sha, MAX_CS)
            aquad = cs_sha
            bquad = cs_sha * (P.gb_mol - max(bsha * P.bbb, 1.0)) - P.mbb * an[SHA] * P.forc_pbot
            cquad = -P.gb_mol * (cs_sha * max(bsha * P.bbb, 1.0) + P.mbb * an[SHA] * P.forc_pbot * P.rh_can)
            r1, r2 = quadratic(aquad, bquad, cquad)
            gs[SHA] = max(r1, r2)
    if an[SUN] >= 0.0:
        if gs[SUN] > 0.0:
            fvalsun = cisun - P.cair + an[SUN] * P.forc_pbot * (1.4 * gs[SUN] + 1.6 * P.gb_mol) / (P.gb_mol * gs[SUN])
        else:
            fvalsun = cisun - P.cair
    if an[SHA] >= 0.0:
        if gs[SHA] > 0.0:
            fvalsha = cisha - P.cair + an[SHA] * P.forc_pbot * (1.4 * gs[SHA] + 1.6 * P.gb_mol) / (P.gb_mol * gs[SHA])
        else:
            fvalsha = cisha - P.cair
    return fvalsun, fvalsha


def brent_PHS(P, M, W, x1sun, x2sun, f1sun, f2sun, x1sha, x2sha, f1sha, f2sha, tol):
    """:4068-4223; returns (xsun, xsha).  Arrays are dicts keyed SUN, SHA."""
    itmax, eps = 20, 1.e-4
    a = {SUN: x1sun, SHA: x1sha}
    b = {SUN: x2sun, SHA: x2sha}
    fa = {SUN: f1sun, SHA: f1sha}
    fb = {SUN: f2sun, SHA: f2sha}
    for ph in (SUN, SHA):
        if (fa[ph] > 0.0 and fb[ph] > 0.0) or (fa[ph] < 0.0 and fb[ph] < 0.0):
            raise EndRun("root must be bracketed for brent")
    c, fc = dict(b), dict(fb)
    d, e, s, pp_, q, r = {}, {}, {}, {}, {}, {}
    it = 0
    while True:
        if it == itmax:
            break
        it += 1
        for ph in (SUN, SHA):
            if (fb[ph] > 0.0 and fc[ph] > 0.0) or (fb[ph] < 0.0 and fc[ph] < 0.0):
                c[ph] = a[ph]
                fc[ph] = fa[ph]
                d[ph] = b[ph] - a[ph]
                e[ph] = d[ph]
            if abs(fc[ph]) < abs(fb[ph]):
                a[ph] = b[ph]
                b[ph] = c[ph]
                c[ph] = a[ph]
                fa[ph] = fb[ph]
                fb[ph] = fc[ph]
                fc[ph] = fa[ph]
        tol1 = {ph: 2.0 * eps * abs(b[ph]) + 0.5 * tol for ph in (SUN, SHA)}
        xm = {ph: 0.5 * (c[ph] - b[ph]) for ph in (SUN, SHA)}
        if abs(xm[SUN]) <= tol1[SUN] or fb[SUN] == 0.0:
            if abs(xm[SHA]) <= tol1[SHA] or fb[SHA] == 0.0:
                W.brent_iters += it
                return b[SUN], b[SHA]
        for ph in (SUN, SHA):
            if abs(e[ph]) >= tol1[ph] and abs(fa[ph]) > abs(fb[ph]):
                s[ph] = fb[ph] / fa[ph]
                if a[ph] == c[ph]:
                    pp_[ph] = 2.0 * xm[ph] * s[ph]
                    q[ph] = 1.0 - s[ph]
                else:
                    q[ph] = fa[ph] / fc[ph]
                    r[ph] = fb[ph] / fc[ph]
                    pp_[ph] = s[ph] * (2.0 * xm[ph] * q[ph] * (q[ph] - r[ph]) - (b[ph] - a[ph]) * (r[ph] - 1.0))
                    q[ph] = (q[ph] - 1.0) * (r[ph] - 1.0) * (s[ph] - 1.0)
                if pp_[ph] > 0.0:
                    q[ph] = -q[ph]
                pp_[ph] = abs(pp_[ph])
                if 2.0 * pp_[ph] < min(3.0 * xm[ph] * q[ph] - abs(tol1[ph] * q[ph]), abs(e[ph] * q[ph])):
                    e[ph] = d[ph]
                    d[ph] = pp_[ph] / q[ph]
                else:
                    d[ph] = xm[ph]
                    e[ph] = d[ph]
            else:
                d[ph] = xm[ph]
                e[ph] = d[ph]
            a[ph] = b[ph]
            fa[ph] = fb[ph]
            if abs(d[ph]) > tol1[ph]:
                b[ph] = b[ph] + d[ph]
            else:
                b[ph] = b[ph] + math.copysign(tol1[ph], xm[ph])
        gs0sun, gs0sha = W.gs_mol[SUN], W.gs_mol[SHA]
        fb[SUN], fb[SHA] = ci_func_PHS(P, M, W, None, b[SUN], b[SHA], False, gs0sun, gs0sha, fb[SUN], fb[SHA])
        if fb[SUN] == 0.0 and fb[SHA] == 0.0:
            break
    W.brent_iters += it
    return b[SUN], b[SHA]


def hybrid_PHS(P, M, W, x0sun, x0sha):
    """:3815-4064; returns (x0sun, x0sha) = the converged ci pair.  W.vegwp (dict 1..4) is vegwp(p,:)"""
    toldb, eps, eps1, itmax = 1.e-2, 1.e-2, 1.e-4, 3
    x1sun, x1sha = x0sun, x0sha
    bflag = False
    b0sun, b0sha = -1.0, -1.0
    gs0sun, gs0sha = 0.0, 0.0
    W.bsun, W.bsha = 1.0, 1.0
    f0sun = f0sha = f1sun = f1sha = float("nan")
    minf = minxsun = minxsha = float("nan")
    iter1 = 0
    while True:
        x = dict(W.vegwp)
        iter1 += 1
        iter2 = 0
        x0sun = max(0.1, x1sun)
        x1sun = 0.99 * x1sun
        x0sha = max(0.1, x1sha)
        x1sha = 0.99 * x1sha
        tolsun = abs(x1sun) * eps
        tolsha = abs(x1sha) * eps
        f0sun, f0sha = ci_func_PHS(P, M, W, x, x0sun, x0sha, bflag, gs0sun, gs0sha, f0sun, f0sha)
        dbsun = b0sun - W.bsun
        dbsha = b0sha - W.bsha
        b0sun, b0sha = W.bsun, W.bsha
        bflag = False
        f1sun, f1sha = ci_func_PHS(P, M, W, x, x1sun, x1sha, bflag, gs0sun, gs0sha, f1sun, f1sha)
        while True:
            if abs(f0sun) < eps1 and abs(f0sha) < eps1:
                x1sun, x1sha = x0sun, x0sha
                break
            if abs(f1sun) < eps1 and abs(f1sha) < eps1:
                break
            iter2 += 1
            if (f1sun - f0sun) == 0.0:
                dxsun = 0.5 * (x1sun + x0sun) - x1sun
            else:
                dxsun = -f1sun * (x1sun - x0sun) / (f1sun - f0sun)
            if (f1sha - f0sha) == 0.0:
                dxsha = 0.5 * (x1sha + x0sha) - x1sha
            else:
                dxsha = -f1sha * (x1sha - x0sha) / (f1sha - f0sha)
            x0sun = x1sun
            x1sun = x1sun + dxsun
            x0sha = x1sha
            x1sha = x1sha + dxsha
            f1sun, f1sha = ci_func_PHS(P, M, W, x, x1sun, x1sha, bflag, gs0sun, gs0sha, f1sun, f1sha)
            if abs(dxsun) < tolsun and abs(dxsha) < tolsha:
                x0sun, x0sha = x1sun, x1sha
                break
            if iter2 == 1:
                minf = abs(f1sun + f1sha)
                minxsun, minxsha = x1sun, x1sha
            else:
                if abs(f1sun + f1sha) < minf:
                    minf = abs(f1sun + f1sha)
                    minxsun, minxsha = x1sun, x1sha
            if abs(f1sun) < eps1 and abs(f1sha) < eps1:
                break
            if f1sun * f0sun < 0.0 and f1sha * f0sha < 0.0:
                xsun, xsha = brent_PHS(P, M, W, x0sun, x1sun, f0sun, f1sun, x0sha, x1sha, f0sha, f1sha, tolsun)
                x0sun, x0sha = xsun, xsha
                W.brent_calls += 1
                break
            if iter2 > itmax:
                x1sun, x1sha = minxsun, minxsha
                f1sun, f1sha = ci_func_PHS(P, M, W, x, x1sun, x1sha, bflag, gs0sun, gs0sha, f1sun, f1sha)
                W.itmax_exits += 1
                break
        if W.bsun > 0.01:
            gs0sun = W.gs_mol[SUN] / W.bsun
        if W.bsha > 0.01:
            gs0sha = W.gs_mol[SHA] / W.bsha
        bflag = True
        if abs(dbsun) < toldb and abs(dbsha) < toldb:
            break
        if iter1 > itmax:
            break
    x0sun, x0sha = x1sun, x1sha
    soilflux = getvegwp(P, x, P.gb_mol, W.gs_mol[SUN], W.gs_mol[SHA], P.qsatl, P.qaf)
    W.vegwp = dict(x)
    if P.near_local_noon:
        W.vegwp_ln = dict(W.vegwp)
    else:
        W.vegwp_ln = {i: SPVAL for i in range(1, 5)}
    if soilflux < 0.0:
        soilflux = 0.0
    W.qflx_tran_veg = soilflux
    W.iter1 += iter1
    return x0sun, x0sha


def photosynthesis_hydraulic_stress(P, M):
    """:2704-3811 for one patch (nlevcan = 1: the canopy-layer loops run iv = 1 .. nrad <= 1).  P: the patch's inputs (see
    tests/test_oracle_phs.py::phs_patch_inputs); M: the parameter / namelist values.  Returns the namespace W of everything the
    routine writes for that patch."""
    croot_lateral_length, c_to_b = 0.25, 2.0
    W = SimpleNamespace(calcstress_calls=0, brent_calls=0, brent_iters=0, itmax_exits=0, iter1=0, qflx_tran_veg=None, vegwp_pd=None,
                        vegwp_ln=None, vegwp=dict(P.vegwp))
    medlyn = M.stomatalcond_mtd == MEDLYN2011
    lmrc = fth25(M.lmrhd, M.lmrse)
    # root-soil interface conductance (:3063-3114)
    W.root_conductance, W.soil_conductance, W.k_soil_root = {}, {}, {}
    for j in range(1, NLEVSOI + 1):
        root_biomass_density = c_to_b * P.froot_carbon * P.rootfr[j] / P.dz[j]
        root_biomass_density = max(c_to_b * 1.0, root_biomass_density)
        root_cross_sec_area = RPI * (P.root_radius * P.root_radius)        # x**2 is x*x in Fortran
        root_length_density = root_biomass_density / (P.root_density * root_cross_sec_area)
        rai = (P.tsai + P.tlai) * P.froot_leaf * P.rootfr[j]
        croot_average_length = croot_lateral_length
        r_soil = math.sqrt(1. / (RPI * root_length_density))
        soil_conductance = min(P.hksat[j], P.hk_l[j]) / (1.e3 * r_soil)
        fs = plc(P.smp[j], P, ROOT)
        root_conductance = (fs * rai * P.krmax) / (croot_average_length + P.z[j])
        soil_conductance = max(soil_conductance, 1.e-16)
        root_conductance = max(root_conductance, 1.e-16)
        W.root_conductance[j] = root_conductance
        W.soil_conductance[j] = soil_conductance
        rs_resis = 1.0 / soil_conductance + 1.0 / root_conductance
        if rai * P.rootfr[j] > 0.0 and j > 1:
            W.k_soil_root[j] = 1.0 / rs_resis
        else:
            W.k_soil_root[j] = 0.0
    P.k = W.k_soil_root
    # :3118-3164
    if round(P.c3psn) == 1:
        P.c3flag = True
    elif round(P.c3psn) == 0:
        P.c3flag = False
    W.c3flag = P.c3flag
    if P.c3flag:
        P.qe = 0.0
        bbbopt = BBBOPT_C3
    else:
        P.qe = 0.05
        bbbopt = BBBOPT_C4
    if not medlyn:
        P.bbb = bbbopt
        P.mbb = P.mbbopt
    kc25 = M.kc25_coef * P.forc_pbot
    ko25 = M.ko25_coef * P.forc_pbot
    sco = 0.5 * 0.209 / M.cp25_yr2000
    cp25 = 0.5 * P.oair / sco
    P.kc = kc25 * ft(P.t_veg, M.kcha)
    P.ko = ko25 * ft(P.t_veg, M.koha)
    P.cp = cp25 * ft(P.t_veg, M.cpha)
    W.qe, W.kc, W.ko, W.cp = P.qe, P.kc, P.ko, P.cp
    # :3170-3469
    lnc = 1.0 / (P.slatop * P.leafcn)
    lnc = min(lnc, 10.0)
    W.lnc = lnc
    vcmax25top = lnc * P.flnr * M.fnr * M.act25 * P.dayl_factor
    vcmax25top = vcmax25top * P.fnitr                       # .not. use_cn
    t10c = min(max((P.t10 - TFRZ), 11.0), 35.0)
    jmax25top = ((2.59 - 0.035 * t10c) * vcmax25top) * M.jmax25top_sf
    tpu25top = M.tpu25ratio * vcmax25top
    kp25top = M.kp25ratio * vcmax25top
    W.luvcmax25top, W.lujmax25top, W.lutpu25top = vcmax25top, jmax25top, tpu25top
    if P.c3flag:
        lmr25top = vcmax25top * M.leaf_mr_vcm
    else:
        lmr25top = vcmax25top * 0.025
    luna = M.use_luna and P.c3flag and P.crop == 0
    P.vcmax_z, P.tpu_z, P.kp_z, P.lmr_z, jmax_z = {}, {}, {}, {}, {}
    t_veg = P.t_veg
    for iv in range(1, P.nrad + 1):
        nscaler_sun = P.vcmaxcintsun
        nscaler_sha = P.vcmaxcintsha
        lmr25_sun = lmr25top * nscaler_sun
        lmr25_sha = lmr25top * nscaler_sha
        if luna:
            lmr25_sun = M.leaf_mr_vcm * P.vcmx25_z
            lmr25_sha = M.leaf_mr_vcm * P.vcmx25_z
        if P.c3flag:
            P.lmr_z[SUN] = lmr25_sun * ft(t_veg, M.lmrha) * fth(t_veg, M.lmrhd, M.lmrse, lmrc)
            P.lmr_z[SHA] = lmr25_sha * ft(t_veg, M.lmrha) * fth(t_veg, M.lmrhd, M.lmrse, lmrc)
        else:
            P.lmr_z[SUN] = lmr25_sun * 2.0 ** ((t_veg - (TFRZ + 25.0)) / 10.0)
            P.lmr_z[SUN] = P.lmr_z[SUN] / (1.0 + math.exp(1.3 * (t_veg - (TFRZ + 55.0))))
            P.lmr_z[SHA] = lmr25_sha * 2.0 ** ((t_veg - (TFRZ + 25.0)) / 10.0)
            P.lmr_z[SHA] = P.lmr_z[SHA] / (1.0 + math.exp(1.3 * (t_veg - (TFRZ + 55.0))))
        P.lmr_z[SUN] = P.lmr_z[SUN] * min((0.2 * math.exp(3.218 * P.tlai_z)), 1.0)
        P.lmr_z[SHA] = P.lmr_z[SHA] * min((0.2 * math.exp(3.218 * P.tlai_z)), 1.0)
        if P.par_z[SUN] <= 0.0:
            for s in (SUN, SHA):
                P.vcmax_z[s] = 0.0
                jmax_z[s] = 0.0
                P.tpu_z[s] = 0.0
                P.kp_z[s] = 0.0
        else:
            if luna:
                vcmax25_sun = P.vcmx25_z
                vcmax25_sha = P.vcmx25_z
                jmax25_sun = P.jmx25_z
                jmax25_sha = P.jmx25_z
                tpu25_sun = M.tpu25ratio * vcmax25_sun
                tpu25_sha = M.tpu25ratio * vcmax25_sha
                if P.vcmaxcintsun > 0.0:                      # .and. nlevcan == 1
                    vcmax25_sha = vcmax25_sun * P.vcmaxcintsha / P.vcmaxcintsun
                    jmax25_sha = jmax25_sun * P.vcmaxcintsha / P.vcmaxcintsun
                    tpu25_sha = tpu25_sun * P.vcmaxcintsha / P.vcmaxcintsun
            else:
                vcmax25_sun = vcmax25top * nscaler_sun
                jmax25_sun = jmax25top * nscaler_sun
                tpu25_sun = tpu25top * nscaler_sun
                vcmax25_sha = vcmax25top * nscaler_sha
                jmax25_sha = jmax25top * nscaler_sha
                tpu25_sha = tpu25top * nscaler_sha
            kp25_sun = kp25top * nscaler_sun
            kp25_sha = kp25top * nscaler_sha
            vcmaxse = (668.39 - 1.07 * t10c) * M.vcmaxse_sf
            jmaxse = (659.70 - 0.75 * t10c) * M.jmaxse_sf
            tpuse = (668.39 - 1.07 * t10c) * M.tpuse_sf
            vcmaxc = fth25(M.vcmaxhd, vcmaxse)
            jmaxc = fth25(M.jmaxhd, jmaxse)
            tpuc = fth25(M.tpuhd, tpuse)
            P.vcmax_z[SUN] = vcmax25_sun * ft(t_veg, M.vcmaxha) * fth(t_veg, M.vcmaxhd, vcmaxse, vcmaxc)
            jmax_z[SUN] = jmax25_sun * ft(t_veg, M.jmaxha) * fth(t_veg, M.jmaxhd, jmaxse, jmaxc)
            P.tpu_z[SUN] = tpu25_sun * ft(t_veg, M.tpuha) * fth(t_veg, M.tpuhd, tpuse, tpuc)
            P.vcmax_z[SHA] = vcmax25_sha * ft(t_veg, M.vcmaxha) * fth(t_veg, M.vcmaxhd, vcmaxse, vcmaxc)
            jmax_z[SHA] = jmax25_sha * ft(t_veg, M.jmaxha) * fth(t_veg, M.jmaxhd, jmaxse, jmaxc)
            P.tpu_z[SHA] = tpu25_sha * ft(t_veg, M.tpuha) * fth(t_veg, M.tpuhd, tpuse, tpuc)
            if not P.c3flag:
                for s, v25 in ((SUN, vcmax25_sun), (SHA, vcmax25_sha)):
                    P.vcmax_z[s] = v25 * 2.0 ** ((t_veg - (TFRZ + 25.0)) / 10.0)
                    P.vcmax_z[s] = P.vcmax_z[s] / (1.0 + math.exp(0.2 * ((TFRZ + 15.0) - t_veg)))
                    P.vcmax_z[s] = P.vcmax_z[s] / (1.0 + math.exp(0.3 * (t_veg - (TFRZ + 40.0))))
            P.kp_z[SUN] = kp25_sun * 2.0 ** ((t_veg - (TFRZ + 25.0)) / 10.0)
            P.kp_z[SHA] = kp25_sha * 2.0 ** ((t_veg - (TFRZ + 25.0)) / 10.0)
        if M.light_inhibit and P.par_z[SUN] > 0.0:
            P.lmr_z[SUN] = P.lmr_z[SUN] * 0.67
        if M.light_inhibit and P.par_z[SHA] > 0.0:
            P.lmr_z[SHA] = P.lmr_z[SHA] * 0.67
    W.vcmax_z, W.tpu_z, W.kp_z, W.lmr_z = P.vcmax_z, P.tpu_z, P.kp_z, P.lmr_z
    # leaf-level photosynthesis and stomatal conductance (:3475-3711)
    rsmax0 = 2.e4
    cf = P.forc_pbot / (RGAS * 1.e-3 * P.tgcm) * 1.e06
    gb = 1.0 / P.rb
    P.gb_mol = gb * cf
    W.gb_mol = P.gb_mol
    W.ac, W.aj, W.ap, W.ag, W.an, W.gs_mol = {}, {}, {}, {}, {}, dict(P.gs_mol)
    W.bsun, W.bsha = P.bsun_in, P.bsha_in
    W.psn_z, W.psn_wc_z, W.psn_wj_z, W.psn_wp_z, W.rs_z, W.ci_z, W.gs_mol_ln = {}, {}, {}, {}, {}, {}, None
    W.vpd_can = None
    lessen = P.crop == 0 or not M.modifyphoto_and_lmr_forcrop
    for iv in range(1, P.nrad + 1):
        if P.par_z[SUN] <= 0.0:
            W.vegwp[SUN] = 1.0
            gsminsun = gsminsha = _gsmin(P, M)
            x = W.vegwp                                               # calcstress works on vegwp(p,:) itself here
            r = calcstress(P, x, P.gb_mol, gsminsun, gsminsha, P.qsatl, P.qaf)
            W.calcstress_calls += 1
            W.bsun, W.bsha, W.vegwp_pd = r.bsun, r.bsha, r.vegwp_pd
            if r.tran is not None:
                W.qflx_tran_veg = r.tran
            for s, b, gsmin in ((SUN, W.bsun, gsminsun), (SHA, W.bsha, gsminsha)):
                W.ac[s] = W.aj[s] = W.ap[s] = W.ag[s] = 0.0
                if lessen:
                    W.an[s] = W.ag[s] - b * P.lmr_z[s]
                else:
                    W.an[s] = W.ag[s] - P.lmr_z[s]
                W.psn_z[s] = W.psn_wc_z[s] = W.psn_wj_z[s] = W.psn_wp_z[s] = 0.0
                W.rs_z[s] = min(rsmax0, 1.0 / (max(b * gsmin, 1.0)) * cf)
                W.ci_z[s] = 0.0
            W.gs_mol[SUN] = cf / W.rs_z[SUN]
            W.gs_mol[SHA] = cf / W.rs_z[SHA]
        else:
            ceair = min(P.eair, P.esat_tv)
            if not medlyn:
                P.rh_can = ceair / P.esat_tv
            else:
                P.rh_can = max((P.esat_tv - ceair), MEDLYN_RH_CAN_MAX) * MEDLYN_RH_CAN_FACT
                W.vpd_can = P.rh_can
            P.je = {}
            for s in (SUN, SHA):
                qabs = 0.5 * (1.0 - M.fnps) * P.par_z[s] * 4.6
                r1, r2 = quadratic(M.theta_psii, -(qabs + jmax_z[s]), qabs * jmax_z[s])
                P.je[s] = min(r1, r2)
            ci0 = 0.7 * P.cair if P.c3flag else 0.4 * P.cair
            hybrid_PHS(P, M, W, ci0, ci0)
            if medlyn:
                gsminsun = gsminsha = P.medlynintercept
                gs_slope = P.medlynslope
            else:
                gsminsun = gsminsha = P.bbb
                gs_slope = P.mbb
            if W.an[SUN] < 0.0:
                W.gs_mol[SUN] = max(W.bsun * gsminsun, 1.0)
            if W.an[SHA] < 0.0:
                W.gs_mol[SHA] = max(W.bsha * gsminsha, 1.0)
            if P.near_local_noon:
                W.gs_mol_ln = dict(W.gs_mol)
            else:
                W.gs_mol_ln = {SUN: SPVAL, SHA: SPVAL}
            cs = {}
            for s in (SUN, SHA):
                cs[s] = P.cair - 1.4 / P.gb_mol * W.an[s] * P.forc_pbot
                cs[s] = max(cs[s], MAX_CS)
                W.ci_z[s] = P.cair - W.an[s] * P.forc_pbot * (1.4 * W.gs_mol[s] + 1.6 * P.gb_mol) / (P.gb_mol * W.gs_mol[s])
                W.ci_z[s] = max(W.ci_z[s], 1.e-06)
            for s in (SUN, SHA):
                gs = W.gs_mol[s] / cf
                W.rs_z[s] = min(1.0 / gs, rsmax0)
                W.rs_z[s] = W.rs_z[s] / P.o3coefg[s]
            for s in (SUN, SHA):
                W.psn_z[s] = W.ag[s]
                W.psn_z[s] = W.psn_z[s] * P.o3coefv[s]
                W.psn_wc_z[s] = W.psn_wj_z[s] = W.psn_wp_z[s] = 0.0
                if W.ac[s] <= W.aj[s] and W.ac[s] <= W.ap[s]:
                    W.psn_wc_z[s] = W.psn_z[s]
                elif W.aj[s] < W.ac[s] and W.aj[s] <= W.ap[s]:
                    W.psn_wj_z[s] = W.psn_z[s]
                elif W.ap[s] < W.ac[s] and W.ap[s] < W.aj[s]:
                    W.psn_wp_z[s] = W.psn_z[s]
            if W.gs_mol[SUN] < 0.0 or W.gs_mol[SHA] < 0.0:
                raise EndRun("Negative stomatal conductance")
            W.gs_mol_err = {}
            for s, b, gsmin in ((SUN, W.bsun, gsminsun), (SHA, W.bsha, gsminsha)):
                hs = (P.gb_mol * ceair + W.gs_mol[s] * P.esat_tv) / ((P.gb_mol + W.gs_mol[s]) * P.esat_tv)
                W.gs_mol_err[s] = gs_slope * max(W.an[s], 0.0) * hs / cs[s] * P.forc_pbot + max(b * gsmin, 1.0)
    # canopy sums (:3715-3807)
    W.psn, W.psn_wc, W.psn_wj, W.psn_wp, W.lmr, W.rs = {}, {}, {}, {}, {}, {}
    laican = {}
    for s, b in ((SUN, W.bsun), (SHA, W.bsha)):
        psncan = psncan_wc = psncan_wj = psncan_wp = lmrcan = gscan = 0.0
        laican[s] = 0.0
        for iv in range(1, P.nrad + 1):
            psncan = psncan + W.psn_z[s] * P.lai_z[s]
            psncan_wc = psncan_wc + W.psn_wc_z[s] * P.lai_z[s]
            psncan_wj = psncan_wj + W.psn_wj_z[s] * P.lai_z[s]
            psncan_wp = psncan_wp + W.psn_wp_z[s] * P.lai_z[s]
            if P.crop == 0 and M.modifyphoto_and_lmr_forcrop:
                lmrcan = lmrcan + P.lmr_z[s] * P.lai_z[s] * b
            else:
                lmrcan = lmrcan + P.lmr_z[s] * P.lai_z[s]
            gscan = gscan + P.lai_z[s] / (P.rb + W.rs_z[s])
            laican[s] = laican[s] + P.lai_z[s]
        if laican[s] > 0.0:
            W.psn[s] = psncan / laican[s]
            W.psn_wc[s] = psncan_wc / laican[s]
            W.psn_wj[s] = psncan_wj / laican[s]
            W.psn_wp[s] = psncan_wp / laican[s]
            W.lmr[s] = lmrcan / laican[s]
            W.rs[s] = laican[s] / gscan - P.rb
        else:
            W.psn[s] = W.psn_wc[s] = W.psn_wj[s] = W.psn_wp[s] = W.lmr[s] = W.rs[s] = 0.0
    if laican[SHA] + laican[SUN] > 0.0:
        W.btran = W.bsun * (laican[SUN] / (laican[SUN] + laican[SHA])) + W.bsha * (laican[SHA] / (laican[SUN] + laican[SHA]))
    else:
        W.btran = W.bsun
    return W


# ------------------------------------------------------------------------------------------------------------------------------
# Photosynthesis without plant hydraulic stress (PhotosynthesisMod.F90:1243-2062) and its root finder hybrid / brent / ci_func
# (:2251-2485, :2564-2701), one leaf class (phase 'sun' or 'sha') per call, same configuration as above
# ------------------------------------------------------------------------------------------------------------------------------
def ci_func(P, M, W, ci):
    """:2564-2701; W carries ac, aj, ap, ag, an, gs_mol of the patch; returns fval"""
    if P.c3flag:
        W.ac = W.vcmax_z * max(ci - P.cp, 0.0) / (ci + P.kc * (1.0 + P.oair / P.ko))
        W.aj = W.je * max(ci - P.cp, 0.0) / (4.0 * ci + 8.0 * P.cp)
        W.ap = 3.0 * W.tpu_z
    else:
        W.ac = W.vcmax_z
        W.aj = P.qe * W.par_z * 4.6
        W.ap = W.kp_z * max(ci, 0.0) / P.forc_pbot
    r1, r2 = quadratic(P.theta_cj, -(W.ac + W.aj), W.ac * W.aj)
    ai = min(r1, r2)
    r1, r2 = quadratic(M.theta_ip, -(ai + W.ap), ai * W.ap)
    W.ag = max(0.0, min(r1, r2))
    W.an = W.ag - W.lmr_z
    W.ci_evals += 1
    if W.an < 0.0:
        return 0.0
    cs = P.cair - 1.4 / P.gb_mol * W.an * P.forc_pbot
    cs = max(cs, MAX_CS)
    if M.stomatalcond_mtd == MEDLYN2011:
        mi, ms = P.medlynintercept, P.medlynslope
        term = 1.6 * W.an / (cs / P.forc_pbot * 1.e06)
        aquad = 1.0
        bquad = -(2.0 * (mi * 1.e-06 + term) + (ms * term) * (ms * term) / (P.gb_mol * 1.e-06 * W.rh_can))
        cquad = mi * mi * 1.e-12 + (2.0 * mi * 1.e-06 + term * (1.0 - ms * ms / W.rh_can)) * term
        r1, r2 = quadratic(aquad, bquad, cquad)
        W.gs_mol = max(r1, r2) * 1.e06
    elif M.stomatalcond_mtd == BB1987:
        aquad = cs
        bquad = cs * (P.gb_mol - P.bbb) - P.mbb * W.an * P.forc_pbot
        cquad = -P.gb_mol * (cs * P.bbb + P.mbb * W.an * P.forc_pbot * W.rh_can)
        r1, r2 = quadratic(aquad, bquad, cquad)
        W.gs_mol = max(r1, r2)
    return ci - P.cair + W.an * P.forc_pbot * (1.4 * W.gs_mol + 1.6 * P.gb_mol) / (P.gb_mol * W.gs_mol)


def brent(P, M, W, x1, x2, f1, f2, tol):
    """:2371-2485; returns the root"""
    itmax, eps = 20, 1.e-2
    a, b, fa, fb = x1, x2, f1, f2
    if (fa > 0.0 and fb > 0.0) or (fa < 0.0 and fb < 0.0):
        raise EndRun("root must be bracketed for brent")
    c, fc = b, fb
    d = e = None
    it = 0
    while True:
        if it == itmax:
            break
        it += 1
        if (fb > 0.0 and fc > 0.0) or (fb < 0.0 and fc < 0.0):
            c = a
            fc = fa
            d = b - a
            e = d
        if abs(fc) < abs(fb):
            a = b
            b = c
            c = a
            fa = fb
            fb = fc
            fc = fa
        tol1 = 2.0 * eps * abs(b) + 0.5 * tol
        xm = 0.5 * (c - b)
        if abs(xm) <= tol1 or fb == 0.0:
            W.brent_iters += it
            return b
        if abs(e) >= tol1 and abs(fa) > abs(fb):
            s = fb / fa
            if a == c:
                p = 2.0 * xm * s
                q = 1.0 - s
            else:
                q = fa / fc
                r = fb / fc
                p = s * (2.0 * xm * q * (q - r) - (b - a) * (r - 1.0))
                q = (q - 1.0) * (r - 1.0) * (s - 1.0)
            if p > 0.0:
                q = -q
            p = abs(p)
            if 2.0 * p < min(3.0 * xm * q - abs(tol1 * q), abs(e * q)):
                e = d
                d = p / q
            else:
                d = xm
                e = d
        else:
            d = xm
            e = d
        a = b
        fa = fb
        if abs(d) > tol1:
            b = b + d
        else:
            b = b + math.copysign(tol1, xm)
        fb = ci_func(P, M, W, b)
        if fb == 0.0:
            break
    W.brent_iters += it
    return b


def hybrid(P, M, W, x0):
    """:2251-2368; returns x0 (the ci the routine hands back)"""
    eps, eps1, itmax = 1.e-2, 1.e-4, 40
    f0 = ci_func(P, M, W, x0)
    if f0 == 0.0:
        return x0
    minx, minf = x0, f0
    x1 = x0 * 0.99
    f1 = ci_func(P, M, W, x1)
    if f1 == 0.0:
        return x1
    if f1 < minf:
        minx, minf = x1, f1
    it = 0
    while True:
        it += 1
        dx = -f1 * (x1 - x0) / (f1 - f0)
        x = x1 + dx
        tol = abs(x) * eps
        if abs(dx) < tol:
            x0 = x
            break
        x0 = x1
        f0 = f1
        x1 = x
        f1 = ci_func(P, M, W, x1)
        if f1 < minf:
            minx, minf = x1, f1
        if abs(f1) <= eps1:
            x0 = x1
            break
        if f1 * f0 < 0.0:
            x = brent(P, M, W, x0, x1, f0, f1, tol)
            W.brent_calls += 1
            x0 = x
            break
        if it > itmax:
            f1 = ci_func(P, M, W, minx)
            W.itmax_exits += 1
            break
    return x0


def photosynthesis(P, M, phase, btran):
    """:1243-2062 for one patch and one leaf class (phase SUN or SHA), nlevcan = 1.  Sets the patch-level values both phases share
    on P (c3flag, qe, kc, ko, cp, bbb, mbb, gb_mol) and returns the namespace W of what the call writes."""
    W = SimpleNamespace(ci_evals=0, brent_calls=0, brent_iters=0, itmax_exits=0, gs_mol=P.gs_mol_in, vpd_can=None, gs_mol_ln=None)
    medlyn = M.stomatalcond_mtd == MEDLYN2011
    W.par_z = P.par_z[phase]
    lai_z = P.lai_z[phase]
    vcmaxcint = P.vcmaxcintsun if phase == SUN else P.vcmaxcintsha
    lmrc = fth25(M.lmrhd, M.lmrse)
    if round(P.c3psn) == 1:
        P.c3flag = True
    elif round(P.c3psn) == 0:
        P.c3flag = False
    if P.c3flag:
        P.qe = 0.0
        bbbopt = BBBOPT_C3
    else:
        P.qe = 0.05
        bbbopt = BBBOPT_C4
    if not medlyn:
        P.bbb = max(bbbopt * btran, 1.0)
        P.mbb = P.mbbopt
    kc25 = M.kc25_coef * P.forc_pbot
    ko25 = M.ko25_coef * P.forc_pbot
    sco = 0.5 * 0.209 / M.cp25_yr2000
    cp25 = 0.5 * P.oair / sco
    P.kc = kc25 * ft(P.t_veg, M.kcha)
    P.ko = ko25 * ft(P.t_veg, M.koha)
    P.cp = cp25 * ft(P.t_veg, M.cpha)
    W.c3flag, W.qe, W.kc, W.ko, W.cp = P.c3flag, P.qe, P.kc, P.ko, P.cp
    if (P.slatop * P.leafcn) <= 0.0:
        raise EndRun("slatop or leafcn is zero")
    lnc = 1.0 / (P.slatop * P.leafcn)
    lnc = min(lnc, 10.0)
    W.lnc = lnc
    vcmax25top = lnc * P.flnr * M.fnr * M.act25 * P.dayl_factor
    vcmax25top = vcmax25top * P.fnitr
    t10c = min(max((P.t10 - TFRZ), 11.0), 35.0)
    jmax25top = ((2.59 - 0.035 * t10c) * vcmax25top) * M.jmax25top_sf
    tpu25top = M.tpu25ratio * vcmax25top
    kp25top = M.kp25ratio * vcmax25top
    lmr25top = vcmax25top * M.leaf_mr_vcm if P.c3flag else vcmax25top * 0.025
    luna = M.use_luna and P.c3flag and P.crop == 0
    t_veg = P.t_veg
    jmax_z = 0.0
    for iv in range(1, P.nrad + 1):
        nscaler = vcmaxcint
        lmr25 = lmr25top * nscaler
        if luna:
            lmr25 = M.leaf_mr_vcm * P.vcmx25_z
        if P.c3flag:
            W.lmr_z = lmr25 * ft(t_veg, M.lmrha) * fth(t_veg, M.lmrhd, M.lmrse, lmrc)
        else:
            W.lmr_z = lmr25 * 2.0 ** ((t_veg - (TFRZ + 25.0)) / 10.0)
            W.lmr_z = W.lmr_z / (1.0 + math.exp(1.3 * (t_veg - (TFRZ + 55.0))))
        if W.par_z <= 0.0:
            W.vcmax_z = jmax_z = W.tpu_z = W.kp_z = 0.0
        else:
            if luna:
                vcmax25 = P.vcmx25_z
                jmax25 = P.jmx25_z
                tpu25 = M.tpu25ratio * vcmax25
                if phase == SHA and P.vcmaxcintsun > 0.0:
                    vcmax25 = vcmax25 * P.vcmaxcintsha / P.vcmaxcintsun
                    jmax25 = jmax25 * P.vcmaxcintsha / P.vcmaxcintsun
                    tpu25 = tpu25 * P.vcmaxcintsha / P.vcmaxcintsun
            else:
                vcmax25 = vcmax25top * nscaler
                jmax25 = jmax25top * nscaler
                tpu25 = tpu25top * nscaler
            kp25 = kp25top * nscaler
            vcmaxse = (668.39 - 1.07 * t10c) * M.vcmaxse_sf
            jmaxse = (659.70 - 0.75 * t10c) * M.jmaxse_sf
            tpuse = (668.39 - 1.07 * t10c) * M.tpuse_sf
            vcmaxc = fth25(M.vcmaxhd, vcmaxse)
            jmaxc = fth25(M.jmaxhd, jmaxse)
            tpuc = fth25(M.tpuhd, tpuse)
            W.vcmax_z = vcmax25 * ft(t_veg, M.vcmaxha) * fth(t_veg, M.vcmaxhd, vcmaxse, vcmaxc)
            jmax_z = jmax25 * ft(t_veg, M.jmaxha) * fth(t_veg, M.jmaxhd, jmaxse, jmaxc)
            W.tpu_z = tpu25 * ft(t_veg, M.tpuha) * fth(t_veg, M.tpuhd, tpuse, tpuc)
            if not P.c3flag:
                W.vcmax_z = vcmax25 * 2.0 ** ((t_veg - (TFRZ + 25.0)) / 10.0)
                W.vcmax_z = W.vcmax_z / (1.0 + math.exp(0.2 * ((TFRZ + 15.0) - t_veg)))
                W.vcmax_z = W.vcmax_z / (1.0 + math.exp(0.3 * (t_veg - (TFRZ + 40.0))))
            W.kp_z = kp25 * 2.0 ** ((t_veg - (TFRZ + 25.0)) / 10.0)
        W.vcmax_z = W.vcmax_z * btran
        W.lmr_z = W.lmr_z * btran
        if M.light_inhibit and W.par_z > 0.0:
            W.lmr_z = W.lmr_z * 0.67
    rsmax0 = 2.e4
    cf = P.forc_pbot / (RGAS * 1.e-3 * P.tgcm) * 1.e06
    gb = 1.0 / P.rb
    P.gb_mol = gb * cf
    W.gb_mol = P.gb_mol
    for iv in range(1, P.nrad + 1):
        if W.par_z <= 0.0:
            W.ac = W.aj = W.ap = W.ag = 0.0
            W.an = W.ag - W.lmr_z
            W.psn_z = W.psn_wc_z = W.psn_wj_z = W.psn_wp_z = 0.0
            if not medlyn:
                W.rs_z = min(rsmax0, 1.0 / P.bbb * cf)
            else:
                W.rs_z = min(rsmax0, 1.0 / P.medlynintercept * cf)
            W.ci_z = 0.0
            W.gs_mol_phase = cf / W.rs_z
        else:
            ceair = min(P.eair, P.esat_tv)
            if not medlyn:
                W.rh_can = ceair / P.esat_tv
            else:
                W.rh_can = max((P.esat_tv - ceair), MEDLYN_RH_CAN_MAX) * MEDLYN_RH_CAN_FACT
                W.vpd_can = W.rh_can
            qabs = 0.5 * (1.0 - M.fnps) * W.par_z * 4.6
            r1, r2 = quadratic(M.theta_psii, -(qabs + jmax_z), qabs * jmax_z)
            W.je = min(r1, r2)
            W.ci_z = 0.7 * P.cair if P.c3flag else 0.4 * P.cair
            ciold = W.ci_z
            hybrid(P, M, W, ciold)
            if W.an < 0.0:
                W.gs_mol = P.bbb if not medlyn else P.medlynintercept
            W.gs_mol_phase = W.gs_mol
            W.gs_mol_ln = W.gs_mol if P.near_local_noon else SPVAL
            cs = P.cair - 1.4 / P.gb_mol * W.an * P.forc_pbot
            cs = max(cs, MAX_CS)
            W.ci_z = P.cair - W.an * P.forc_pbot * (1.4 * W.gs_mol + 1.6 * P.gb_mol) / (P.gb_mol * W.gs_mol)
            W.ci_z = max(W.ci_z, 1.e-06)
            gs = W.gs_mol / cf
            W.rs_z = min(1.0 / gs, rsmax0)
            W.rs_z = W.rs_z / P.o3coefg[phase]
            W.psn_z = W.ag
            W.psn_z = W.psn_z * P.o3coefv[phase]
            W.psn_wc_z = W.psn_wj_z = W.psn_wp_z = 0.0
            if W.ac <= W.aj and W.ac <= W.ap:
                W.psn_wc_z = W.psn_z
            elif W.aj < W.ac and W.aj <= W.ap:
                W.psn_wj_z = W.psn_z
            elif W.ap < W.ac and W.ap < W.aj:
                W.psn_wp_z = W.psn_z
            if W.gs_mol < 0.0:
                raise EndRun("Negative stomatal conductance")
    psncan = psncan_wc = psncan_wj = psncan_wp = lmrcan = gscan = laican = 0.0
    for iv in range(1, P.nrad + 1):
        psncan = psncan + W.psn_z * lai_z
        psncan_wc = psncan_wc + W.psn_wc_z * lai_z
        psncan_wj = psncan_wj + W.psn_wj_z * lai_z
        psncan_wp = psncan_wp + W.psn_wp_z * lai_z
        lmrcan = lmrcan + W.lmr_z * lai_z
        gscan = gscan + lai_z / (P.rb + W.rs_z)
        laican = laican + lai_z
    if laican > 0.0:
        W.psn, W.psn_wc, W.psn_wj, W.psn_wp = psncan / laican, psncan_wc / laican, psncan_wj / laican, psncan_wp / laican
        W.lmr = lmrcan / laican
        W.rs = laican / gscan - P.rb
    else:
        W.psn = W.psn_wc = W.psn_wj = W.psn_wp = W.lmr = W.rs = 0.0
    return W
