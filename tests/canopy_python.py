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


# ------------------------------------------------------------------------------------------------------------------------------
# CanopyFluxes for one exposed-vegetation patch (CanopyFluxesMod.F90:191-1765; use_fates = use_lch4 = use_cn = .false., nlevcan = 1,
# perchroot = .false.).  The filter loops of the reference become the life of one patch: it leaves the iteration when it converged.
# ------------------------------------------------------------------------------------------------------------------------------
SB = 5.67e-8
CPAIR = 1.00464e3
HVAP = 2.501e6
DENH2O = 1.000e3
DENICE = 0.917e3
C_WATER = 4.188e3
C_DRY_BIOMASS = 1400.0
C_TO_B = 2.0
ALPHA_AERO = 1.0
TLSAI_CRIT = 2.0
NU_PARAM = 1.5e-5
CD1_PARAM = 7.5
NLEVGRND = 25
SPVAL = 1.0e36


def _p3(x):
    """x**3 with an integer exponent, as the compiler expands it"""
    return (x * x) * x


def _p4(x):
    """x**4 with an integer exponent: two squarings"""
    x2 = x * x
    return x2 * x2


def root_moist_stress(P):
    """calc_effective_soilporosity, calc_volumetric_h2oliq, calc_root_moist_stress (SoilMoistStressMod.F90:70-514, clm45 default
    method, perchroot off) + soil_suction (SoilWaterRetentionCurveClappHornberg1978Mod.F90:87-124).
    Returns (eff_porosity, h2osoi_liqvol, rresis (None where the reference leaves it), rootr, btran)"""
    btran0 = 0.0
    eff_por, vol_liq, rresis, rootr = {}, {}, {}, {}
    for j in range(1, NLEVGRND + 1):
        vol_ice = min(P.watsat[j], P.h2osoi_ice[j] / (DENICE * P.dz[j]))
        eff_por[j] = P.watsat[j] - vol_ice
    for j in range(1, NLEVGRND + 1):
        vol_liq[j] = min(eff_por[j], P.h2osoi_liq[j] / (P.dz[j] * DENH2O))
    btran = btran0
    for j in range(1, NLEVGRND + 1):
        rresis[j] = None
        if vol_liq[j] <= 0.0 or P.t_soisno[j] <= TKFRZ - 2.0:
            rootr[j] = 0.0
        else:
            s_node = max(vol_liq[j] / eff_por[j], 0.01)
            smp_node = -P.sucsat[j] * s_node ** (-P.bsw[j])
            smp_node = max(P.smpsc, smp_node)
            rresis[j] = min((eff_por[j] / P.watsat[j]) * (smp_node - P.smpsc) / (P.smpso - P.smpsc), 1.0)
            rootr[j] = P.rootfr[j] * rresis[j]
            btran = btran + max(rootr[j], 0.0)
    for j in range(1, NLEVGRND + 1):
        if btran > btran0:
            rootr[j] = rootr[j] / btran
        else:
            rootr[j] = 0.0
    return eff_por, vol_liq, rresis, rootr, btran


def canopy_fluxes_patch(P, M, phs):
    """P: the patch's inputs (tests/test_oracle_canopy_pin.py::canopy_patch_inputs); M: parameters; phs(P, M): PhotosynthesisHydraulicStress
    for the patch (tests/phs_python.py), which reads P.esat_tv, eair, oair, cair, rb, dayl_factor, qsatl, qaf, t_veg, vegwp, gs_mol ...
    Returns the namespace O of everything the routine writes for the patch and its column."""
    btran0, zii, beta, delmax, dlemin, dtmin, itmin, ria = 0.0, 1000.0, 1.0, 1.0, 0.1, 0.01, 2, 0.5
    k_vert, k_cyl_vol, k_cyl_area, k_internal, min_stem_diameter, min_lai = 0.1, 1.0, 1.0, 0.0, 0.05, 0.1
    dtime = M.dtime
    O = SimpleNamespace()
    del_ = efeb = wtlq0 = wtalq = wtgq = wtaq0 = obuold = 0.0
    O.dhsdt_canopy = 0.0
    eflx_sh_stem = 0.0
    elai, esai, htop = P.elai, P.esai, P.htop
    stem_biomass, leaf_biomass = P.stem_biomass, P.leaf_biomass
    if M.use_biomass_heat_storage:
        frac_rad_abs_by_stem = (esai) / (elai + esai)
        if elai > 0.0:
            frac_rad_abs_by_stem = k_vert * frac_rad_abs_by_stem
        dbh = P.dbh_param
        sa_leaf = elai
        sa_leaf = 2.0 * sa_leaf
        sa_stem = P.nstem * (htop * PI * dbh)
        sa_stem = k_cyl_area * sa_stem
        if (not (P.is_tree or P.is_shrub)) or dbh < min_stem_diameter:
            frac_rad_abs_by_stem = 0.0
            sa_stem = 0.0
            sa_leaf = sa_leaf + esai
        else:
            if elai < min_lai:
                sa_leaf = sa_leaf + esai
        leaf_biomass = (1.e-3 * C_TO_B / P.slatop) * max(0.01, 0.5 * sa_leaf) / (1.0 - P.fbw)
        carea_stem = PI * ((dbh * 0.5) * (dbh * 0.5))
        stem_biomass = carea_stem * htop * k_cyl_vol * P.nstem * P.wood_density / (1.0 - P.fbw)
        sa_internal = min(sa_leaf, sa_stem)
        sa_internal = k_internal * sa_internal
        cp_leaf = leaf_biomass * (C_DRY_BIOMASS * (1.0 - P.fbw) + (P.fbw) * C_WATER)
        cp_stem = stem_biomass * (C_DRY_BIOMASS * (1.0 - P.fbw) + (P.fbw) * C_WATER)
        cp_stem = k_cyl_vol * cp_stem
        rstem = P.rstem_per_dbh * dbh
    else:
        sa_leaf = (elai + esai)
        frac_rad_abs_by_stem = sa_stem = sa_internal = cp_leaf = cp_stem = rstem = 0.0
    O.leaf_biomass, O.stem_biomass = leaf_biomass, stem_biomass
    P.dayl_factor = min(1.0, max(0.01, (P.dayl * P.dayl) / (P.max_dayl * P.max_dayl)))
    O.eff_porosity, O.h2osoi_liqvol, O.rresis, O.rootr, btran = root_moist_stress(P)
    # roughness length and displacement height (:902-946)
    displa, z0mv = P.displa, P.z0mv
    if M.z0param_method == 1:
        lt = min(elai + esai, TLSAI_CRIT)
        egvf = (1.0 - ALPHA_AERO * math.exp(-lt)) / (1.0 - ALPHA_AERO * math.exp(-TLSAI_CRIT))
        displa = egvf * displa
        z0mv = math.exp(egvf * math.log(z0mv) + (1.0 - egvf) * math.log(P.z0mg))
    elif M.z0param_method == 2:
        lt = max(1.e-5, elai + esai)
        displa = htop * (1.0 - (1.0 - math.exp(-(CD1_PARAM * lt) ** 0.5)) / (CD1_PARAM * lt) ** 0.5)
        lt = min(lt, P.z0v_LAImax)
        delt = 2.0
        U_ustar_ini = (P.z0v_Cs + P.z0v_Cr * lt * 0.5) ** (-0.5) * P.z0v_c * lt * 0.25
        U_ustar = U_ustar_ini
        while delt > 1.e-4:
            U_ustar_prev = U_ustar
            U_ustar = U_ustar_ini * math.exp(U_ustar_prev)
            delt = abs(U_ustar - U_ustar_prev)
        U_ustar = 4.0 * U_ustar / lt / P.z0v_c
        z0mv = htop * (1.0 - displa / htop) * math.exp(-VKC * U_ustar + math.log(P.z0v_cw) - 1.0 + P.z0v_cw ** (-1.0))
    else:
        raise ValueError("unknown z0param_method")
    z0hv = z0mv
    z0qv = z0mv
    O.displa, O.z0mv, O.z0hv, O.z0qv = displa, z0mv, z0hv, z0qv
    hgt_u = O.forc_hgt_u_patch = P.forc_hgt_u + z0mv + displa
    hgt_t = O.forc_hgt_t_patch = P.forc_hgt_t + z0hv + displa
    hgt_q = O.forc_hgt_q_patch = P.forc_hgt_q + z0qv + displa
    # initial conditions (:952-1015)
    emv, emg = P.emv, P.emg
    t_veg, t_stem, t_grnd, thm = P.t_veg, P.t_stem, P.t_grnd, P.thm
    air = emv * (1.0 + (1.0 - emv) * (1.0 - emg)) * P.forc_lwrad
    bir = -(2.0 - emv * (1.0 - emg)) * emv * SB
    cir = emv * emg * SB
    qsatl, el, qsatldT, _ = qsat(t_veg, P.forc_pbot)
    P.cair = P.forc_pco2
    P.oair = P.forc_po2
    nmozsgn = 0
    taf = (t_grnd + thm) / 2.0
    qaf = (P.forc_q + P.qg) / 2.0
    ur = max(M.wind_min, math.sqrt(P.forc_u * P.forc_u + P.forc_v * P.forc_v))
    dth = thm - taf
    dqh = P.forc_q - qaf
    delq = P.qg - qaf
    dthv = dth * (1.0 + 0.61 * P.forc_q) + 0.61 * P.forc_th * dqh
    zldis = hgt_u - displa
    if zldis < 0.0:
        raise ValueError("Forcing height is below canopy height")
    um, obu = monin_obuk_ini(M.zetamaxstable, ur, P.thv, dthv, zldis, z0mv)
    num_iter = 0
    tl_ini, ts_ini = t_veg, t_stem
    itlef = 0
    fm = None
    active = True
    del2 = 0.0
    O.phs = None
    while itlef <= M.itmax_canopy_fluxes and active:
        fv = friction_velocity(hgt_u, hgt_t, hgt_q, displa, z0mv, z0hv, z0qv, obu, itlef + 1, ur, um, fm)
        ustar, temp1, temp2, temp12m, temp22m, fm = fv.ustar, fv.temp1, fv.temp2, fv.temp12m, fv.temp22m, fv.fm
        O.vds, O.u10_clm, O.va, O.u10, O.fv = fv.vds, fv.u10_clm, fv.va, fv.u10, fv.fv
        tlbef = t_veg
        del2 = del_
        ram1 = 1.0 / (ustar * ustar / um)
        rah_above = 1.0 / (temp1 * ustar)
        raw_above = 1.0 / (temp2 * ustar)
        uaf = um * math.sqrt(1.0 / (ram1 * um))
        uuc = min(0.4, (0.03 * um / ustar))
        dleaf_patch = P.dleaf
        cf = M.cv / (math.sqrt(uaf) * math.sqrt(dleaf_patch))
        rb = 1.0 / (cf * uaf)
        rb1 = rb
        w = math.exp(-(elai + esai))
        csoilb = VKC / (M.a_coef * (P.z0mg * uaf / NU_PARAM) ** M.a_exp)
        ri = (GRAV * htop * (taf - t_grnd)) / (taf * uaf ** 2.00)
        if M.use_undercanopy_stability and (taf - t_grnd) > 0.0:
            ricsoilc = M.csoilc / (1.00 + ria * min(ri, 10.0))
            csoilcn = csoilb * w + ricsoilc * (1.0 - w)
        else:
            csoilcn = csoilb * w + M.csoilc * (1.0 - w)
        if M.use_biomass_heat_storage:
            rah_below = 1.0 / (csoilcn * uuc)
        else:
            rah_below = 1.0 / (csoilcn * uaf)
        raw_below = rah_below
        svpts = el
        eah = P.forc_pbot * qaf / 0.622
        rhaf = eah / svpts
        vpd = max((svpts - eah), 50.0) * 0.001
        # photosynthesis with plant hydraulic stress
        P.esat_tv, P.eair, P.rb, P.qsatl, P.qaf, P.t_veg = svpts, eah, rb, qsatl, qaf, t_veg
        if M.use_hydrstress:
            W = phs(P, M)
            P.vegwp, P.gs_mol = W.vegwp, W.gs_mol
            P.bsun_in, P.bsha_in = W.bsun, W.bsha
            btran = W.btran
            qflx_tran_veg = W.qflx_tran_veg
            rssun, rssha = W.rs[1], W.rs[2]
            O.phs = W
        else:                                                         # Photosynthesis for sunlit, then shaded leaves (:1143-1166)
            P.gs_mol_in = P.gs_mol_patch
            O.psn_sun = phs(P, M, 1, btran)
            P.gs_mol_in = O.psn_sun.gs_mol
            O.psn_sha = phs(P, M, 2, btran)
            P.gs_mol_patch = O.psn_sha.gs_mol
            rssun, rssha = O.psn_sun.rs, O.psn_sha.rs
        # fluxes and the leaf temperature update (:1176-1369)
        wta = 1.0 / rah_above
        wtl = sa_leaf / rb
        wtg = 1.0 / rah_below
        wtstem = sa_stem / (rstem + rb)
        wtshi = 1.0 / (wta + wtl + wtstem + wtg)
        wtl0 = wtl * wtshi
        wtg0 = wtg * wtshi
        wta0 = wta * wtshi
        wtstem0 = wtstem * wtshi
        wtga = wta0 + wtg0 + wtstem0
        wtal = wta0 + wtl0 + wtstem0
        lw_stem = sa_internal * emv * SB * _p4(t_stem)
        lw_leaf = sa_internal * emv * SB * _p4(t_veg)
        if P.fdry > 0.0:
            rppdry = P.fdry * rb * (P.laisun / (rb + rssun) + P.laisha / (rb + rssha)) / elai
        else:
            rppdry = 0.0
        efpot = P.forc_rho * ((elai + esai) / rb) * (qsatl - qaf)
        h2ocan = P.liqcan + P.snocan
        if M.use_hydrstress:
            if efpot > 0.0:
                if btran > btran0:
                    rpp = rppdry + P.fwet
                else:
                    rpp = P.fwet
                rpp = min(rpp, (qflx_tran_veg + h2ocan / dtime) / efpot)
            else:
                rpp = 1.0
        else:
            if efpot > 0.0:
                if btran > btran0:
                    qflx_tran_veg = efpot * rppdry
                    rpp = rppdry + P.fwet
                else:
                    rpp = P.fwet
                    qflx_tran_veg = 0.0
                rpp = min(rpp, (qflx_tran_veg + h2ocan / dtime) / efpot)
            else:
                rpp = 1.0
                qflx_tran_veg = 0.0
        wtaq = P.frac_veg_nosno / raw_above
        wtlq = P.frac_veg_nosno * (elai + esai) / rb * rpp
        snow_depth_c = M.z_dl
        fsno_dl = P.snow_depth / snow_depth_c
        elai_dl = M.lai_dl * (1.0 - min(fsno_dl, 1.0))
        rdl = (1.0 - math.exp(-elai_dl)) / (0.004 * uaf)
        if delq < 0.0:
            wtgq = P.frac_veg_nosno / (raw_below + rdl)
        else:
            if M.soil_resis_method == 0:
                wtgq = P.soilbeta * P.frac_veg_nosno / (raw_below + rdl)
            if M.soil_resis_method == 1:
                wtgq = P.frac_veg_nosno / (raw_below + P.soilresis)
        wtsqi = 1.0 / (wtaq + wtlq + wtgq)
        wtgq0 = wtgq * wtsqi
        wtlq0 = wtlq * wtsqi
        wtaq0 = wtaq * wtsqi
        wtgaq = wtaq0 + wtgq0
        wtalq = wtaq0 + wtlq0
        dc1 = P.forc_rho * CPAIR * wtl
        dc2 = HVAP * P.forc_rho * wtlq
        efsh = dc1 * (wtga * t_veg - wtg0 * t_grnd - wta0 * thm - wtstem0 * t_stem)
        eflx_sh_stem = P.forc_rho * CPAIR * wtstem * ((wta0 + wtg0 + wtl0) * t_stem - wtg0 * t_grnd - wta0 * thm - wtl0 * t_veg)
        efe = dc2 * (wtgaq * qsatl - wtgq0 * P.qg - wtaq0 * P.forc_q)
        erre = 0.0
        if efe * efeb < 0.0:
            efeold = efe
            efe = 0.1 * efeold
            erre = efe - efeold
        lw_grnd = (P.frac_sno * _p4(P.t_soisno[P.snl + 1]) + (1.0 - P.frac_sno - P.frac_h2osfc) * _p4(P.t_soisno[1])
                   + P.frac_h2osfc * _p4(P.t_h2osfc))
        dt_veg = (((1.0 - frac_rad_abs_by_stem) * (P.sabv + air + bir * _p4(t_veg) + cir * lw_grnd) - efsh - efe - lw_leaf + lw_stem
                   - (cp_leaf / dtime) * (t_veg - tl_ini))
                  / ((1.0 - frac_rad_abs_by_stem) * (-4.0 * bir * _p3(t_veg)) + 4.0 * sa_internal * emv * SB * _p3(t_veg)
                     + dc1 * wtga + dc2 * wtgaq * qsatldT + cp_leaf / dtime))
        t_veg = tlbef + dt_veg
        dels = dt_veg
        del_ = abs(dels)
        err = 0.0
        if del_ > delmax:
            dt_veg = delmax * dels / del_
            t_veg = tlbef + dt_veg
            err = ((1.0 - frac_rad_abs_by_stem) * (P.sabv + air + bir * _p3(tlbef) * (tlbef + 4.0 * dt_veg) + cir * lw_grnd)
                   - sa_internal * emv * SB * _p3(tlbef) * (tlbef + 4.0 * dt_veg) + lw_stem
                   - (efsh + dc1 * wtga * dt_veg) - (efe + dc2 * wtgaq * qsatldT * dt_veg) - (cp_leaf / dtime) * (t_veg - tl_ini))
        efpot = P.forc_rho * ((elai + esai) / rb) * (wtgaq * (qsatl + qsatldT * dt_veg) - wtgq0 * P.qg - wtaq0 * P.forc_q)
        qflx_evap_veg = rpp * efpot
        if not M.use_hydrstress:
            if efpot > 0.0 and btran > btran0:
                qflx_tran_veg = efpot * rppdry
            else:
                qflx_tran_veg = 0.0
        ecidif = max(0.0, qflx_evap_veg - qflx_tran_veg - h2ocan / dtime)
        qflx_evap_veg = min(qflx_evap_veg, qflx_tran_veg + h2ocan / dtime)
        eflx_sh_veg = efsh + dc1 * wtga * dt_veg + err + erre + HVAP * ecidif
        eflx_sh_stem = eflx_sh_stem + P.forc_rho * CPAIR * wtstem * (-wtl0 * dt_veg)
        lw_leaf = sa_internal * emv * SB * _p3(tlbef) * (tlbef + 4.0 * dt_veg)
        qsatl, el, qsatldT, _ = qsat(t_veg, P.forc_pbot)
        taf = wtg0 * t_grnd + wta0 * thm + wtl0 * t_veg + wtstem0 * t_stem
        qaf = wtlq0 * qsatl + wtgq0 * P.qg + P.forc_q * wtaq0
        dth = thm - taf
        dqh = P.forc_q - qaf
        delq = wtalq * P.qg - wtlq0 * qsatl - wtaq0 * P.forc_q
        tstar = temp1 * dth
        qstar = temp2 * dqh
        thvstar = tstar * (1.0 + 0.61 * P.forc_q) + 0.61 * P.forc_th * qstar
        zeta = zldis * VKC * GRAV * thvstar / (ustar * ustar * P.thv)
        if zeta >= 0.0:
            zeta = min(M.zetamaxstable, max(zeta, 0.01))
            um = max(ur, 0.1)
        else:
            zeta = max(-100.0, min(zeta, -0.01))
            if ustar * thvstar > 0.0:
                wc = 0.0
            else:
                wc = beta * (-GRAV * ustar * thvstar * zii / P.thv) ** 0.333
            um = math.sqrt(ur * ur + wc * wc)
        obu = zldis / zeta
        if obuold * obu < 0.0:
            nmozsgn = nmozsgn + 1
        if nmozsgn >= 4:
            obu = zldis / (-0.01)
        obuold = obu
        itlef = itlef + 1
        if itlef > itmin:
            dele = abs(efe - efeb)
            efeb = efe
            det = max(del_, del2)
            num_iter = itlef
            if det < dtmin and dele < dlemin:
                active = False
    # after the iteration (:1460-1660)
    lw_grnd = (P.frac_sno * _p4(P.t_soisno[P.snl + 1]) + (1.0 - P.frac_sno - P.frac_h2osfc) * _p4(P.t_soisno[1])
               + P.frac_h2osfc * _p4(P.t_h2osfc))
    err = ((1.0 - frac_rad_abs_by_stem) * (P.sabv + air + bir * _p3(tlbef) * (tlbef + 4.0 * dt_veg) + cir * lw_grnd)
           - lw_leaf + lw_stem - eflx_sh_veg - HVAP * qflx_evap_veg - ((t_veg - tl_ini) * cp_leaf / dtime))
    if M.use_biomass_heat_storage:
        if stem_biomass > 0.0:
            dt_stem = ((frac_rad_abs_by_stem * (P.sabv + air + bir * _p4(ts_ini) + cir * lw_grnd) - eflx_sh_stem + lw_leaf - lw_stem)
                       / (cp_stem / dtime - frac_rad_abs_by_stem * bir * 4.0 * _p3(ts_ini)))
        else:
            dt_stem = 0.0
        O.dhsdt_canopy = dt_stem * cp_stem / dtime + (t_veg - tl_ini) * cp_leaf / dtime
        t_stem = t_stem + dt_stem
    else:
        dt_stem = 0.0
    delt = wtal * t_grnd - wtl0 * t_veg - wta0 * thm - wtstem0 * t_stem
    O.taux = -P.forc_rho * P.forc_u / ram1
    O.tauy = -P.forc_rho * P.forc_v / ram1
    O.eflx_sh_grnd = CPAIR * P.forc_rho * wtg * delt
    delt_snow = wtal * P.t_soisno[P.snl + 1] - wtl0 * t_veg - wta0 * thm - wtstem0 * t_stem
    delt_soil = wtal * P.t_soisno[1] - wtl0 * t_veg - wta0 * thm - wtstem0 * t_stem
    delt_h2osfc = wtal * P.t_h2osfc - wtl0 * t_veg - wta0 * thm - wtstem0 * t_stem
    O.eflx_sh_snow = CPAIR * P.forc_rho * wtg * delt_snow
    O.eflx_sh_soil = CPAIR * P.forc_rho * wtg * delt_soil
    O.eflx_sh_h2osfc = CPAIR * P.forc_rho * wtg * delt_h2osfc
    O.qflx_evap_soi = P.forc_rho * wtgq * delq
    delq_snow = wtalq * P.qg_snow - wtlq0 * qsatl - wtaq0 * P.forc_q
    O.qflx_ev_snow = P.forc_rho * wtgq * delq_snow
    delq_soil = wtalq * P.qg_soil - wtlq0 * qsatl - wtaq0 * P.forc_q
    O.qflx_ev_soil = P.forc_rho * wtgq * delq_soil
    delq_h2osfc = wtalq * P.qg_h2osfc - wtlq0 * qsatl - wtaq0 * P.forc_q
    O.qflx_ev_h2osfc = P.forc_rho * wtgq * delq_h2osfc
    O.t_ref2m = thm + temp1 * dth * (1.0 / temp12m - 1.0 / temp1)
    O.t_ref2m_r = O.t_ref2m
    O.q_ref2m = P.forc_q + temp2 * dqh * (1.0 / temp22m - 1.0 / temp2)
    qsat_ref2m, e_ref2m, _, _ = qsat(O.t_ref2m, P.forc_pbot)
    O.rh_ref2m = min(100.0, O.q_ref2m / qsat_ref2m * 100.0)
    O.rh_ref2m_r = O.rh_ref2m
    O.vpd_ref2m = e_ref2m * (1.0 - O.rh_ref2m / 100.0)
    O.dlrad = ((1.0 - emv) * emg * P.forc_lwrad
               + emv * emg * SB * _p3(tlbef) * (tlbef + 4.0 * dt_veg) * (1.0 - frac_rad_abs_by_stem)
               + emv * emg * SB * _p3(ts_ini) * (ts_ini + 4.0 * dt_stem) * frac_rad_abs_by_stem)
    O.ulrad = ((1.0 - emg) * (1.0 - emv) * (1.0 - emv) * P.forc_lwrad
               + emv * (1.0 + (1.0 - emg) * (1.0 - emv)) * SB * _p3(tlbef) * (tlbef + 4.0 * dt_veg) * (1.0 - frac_rad_abs_by_stem)
               + emv * (1.0 + (1.0 - emg) * (1.0 - emv)) * SB * _p3(ts_ini) * (ts_ini + 4.0 * dt_stem) * frac_rad_abs_by_stem
               + emg * (1.0 - emv) * SB * lw_grnd)
    O.t_skin = emv * t_veg + (1.0 - emv) * math.sqrt(math.sqrt(lw_grnd))
    O.cgrnds = P.cgrnds + CPAIR * P.forc_rho * wtg * wtal
    O.cgrndl = P.cgrndl + P.forc_rho * wtgq * wtalq * P.dqgdT
    O.cgrnd = O.cgrnds + O.cgrndl * P.htvp
    snocan, liqcan = P.snocan, P.liqcan
    snocan_baseline = snocan
    if t_veg > TKFRZ:
        if (qflx_evap_veg - qflx_tran_veg) * dtime > liqcan:
            snocan = max(0.0, snocan + liqcan + (qflx_tran_veg - qflx_evap_veg) * dtime)
        liqcan = max(0.0, liqcan + (qflx_tran_veg - qflx_evap_veg) * dtime)
    elif t_veg <= TKFRZ:
        if (qflx_evap_veg - qflx_tran_veg) * dtime > snocan:
            liqcan = liqcan + snocan + (qflx_tran_veg - qflx_evap_veg) * dtime
        snocan = max(0.0, snocan + (qflx_tran_veg - qflx_evap_veg) * dtime)
    if abs(snocan) < 1.e-10 * abs(snocan_baseline):                  # truncate_small_values (NumericsMod.F90:53-99)
        snocan = 0.0
    O.snocan, O.liqcan = snocan, liqcan
    if M.use_hydrstress:
        W = O.phs
        psn, psn_wc, psn_wj, psn_wp, gs_ss = W.psn, W.psn_wc, W.psn_wj, W.psn_wp, W.gs_mol
    else:
        A, B = O.psn_sun, O.psn_sha
        psn, psn_wc, psn_wj, psn_wp = {1: A.psn, 2: B.psn}, {1: A.psn_wc, 2: B.psn_wc}, {1: A.psn_wj, 2: B.psn_wj}, {1: A.psn_wp, 2: B.psn_wp}
        gs_ss = {1: A.gs_mol_phase, 2: B.gs_mol_phase}
    O.fpsn = psn[1] * P.laisun + psn[2] * P.laisha                    # PhotosynthesisTotal (PhotosynthesisMod.F90:2065-2151)
    O.fpsn_wc = psn_wc[1] * P.laisun + psn_wc[2] * P.laisha
    O.fpsn_wj = psn_wj[1] * P.laisun + psn_wj[2] * P.laisha
    O.fpsn_wp = psn_wp[1] * P.laisun + psn_wp[2] * P.laisha
    if P.near_local_noon and O.fpsn > 0.0:
        gs = 1.e-6 * (P.laisun * gs_ss[1] + P.laisha * gs_ss[2])
        O.iwue_ln = O.fpsn / gs if gs > 0.0 else SPVAL
    else:
        O.iwue_ln = SPVAL
    O.t_veg, O.t_stem, O.btran, O.qflx_tran_veg, O.qflx_evap_veg = t_veg, t_stem, btran, qflx_tran_veg, qflx_evap_veg
    O.eflx_sh_veg, O.eflx_sh_stem, O.num_iter, O.err = eflx_sh_veg, eflx_sh_stem, num_iter, err
    O.ram1, O.rb1, O.rah1, O.raw1, O.rah2, O.raw2 = ram1, rb1, rah_above, raw_above, rah_below, raw_below
    O.ustar, O.um, O.uaf, O.taf, O.qaf, O.obu, O.zeta, O.vpd, O.rh_af = ustar, um, uaf, taf, qaf, obu, zeta, vpd, rhaf
    O.dleaf_patch = dleaf_patch
    return O


# ------------------------------------------------------------------------------------------------------------------------------
# BareGroundFluxes for one patch without exposed vegetation (BareGroundFluxesMod.F90:63-579; the human-stress indices are pinned
# separately, tests/test_oracle_canopy.py)
# ------------------------------------------------------------------------------------------------------------------------------
RGAS_SHR = 6.02214e26 * 1.38065e-23
BETA_PARAM = 7.2
MEIER_PARAM3 = 70.0


def dewpoint(e, t):
    """BareGroundFluxesMod.F90:549-579"""
    if t < TKFRZ:
        d = 273.86 * math.log(e / 611.21) / (22.587 - math.log(e / 611.21))
    else:
        d = 243.04 * math.log(e / 610.94) / (17.625 - math.log(e / 610.94))
    return d + TKFRZ


def bare_ground_fluxes_patch(P, M):
    niters = 3
    O = SimpleNamespace()
    O.btran = 0.0
    O.t_veg = P.forc_t
    cf_bare = P.forc_pbot / (RGAS_SHR * 0.001 * P.thm) * 1.e06
    O.rssun = 1.0 / 1.e15 * cf_bare
    O.rssha = 1.0 / 1.e15 * cf_bare
    displa = 0.0
    ur = max(M.wind_min, math.sqrt(P.forc_u * P.forc_u + P.forc_v * P.forc_v))
    dth = P.thm - P.t_grnd
    dqh = P.forc_q - P.qg
    dthv = dth * (1.0 + 0.61 * P.forc_q) + 0.61 * P.forc_th * dqh
    zldis = P.forc_hgt_u_patch
    z0mg, z0hg, z0qg = P.z0mg, P.z0hg, P.z0qg
    um, obu = monin_obuk_ini(M.zetamaxstable, ur, P.thv, dthv, zldis, z0mg)
    hgt_u, hgt_t, hgt_q = P.forc_hgt_u_patch, P.forc_hgt_t_patch, P.forc_hgt_q_patch
    fm = None
    for it in range(1, niters + 1):
        fv = friction_velocity(hgt_u, hgt_t, hgt_q, displa, z0mg, z0hg, z0qg, obu, it, ur, um, fm)
        ustar, temp1, temp2, temp12m, temp22m, fm = fv.ustar, fv.temp1, fv.temp2, fv.temp12m, fv.temp22m, fv.fm
        tstar = temp1 * dth
        qstar = temp2 * dqh
        if M.z0param_method == 1:
            z0hg = z0mg / math.exp(M.a_coef * (ustar * z0mg / NU_PARAM) ** M.a_exp)
        elif M.z0param_method == 2:
            z0hg = MEIER_PARAM3 * NU_PARAM / ustar * math.exp(-BETA_PARAM * ustar ** (0.5) * (abs(tstar)) ** (0.25))
        z0qg = z0hg
        hgt_u = P.forc_hgt_u + z0mg + displa
        hgt_t = P.forc_hgt_t + z0hg + displa
        hgt_q = P.forc_hgt_q + z0qg + displa
        thvstar = tstar * (1.0 + 0.61 * P.forc_q) + 0.61 * P.forc_th * qstar
        zeta = zldis * VKC * GRAV * thvstar / (ustar * ustar * P.thv)
        if zeta >= 0.0:
            zeta = min(M.zetamaxstable, max(zeta, 0.01))
            um = max(ur, 0.1)
        else:
            zeta = max(-100.0, min(zeta, -0.01))
            wc = P.beta * (-GRAV * ustar * thvstar * P.zii / P.thv) ** 0.333
            um = math.sqrt(ur * ur + wc * wc)
        obu = zldis / zeta
    O.num_iter = niters
    O.vds, O.u10, O.u10_clm, O.va, O.fv = fv.vds, fv.u10, fv.u10_clm, fv.va, fv.fv
    ram = 1.0 / (ustar * ustar / um)
    rah = 1.0 / (temp1 * ustar)
    raw = 1.0 / (temp2 * ustar)
    raih = P.forc_rho * CPAIR / rah
    _, forc_esat, _, _ = qsat(P.forc_t, P.forc_pbot)
    forc_e = max((P.forc_q * P.forc_pbot) / (P.forc_q + 0.622), 0.01 * forc_esat)
    forc_dewpoint = dewpoint(forc_e, P.t_grnd)
    raiw = None
    if dqh > 0.0:
        raiw = 0.0 if P.t_grnd > forc_dewpoint else P.forc_rho / (raw)
    else:
        if M.soil_resis_method == 0:
            raiw = 0.0 if P.t_grnd > forc_dewpoint else P.soilbeta * P.forc_rho / (raw)
        if M.soil_resis_method == 1:
            raiw = P.forc_rho / (raw + P.soilresis)
    O.ram1 = ram
    O.cgrnds = raih
    O.cgrndl = raiw * P.dqgdT
    O.cgrnd = O.cgrnds + P.htvp * O.cgrndl
    O.taux = -P.forc_rho * P.forc_u / ram
    O.tauy = -P.forc_rho * P.forc_v / ram
    O.eflx_sh_grnd = -raih * dth
    O.eflx_sh_tot = O.eflx_sh_grnd
    O.eflx_sh_snow = -raih * (P.thm - P.t_soisno[P.snl + 1])
    O.eflx_sh_soil = -raih * (P.thm - P.t_soisno[1])
    O.eflx_sh_h2osfc = -raih * (P.thm - P.t_h2osfc)
    O.qflx_tran_veg = 0.0
    O.qflx_evap_veg = 0.0
    O.qflx_evap_soi = -raiw * dqh
    O.qflx_evap_tot_patch = O.qflx_evap_soi
    O.qflx_ev_snow = -raiw * (P.forc_q - P.qg_snow)
    O.qflx_ev_soil = -raiw * (P.forc_q - P.qg_soil)
    O.qflx_ev_h2osfc = -raiw * (P.forc_q - P.qg_h2osfc)
    O.t_ref2m = P.thm + temp1 * dth * (1.0 / temp12m - 1.0 / temp1)
    O.q_ref2m = P.forc_q + temp2 * dqh * (1.0 / temp22m - 1.0 / temp2)
    qsat_ref2m, _, _, _ = qsat(O.t_ref2m, P.forc_pbot)
    O.rh_ref2m = min(100.0, O.q_ref2m / qsat_ref2m * 100.0)
    O.kbm1 = math.log(z0mg / z0hg)
    O.z0mg_p, O.z0hg_p, O.z0qg_p = z0mg, z0hg, z0qg
    O.forc_hgt_u_patch, O.forc_hgt_t_patch, O.forc_hgt_q_patch = hgt_u, hgt_t, hgt_q
    O.um, O.obu, O.zeta, O.ustar = um, obu, zeta, ustar
    for k in ("displa", "z0mv", "z0hv", "z0qv", "dlrad", "ulrad", "dhsdt_canopy", "eflx_sh_stem"):
        setattr(O, k, 0.0)
    return O
