"""Independent restatement of the snow routines of HydrologyNoDrainage in plain Python, written from the Fortran
(src/biogeophys/SnowHydrologyMod.F90, AerosolMod.F90, SnowCoverFractionSwensonLawrence2012Mod.F90), NOT from oracle/oracle_snow.c.
Test infrastructure: it pins the C oracle (tests/test_oracle_snow.py).  One column at a time, Fortran level indices kept
(a `Lev` is a list addressed with the Fortran index), so the index arithmetic can be read against the source line by line."""
import math

NLEVSNO = 12
DENICE, DENH2O, TFRZ, CPICE, CPLIQ, HFUS = 917.0, 1000.0, 273.15, 2.11727e3, 4.188e3, 3.337e5
AER = ("bcphi", "bcpho", "ocphi", "ocpho", "dst1", "dst2", "dst3", "dst4")


class Lev:
    """1-D array with a Fortran lower bound"""

    def __init__(self, lo, values):
        self.lo, self.v = lo, [x for x in values]

    def __getitem__(self, j):
        assert j - self.lo >= 0
        return self.v[j - self.lo]

    def __setitem__(self, j, x):
        assert j - self.lo >= 0
        self.v[j - self.lo] = x


def column(S, c, names_sno, names_snosoi, scalars):
    """pull column c (0-based) out of the level-major state arrays"""
    col = {k: Lev(-NLEVSNO + 1, S[k][:, c].tolist()) for k in names_sno + names_snosoi}
    col.update({k: (int(S[k][c]) if S[k].dtype.kind == "i" else float(S[k][c])) for k in scalars})
    return col


def dz_limits(prm):
    """InitSnowLayers :2985-3002"""
    dzmin, dzmax_l, dzmax_u = Lev(1, [0.0] * NLEVSNO), Lev(1, [0.0] * NLEVSNO), Lev(1, [0.0] * NLEVSNO)
    dzmin[1], dzmax_l[1], dzmax_u[1] = prm.snow_dzmin_1, prm.snow_dzmax_l_1, prm.snow_dzmax_u_1
    dzmin[2], dzmax_l[2], dzmax_u[2] = prm.snow_dzmin_2, prm.snow_dzmax_l_2, prm.snow_dzmax_u_2
    for j in range(3, NLEVSNO + 1):
        dzmin[j] = dzmax_u[j - 1] * 0.5
        dzmax_u[j] = 2.0 * dzmax_u[j - 1] + 0.01
        dzmax_l[j] = dzmax_u[j] + dzmax_l[j - 1]
        if j == NLEVSNO:
            dzmax_u[j] = dzmax_l[j] = float.fromhex("0x1.fffffffffffffp+1023")
    return dzmin, dzmax_l, dzmax_u


# ---------------------------------------------------------------------------------------------------------------------
def snow_water_column(prm, col, forc_aer):
    """SnowWater :1015-1165 for one snow column; col holds snl, frac_sno_eff, the five top-layer fluxes, int_snow, qflx_snow_drain and
    the Lev arrays h2osoi_ice, h2osoi_liq, dz, mss_*; returns qflx_snow_percolation (Lev) and qflx_rain_plus_snomelt"""
    dt, snl, fse = prm.dtime, col["snl"], col["frac_sno_eff"]
    ice, liq, dz = col["h2osoi_ice"], col["h2osoi_liq"], col["dz"]
    top = snl + 1
    # UpdateState_TopLayerFluxes :1210-1251
    ice0, liq0 = ice[top], liq[top]
    ice[top] = ice[top] + fse * (col["qflx_soliddew_to_top_layer"] - col["qflx_solidevap_from_top_layer"]) * dt
    liq[top] = liq[top] + fse * (col["qflx_liq_grnd"] + col["qflx_liqdew_to_top_layer"] - col["qflx_liqevap_from_top_layer"]) * dt
    if abs(ice[top]) < 1.e-12 * abs(ice0):
        ice[top] = 0.0
    if abs(liq[top]) < 1.e-12 * abs(liq0):
        liq[top] = 0.0
    if ice[top] < 0.0 or liq[top] < 0.0:
        raise ArithmeticError("significantly negative")
    # BulkFlux_SnowPercolation :1331-1380
    levels = range(snl + 1, 1)
    vol_ice, eff_por, vol_liq = {}, {}, {}
    for j in levels:
        vol_ice[j] = min(1.0, ice[j] / (dz[j] * fse * DENICE))
        eff_por[j] = 1.0 - vol_ice[j]
        vol_liq[j] = min(eff_por[j], liq[j] / (dz[j] * fse * DENH2O))
    perc = Lev(-NLEVSNO + 1, [1.0e36] * NLEVSNO)
    for j in levels:
        if j <= -1:
            if eff_por[j] < prm.wimp or eff_por[j + 1] < prm.wimp:
                q = 0.0
            else:
                q = max(0.0, (vol_liq[j] - prm.ssi * eff_por[j]) * dz[j] * fse)
                q = min(q, (1.0 - vol_ice[j + 1] - vol_liq[j + 1]) * dz[j + 1] * fse)
        else:
            q = max(0.0, (vol_liq[j] - prm.ssi * eff_por[j]) * dz[j] * fse)
        perc[j] = (q * 1000.0) / dt
    # UpdateState_SnowPercolation :1483-1494
    for j in levels:
        if j >= snl + 2:
            liq[j] = liq[j] + perc[j - 1] * dt
        liq[j] = liq[j] - perc[j] * dt
    # CalcAndApplyAerosolFluxes :1563-1700
    scv = {"bcphi": prm.scvng_fct_mlt_bcphi, "bcpho": prm.scvng_fct_mlt_bcpho, "ocphi": 0.20, "ocpho": 0.03, "dst1": prm.scvng_fct_mlt_dst1,
           "dst2": prm.scvng_fct_mlt_dst2, "dst3": prm.scvng_fct_mlt_dst3, "dst4": prm.scvng_fct_mlt_dst4}
    qin = {a: 0.0 for a in AER}
    for j in levels:
        for a in AER:
            col["mss_" + a][j] = col["mss_" + a][j] + qin[a] * dt
        mss_liqice = liq[j] + ice[j]
        if mss_liqice < 1e-30:
            mss_liqice = 1e-30
        for a in AER:
            m = col["mss_" + a]
            qout = perc[j] * prm.scvng_fct_mlt_sf * scv[a] * (m[j] / mss_liqice)
            if qout * dt > m[j]:
                qout = m[j] / dt
                m[j] = 0.0
            else:
                m[j] = m[j] - qout * dt
            qin[a] = qout
    # AerosolFluxes, AerosolMod.F90:728-750 and :785-798 (forc_aer: the 14 deposition fluxes of the column's gridcell, 1-based)
    fa = Lev(1, forc_aer if prm.snicar_use_aerosol else [0.0] * 14)
    dep = {"bcphi": (fa[1] + fa[3]) * dt, "bcpho": fa[2] * dt, "ocphi": (fa[4] + fa[6]) * dt, "ocpho": fa[5] * dt,
           "dst1": (fa[8] + fa[7]) * dt, "dst2": (fa[10] + fa[9]) * dt, "dst3": (fa[12] + fa[11]) * dt, "dst4": (fa[14] + fa[13]) * dt}
    for a in AER:
        col["mss_" + a][top] = col["mss_" + a][top] + dep[a]
    # PostPercolation_AdjustLayerThicknesses :1744
    for j in levels:
        dz[j] = max(dz[j], liq[j] / DENH2O + ice[j] / DENICE)
    # BulkDiag_SnowWaterAccumulatedSnow :1796-1800, SumFlux_AddSnowPercolation :1853-1857
    col["int_snow"] = col["int_snow"] + fse * (col["qflx_soliddew_to_top_layer"] + col["qflx_liqdew_to_top_layer"] + col["qflx_liq_grnd"]) * dt
    col["qflx_snow_drain"] = col["qflx_snow_drain"] + perc[0]
    col["qflx_rain_plus_snomelt"] = perc[0] + (1.0 - fse) * col["qflx_liq_grnd"]
    col["qflx_snow_percolation"] = perc


# ---------------------------------------------------------------------------------------------------------------------
def combo(dz, wliq, wice, t, dz2, wliq2, wice2, t2):
    """Combo :3902-3946; returns the combined (dz, wliq, wice, t)"""
    dzc = dz + dz2
    wicec = wice + wice2
    wliqc = wliq + wliq2
    h = (CPICE * wice + CPLIQ * wliq) * (t - TFRZ) + HFUS * wliq
    h2 = (CPICE * wice2 + CPLIQ * wliq2) * (t2 - TFRZ) + HFUS * wliq2
    hc = h + h2
    tc = TFRZ + (hc - HFUS * wliqc) / (CPICE * wicec + CPLIQ * wliqc)
    return dzc, wliqc, wicec, tc


def snow_compaction_column(prm, col, forc_wind):
    """SnowCompaction :1947-2077"""
    c3, c4, c5 = 2.777e-6, 0.04, 2.0
    dt, snl, frac_sno = prm.dtime, col["snl"], col["frac_sno_eff"]
    ice, liq, dz, t = col["h2osoi_ice"], col["h2osoi_liq"], col["dz"], col["t_soisno"]
    burden, zpseudo, mobile = 0.0, 0.0, True
    for j in range(snl + 1, 1):
        wx = ice[j] + liq[j]
        void = 1.0 - (ice[j] / DENICE + liq[j] / DENH2O) / (frac_sno * dz[j])
        if void > 0.001 and ice[j] > 0.1:
            bi = ice[j] / (frac_sno * dz[j])
            fi = ice[j] / wx
            td = TFRZ - t[j]
            dexpf = math.exp(-c4 * td)
            ddz1 = -c3 * dexpf
            if bi > prm.upplim_destruct_metamorph:
                ddz1 = ddz1 * math.exp(-46.0e-3 * (bi - prm.upplim_destruct_metamorph))
            if liq[j] > 0.01 * dz[j] * frac_sno:
                ddz1 = ddz1 * c5
            if prm.snow_overburden_compaction_method == 1:            # Anderson1976 :3789
                ddz2 = -(burden + wx / 2.0) * math.exp(-prm.overburden_compress_Tfactor * td - 23.e-3 * bi) / prm.eta0_anderson
            else:                                                     # Vionnet2012 :3825-3830
                f1 = 1.0 / (1.0 + 60.0 * liq[j] / (DENH2O * dz[j]))
                eta = f1 * 4.0 * (bi / prm.ceta) * math.exp(0.1 * td + 0.023 * bi) * prm.eta0_vionnet
                ddz2 = -(burden + wx / 2.0) / eta
            if col["imelt"][j] == 1:
                if prm.use_subgrid_fluxes:
                    ddz3 = max(0.0, min(1.0, (col["swe_old"][j] - wx) / wx))
                    if (col["swe_old"][j] - wx) > 0.0:
                        wsum = 0.0
                        for jj in range(snl + 1, 1):
                            wsum = wsum + (liq[jj] + ice[jj])
                        smr = min(1.0, wsum / min(col["int_snow"], prm.int_snow_max))          # FracSnowDuringMelt :263-266
                        fsno_melt = 1.0 - (math.acos(min(1.0, 2.0 * smr - 1.0)) / math.pi) ** col["n_melt"]
                        if fsno_melt + col["frac_h2osfc"] > 1.0:
                            fsno_melt = 1.0 - col["frac_h2osfc"]
                        ddz3 = ddz3 - max(0.0, (fsno_melt - frac_sno) / frac_sno)
                    ddz3 = -1.0 / dt * ddz3
                else:
                    ddz3 = -1.0 / dt * max(0.0, (col["frac_iceold"][j] - fi) / col["frac_iceold"][j])
            else:
                ddz3 = 0.0
            ddz4 = 0.0
            if prm.wind_dependent_snow_density and mobile:            # WindDriftCompaction :3872-3897
                frho = 1.25 - 0.0042 * (max(50.0, bi) - 50.0)
                mo = 0.34 * (-0.583 * prm.drift_gs - 0.833 * 1.0 + 0.833) + 0.66 * frho
                si = -2.868 * math.exp(-0.085 * forc_wind) + 1.0 + mo
                if si > 0.0:
                    si = min(si, 3.25)
                    zpseudo = zpseudo + 0.5 * dz[j] * (3.25 - si)
                    gamma_drift = si * math.exp(-zpseudo / 0.1)
                    ddz4 = -max(0.0, prm.rho_max - bi) * (gamma_drift / prm.tau_ref)
                    zpseudo = zpseudo + 0.5 * dz[j] * (3.25 - si)
                else:
                    mobile = False
            pdzdtc = ddz1 + ddz2 + ddz3 + ddz4
            dz[j] = max(dz[j] * (1.0 + pdzdtc * dt), (ice[j] / DENICE + liq[j] / DENH2O) / frac_sno)
        else:
            mobile = False
        burden = burden + wx


def combine_snow_layers_column(prm, col, dzmin):
    """CombineSnowLayers :2186-2503 for a non-lake column (dzminloc = dzmin)"""
    dt = prm.dtime
    ice, liq, dz, t, rds = col["h2osoi_ice"], col["h2osoi_liq"], col["dz"], col["t_soisno"], col["snw_rds"]
    mss = [col["mss_" + a] for a in AER]
    soil = col["lun_itype"] in (1, 2)
    col["qflx_sl_top_soil"] = 0.0
    msn_old = col["snl"]
    for j in range(msn_old + 1, 1):
        if ice[j] <= 0.01:
            if j < 0 or soil:
                liq[j + 1] = liq[j + 1] + liq[j]
                ice[j + 1] = ice[j + 1] + ice[j]
            if j < 0:
                dz[j + 1] = dz[j + 1] + dz[j]
                for m in mss:
                    m[j + 1] = m[j + 1] + m[j]
            if j == 0:
                col["qflx_sl_top_soil"] = (liq[j] + ice[j]) / dt
            if j > col["snl"] + 1 and col["snl"] < -1:
                for i in range(j, col["snl"] + 1, -1):               # do i = j, snl+2, -1
                    for arr in [liq, ice, t, rds, dz] + mss:
                        arr[i] = arr[i - 1]
            col["snl"] = col["snl"] + 1
    snow_depth = h2osno_total = zwice = zwliq = 0.0
    for j in range(-NLEVSNO + 1, 1):
        if j >= col["snl"] + 1:
            zwice = zwice + ice[j]
            zwliq = zwliq + liq[j]
            snow_depth = snow_depth + dz[j]
            h2osno_total = h2osno_total + ice[j] + liq[j]
    fse = col["frac_sno_eff"]
    if snow_depth > 0.0:
        if fse * snow_depth < dzmin[1] or h2osno_total / (fse * snow_depth) < 50.0:
            col["h2osno_no_layers"] = zwice
            if soil:
                liq[1] = liq[1] + zwliq
            col["snl"] = 0
            h2osno_total = col["h2osno_no_layers"]
            for m in mss:
                for j in range(-NLEVSNO + 1, 1):
                    m[j] = 0.0
            if col["h2osno_no_layers"] <= 0.0:
                snow_depth = 0.0
    if h2osno_total <= 0.0:
        snow_depth = 0.0
        col["frac_sno"] = col["frac_sno_eff"] = col["int_snow"] = 0.0
    col["snow_depth"] = snow_depth
    fse = col["frac_sno_eff"]
    if col["snl"] < -1:
        msn_old = col["snl"]
        mssi = 1
        for i in range(msn_old + 1, 1):
            if fse * dz[i] < dzmin[mssi] or (ice[i] + liq[i]) / (fse * dz[i]) < 50.0:
                if i == col["snl"] + 1:
                    neibor = i + 1
                elif i == 0:
                    neibor = i - 1
                else:
                    neibor = i + 1
                    if (dz[i - 1] + dz[i]) < (dz[i + 1] + dz[i]):
                        neibor = i - 1
                j, l = (neibor, i) if neibor > i else (i, neibor)
                for m in mss:
                    m[j] = m[j] + m[l]
                rds[j] = (rds[j] * (liq[j] + ice[j]) + rds[l] * (liq[l] + ice[l])) / (liq[j] + ice[j] + liq[l] + ice[l])
                dz[j], liq[j], ice[j], t[j] = combo(dz[j], liq[j], ice[j], t[j], dz[l], liq[l], ice[l], t[l])
                if j - 1 > col["snl"] + 1:
                    for k in range(j - 1, col["snl"] + 1, -1):       # do k = j-1, snl+2, -1
                        for arr in [ice, liq, t, rds, dz] + mss:
                            arr[k] = arr[k - 1]
                col["snl"] = col["snl"] + 1
                if col["snl"] >= -1:
                    break
            else:
                mssi = mssi + 1


def divide_snow_layers_column(prm, col, dzmax_l, dzmax_u):
    """DivideSnowLayers :2620-2879, is_lake = .false."""
    snl = col["snl"]
    fs = col["frac_sno_eff"]
    n0 = abs(snl)
    z = lambda: Lev(1, [0.0] * NLEVSNO)
    dzsno, swice, swliq, tsno, rds = z(), z(), z(), z(), z()
    m = {a: z() for a in AER}
    for j in range(1, n0 + 1):
        dzsno[j] = fs * col["dz"][j + snl]
        swice[j], swliq[j] = col["h2osoi_ice"][j + snl], col["h2osoi_liq"][j + snl]
        tsno[j], rds[j] = col["t_soisno"][j + snl], col["snw_rds"][j + snl]
        for a in AER:
            m[a][j] = col["mss_" + a][j + snl]
    msno = n0
    k = 1
    while k <= msno and k < NLEVSNO:
        if k == msno:
            if dzsno[k] > dzmax_l[k]:
                msno = msno + 1
                dzsno[k] = dzsno[k] / 2.0
                dzsno[k + 1] = dzsno[k]
                swice[k] = swice[k] / 2.0
                swice[k + 1] = swice[k]
                swliq[k] = swliq[k] / 2.0
                swliq[k + 1] = swliq[k]
                if k == 1:
                    tsno[k + 1] = tsno[k]
                else:
                    dtdz = (tsno[k - 1] - tsno[k]) / ((dzsno[k - 1] + 2 * dzsno[k]) / 2.0)
                    tsno[k + 1] = tsno[k] - dtdz * dzsno[k] / 2.0
                    if tsno[k + 1] >= TFRZ:
                        tsno[k + 1] = tsno[k]
                    else:
                        tsno[k] = tsno[k] + dtdz * dzsno[k] / 2.0
                for a in AER:
                    m[a][k] = m[a][k] / 2.0
                    m[a][k + 1] = m[a][k]
                rds[k + 1] = rds[k]
        if k < msno:
            if dzsno[k] > dzmax_u[k]:
                drr = dzsno[k] - dzmax_u[k] - 0.0
                propor = drr / dzsno[k]
                zwice, zwliq = propor * swice[k], propor * swliq[k]
                zm = {a: propor * m[a][k] for a in AER}
                propor = (dzmax_u[k] + 0.0) / dzsno[k]
                swice[k], swliq[k] = propor * swice[k], propor * swliq[k]
                for a in AER:
                    m[a][k] = propor * m[a][k]
                dzsno[k] = dzmax_u[k] + 0.0
                for a in AER:
                    m[a][k + 1] = m[a][k + 1] + zm[a]
                swtot, zwtot = swliq[k + 1] + swice[k + 1], zwliq + zwice              # MassWeightedSnowRadius :3966-3971
                r = (rds[k + 1] * swtot + rds[k] * zwtot) / (swtot + zwtot)
                rds[k + 1] = 1500.0 if r > 1500.0 else (prm.snw_rds_min if r < prm.snw_rds_min else r)
                dzsno[k + 1], swliq[k + 1], swice[k + 1], tsno[k + 1] = combo(dzsno[k + 1], swliq[k + 1], swice[k + 1], tsno[k + 1],
                                                                              drr, zwliq, zwice, tsno[k])
        k = k + 1
    col["snl"] = snl = -msno
    for j in range(snl + 1, 1):
        col["dz"][j] = dzsno[j - snl] / fs
        col["h2osoi_ice"][j], col["h2osoi_liq"][j] = swice[j - snl], swliq[j - snl]
        col["t_soisno"][j], col["snw_rds"][j] = tsno[j - snl], rds[j - snl]
        for a in AER:
            col["mss_" + a][j] = m[a][j - snl]


def finish_column(col, zi):
    """node depths :2883-2891 and ZeroEmptySnowLayers :2935-2947; zi is a Lev with lower bound -nlevsno"""
    snl = col["snl"]
    for j in range(0, -NLEVSNO, -1):
        if j >= snl + 1:
            col["z"][j] = zi[j] - 0.5 * col["dz"][j]
            zi[j - 1] = zi[j] - col["dz"][j]
    for j in range(-NLEVSNO + 1, 1):
        if j <= snl and snl > -NLEVSNO:
            for k in ("h2osoi_ice", "h2osoi_liq", "t_soisno", "dz", "z"):
                col[k][j] = 0.0
            zi[j - 1] = 0.0


def snow_layers_column(prm, col, zi, forc_wind):
    dzmin, dzmax_l, dzmax_u = dz_limits(prm)
    snow_compaction_column(prm, col, forc_wind)
    combine_snow_layers_column(prm, col, dzmin)
    if col["snl"] < 0:
        divide_snow_layers_column(prm, col, dzmax_l, dzmax_u)
    finish_column(col, zi)
