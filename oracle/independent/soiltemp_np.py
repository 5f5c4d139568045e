"""Second, independent restatement of SoilTemperature in NumPy + SciPy LAPACK (dgbsv).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Written from the Fortran alone, as whole-array operations over all non-lake
columns at once (the C oracle loops column by column and streams a band LU; this one materialises the full banded
matrix per column exactly as SetMatrix / AssembleMatrixFromSubmatrices describe it and hands it to LAPACK's dgbsv
through scipy, like BandDiagonal does).  tests/test_oracle_independent.py holds oracle/oracle_soiltemp.c to it.

Reference: src/biogeophys/SoilTemperatureMod.F90
    SoilTemperature :92-599          SoilThermProp :602-901              PhaseChangeH2osfc :904-1130
    Phasechange :1133-1540           ComputeGroundHeatFluxAndDeriv :1543-1796
    ComputeHeatDiffFluxAndFactor :1799-1910    SetRHSVec* :1913-2353    SetMatrix* :2356-2926
  src/biogeophys/BandDiagonalMod.F90:167-219, src/biogeophys/WaterStateType.F90:856-897 (CalculateTotalH2osno)
Scope = this repo's hot path: non-urban columns (istsoil, istcrop, istice, istwet landunits), use_excess_ice off.

Un-suffixed Fortran literals are REAL(4) constants promoted to double: R4() below.
Array convention: S[name] has shape (levels, n); level j of a (-nlevsno+1:nlevgrnd) field is row j + 11, of a
(-nlevsno:nlevgrnd) field row j + 12, of a (1:...) field row j - 1, of a snow field (-nlevsno+1:0) row j + 11.
"""
import numpy as np
from scipy.linalg import lapack

NLEVSNO, NLEVGRND, NLEVSOI = 12, 25, 20
NL = NLEVSNO + NLEVGRND
TFRZ, DENH2O, DENICE = 273.15, 1.000e3, 0.917e3
CPLIQ, CPICE, HFUS, SB, GRAV = 4.188e3, 2.11727e3, 3.337e5, 5.67e-8, 9.80616
TKAIR, TKICE, TKWAT = 0.023, 2.290, 0.57
CAPR, CNFAC, THK_BEDROCK, CSOL_BEDROCK, THIN = 0.34, 0.5, 3.0, 2.0e6, 1.0e-6
ISTSOIL, ISTCROP, ISTICE, ISTWET = 1, 2, 4, 6


def R4(x):
    return float(np.float32(x))


def soiltemperature(S, filter_nolakep, filter_nolakec, dtime=1800.0, snow_method=2, snow_glc_method=2):
    """In-place update of every OUT / INOUT field of the SoilTemperature group for the given filters (1-based)."""
    with np.errstate(all="ignore"):          # whole-array code evaluates every branch; unused lanes may be inf / nan
        _soiltemperature(S, filter_nolakep, filter_nolakec, dtime, snow_method, snow_glc_method)


def _soiltemperature(S, filter_nolakep, filter_nolakec, dtime, snow_method, snow_glc_method):
    cols = np.asarray(filter_nolakec, dtype=np.int64) - 1
    pats = np.asarray(filter_nolakep, dtype=np.int64) - 1
    n = len(cols)
    if n == 0:
        return
    o = NLEVSNO - 1                                   # row of level 0 in a (-nlevsno+1:nlevgrnd) array
    lev = np.arange(-NLEVSNO + 1, NLEVGRND + 1)[:, None]          # (37, 1) level numbers
    snl = S["snl"][cols].astype(np.int64)
    top = (snl + 1)[None, :]                          # top active level
    active = lev >= top
    ltype = S["lun_itype"][cols]
    soilcrop = ((ltype == ISTSOIL) | (ltype == ISTCROP))[None, :]
    isice = (ltype == ISTICE)[None, :]
    iswet = (ltype == ISTWET)[None, :]
    nbed = S["nbedrock"][cols][None, :]
    dz = S["dz"][:, cols]; z = S["z"][:, cols]; zi = S["zi"][:, cols]        # zi: level j at row j + 12
    t = S["t_soisno"][:, cols].copy()
    liq = S["h2osoi_liq"][:, cols].copy()
    ice = S["h2osoi_ice"][:, cols].copy()
    fsno = S["frac_sno_eff"][cols]; fh2o = S["frac_h2osfc"][cols]
    h2osfc = S["h2osfc"][cols].copy(); t_h2osfc = S["t_h2osfc"][cols].copy()
    noly = S["h2osno_no_layers"][cols].copy()
    snow_depth = S["snow_depth"][cols].copy(); int_snow = S["int_snow"][cols].copy()
    t_grnd_in = S["t_grnd"][cols]
    emg = S["emg"][cols]; htvp = S["htvp"][cols]; forc_lwrad = S["forc_lwrad"][cols]; eflx_bot = S["eflx_bot"][cols]

    def soilrows(a):                                  # (1:nlevgrnd) field -> padded to the 37 rows (snow rows = 0)
        return np.vstack([np.zeros((NLEVSNO, n)), a[:NLEVGRND][:, cols]])

    watsat, tkmg, tkdry, csol = (soilrows(S[k]) for k in ("watsat", "tkmg", "tkdry", "csol"))
    bsw, sucsat = soilrows(S["bsw"]), soilrows(S["sucsat"])
    issoil = lev >= 1
    issnow = (lev <= 0) & active & (top < 1)

    # ---------------- SoilThermProp :602-901
    thk = np.zeros((NL, n))
    with np.errstate(all="ignore"):
        satw = (liq / DENH2O + ice / DENICE) / (dz * watsat)
        satw = np.minimum(1.0, satw)
        dke = np.where(t >= TFRZ, np.maximum(0.0, np.log10(satw) + 1.0), satw)
        fl = (liq / (DENH2O * dz)) / (liq / (DENH2O * dz) + ice / (DENICE * dz))
        dksat = tkmg * np.power(TKWAT, fl * watsat) * np.power(TKICE, (1.0 - fl) * watsat)
        thk_min = np.where(satw > 0.1e-6, dke * dksat + (1.0 - dke) * tkdry, tkdry)
    thk_min = np.where(lev > nbed, THK_BEDROCK, thk_min)
    thk_wat = np.where(t < TFRZ, TKICE, TKWAT)
    thk = np.where(issoil & ~isice & ~iswet, thk_min, thk)
    thk = np.where(issoil & isice, thk_wat, thk)
    thk = np.where(issoil & iswet, np.where(lev > NLEVSOI, THK_BEDROCK, thk_wat), thk)
    with np.errstate(all="ignore"):
        bw_all = (ice + liq) / (fsno[None, :] * dz)
    method = np.where(isice, snow_glc_method, snow_method)
    jordan = TKAIR + (7.75e-5 * bw_all + 1.105e-6 * bw_all * bw_all) * (TKICE - TKAIR)
    bwk = bw_all / 1000
    sturm = np.where(bw_all <= 156, R4(0.023) + R4(0.234) * bwk, R4(0.138) - R4(1.01) * bwk + (R4(3.233) * (bwk * bwk)))
    thk = np.where(issnow, np.where(method == 1, jordan, sturm), thk)
    # interface conductivity :840-857 (row j needs row j+1)
    nxt = lambda a: np.vstack([a[1:], a[-1:]])
    zi_j = zi[1:]                                     # zi(c,j) for j = -11..25 aligned with the 37 rows
    with np.errstate(all="ignore"):
        tk = thk * nxt(thk) * (nxt(z) - z) / (thk * (nxt(z) - zi_j) + nxt(thk) * (zi_j - z))
    tk = np.where(lev == NLEVGRND, 0.0, tk)
    tk = np.where(active, tk, 0.0)
    zh2osfc = R4(1.0e-3) * (0.5 * h2osfc)
    tk_h2osfc = TKWAT * thk[o + 1] * (z[o + 1] + zh2osfc) / (TKWAT * z[o + 1] + thk[o + 1] * zh2osfc)
    cv_min = csol * (1.0 - watsat) * dz + (ice * CPICE + liq * CPLIQ)
    cv_min = np.where(lev > nbed, CSOL_BEDROCK * dz, cv_min)
    cv_w = ice * CPICE + liq * CPLIQ
    cv = np.where(isice, cv_w, np.where(iswet, np.where(lev > nbed, CSOL_BEDROCK * dz, cv_w), cv_min))
    cv[o + 1] = np.where(noly > 0.0, cv[o + 1] + CPICE * noly, cv[o + 1])
    with np.errstate(all="ignore"):
        cv_snow = np.where(fsno[None, :] > 0.0, np.maximum(THIN, (CPLIQ * liq + CPICE * ice) / fsno[None, :]), THIN)
    cv = np.where(issnow, cv_snow, cv)

    # ---------------- ComputeGroundHeatFluxAndDeriv :1543-1796 (non-urban)
    pc = S["column"][pats] - 1                        # column index (0-based, global) of every filter patch
    pos = -np.ones(S["snl"].shape[0], dtype=np.int64); pos[cols] = np.arange(n)
    pk = pos[pc]                                      # position of the patch's column inside `cols`
    w = S["wtcol"][pats]
    t_top = t[snl + 1 + o, np.arange(n)]
    lw = emg * SB * t_grnd_in ** 4
    dlw = 4.0 * emg * SB * t_grnd_in ** 3
    lw_snow = emg * SB * t_top ** 4
    lw_soil = emg * SB * t[o + 1] ** 4
    lw_h2o = emg * SB * t_h2osfc ** 4
    fv = S["frac_veg_nosno"][pats]
    dlrad = S["dlrad"][pats]
    common = dlrad + (1.0 - fv) * emg[pk] * forc_lwrad[pk]
    eflx_gnet = S["sabg"][pats] + common - lw[pk] - (S["eflx_sh_grnd"][pats] + S["qflx_evap_soi"][pats] * htvp[pk])
    sabg_chk = fsno[pk] * S["sabg_snow"][pats] + (1.0 - fsno[pk]) * S["sabg_soil"][pats]
    g_soil = S["sabg_soil"][pats] + common - lw_soil[pk] - (S["eflx_sh_soil"][pats] + S["qflx_ev_soil"][pats] * htvp[pk])
    g_h2o = S["sabg_soil"][pats] + common - lw_h2o[pk] - (S["eflx_sh_h2osfc"][pats] + S["qflx_ev_h2osfc"][pats] * htvp[pk])
    dgnetdT = -S["cgrnd"][pats] - dlw[pk]
    dhsdT = np.zeros(n); hs_soil = np.zeros(n); hs_h2osfc = np.zeros(n)
    np.add.at(dhsdT, pk, dgnetdT * w)                 # patches ascending: the reference's summation order
    np.add.at(hs_soil, pk, g_soil * w)
    np.add.at(hs_h2osfc, pk, g_h2o * w)
    lyr_top = snl[pk] + 1
    sabg_lyr = S["sabg_lyr"][:, pats]                 # (-nlevsno+1:1): level j at row j + 11
    sl_top = sabg_lyr[lyr_top + o, np.arange(len(pats))]
    gtop = sl_top + common - lw[pk] - (S["eflx_sh_grnd"][pats] + S["qflx_evap_soi"][pats] * htvp[pk])
    gtop_snow = sl_top + common - lw_snow[pk] - (S["eflx_sh_snow"][pats] + S["qflx_ev_snow"][pats] * htvp[pk])
    hs_top = np.zeros(n); hs_top_snow = np.zeros(n)
    np.add.at(hs_top, pk, gtop * w)
    np.add.at(hs_top_snow, pk, gtop_snow * w)
    sabg_lyr_col = np.zeros((NLEVSNO + 1, n))         # levels -11..1
    for r in range(NLEVSNO + 1):
        use = (r - o) >= lyr_top
        np.add.at(sabg_lyr_col[r], pk[use], (sabg_lyr[r, use] * w[use]))

    # ---------------- ComputeHeatDiffFluxAndFactor :1868-1893
    prv = lambda a: np.vstack([a[:1], a[:-1]])
    zi_jm1 = zi[:-1]                                  # zi(c,j-1) aligned with the 37 rows
    with np.errstate(all="ignore"):
        fact = dtime / cv
        fact_top = dtime / cv * dz / (0.5 * (z - zi_jm1 + CAPR * (nxt(z) - zi_jm1)))
        fact = np.where(lev == top, fact_top, fact)
        fn = tk * (nxt(t) - t) / (nxt(z) - z)
    fn = np.where(lev == NLEVGRND, eflx_bot[None, :], fn)

    # ---------------- h2osfc layer :285-298
    has_sfc = (h2osfc > THIN) & (fh2o > THIN)
    with np.errstate(all="ignore"):
        c_h2osfc = np.where(has_sfc, np.maximum(THIN, CPLIQ * h2osfc / fh2o), THIN)
        dz_h2osfc = np.where(has_sfc, np.maximum(THIN, R4(1.0e-3) * h2osfc / fh2o), THIN)

    # ---------------- SetRHSVec :1913-2353 and SetMatrix :2356-2926, assembled per unknown.
    # Unknowns of a column: rows snl .. nlevgrnd of the band system; row r <= -1 is snow layer r + 1, row 0 the
    # standing surface water, rows 1..25 soil.  b(k, r): k = 1 couples to r+2, 2 to r+1, 3 diagonal, 4 to r-1, 5 to r-2.
    dzp = nxt(z) - z
    dzm = z - prv(z)
    tkm = prv(tk)
    fnm = prv(fn)
    omc = 1.0 - CNFAC
    is_top = lev == top
    # snow rows (level j <= 0, stored at row j - 1) and soil rows
    with np.errstate(all="ignore"):
        rt_top = t + fact * (hs_top_snow[None, :] - dhsdT[None, :] * t + CNFAC * fn)
        sab37 = np.vstack([sabg_lyr_col, np.zeros((NLEVGRND - 1, n))])
        rt_snow_int = t + CNFAC * fact * (fn - fnm)
        rt_snow_int = rt_snow_int + fact * sab37
        rt_soil1 = t + fact * ((1.0 - fsno[None, :]) * (hs_soil[None, :] - dhsdT[None, :] * t) + CNFAC * (fn - fsno[None, :] * fnm))
        rt_soil1 = rt_soil1 + fsno[None, :] * fact * sab37
        rt_soil_int = t + CNFAC * fact * (fn - fnm)
        rt_soil_bot = t - CNFAC * fact * fnm + fact * fn
        d_top = 1.0 + omc * fact * tk / dzp - fact * dhsdT[None, :]
        d_int = 1.0 + omc * fact * (tk / dzp + tkm / dzm)
        up = -omc * fact * tk / dzp                   # coefficient of the layer below (j+1)
        lo_ = -omc * fact * tkm / dzm                 # coefficient of the layer above (j-1)
        d_soil1 = 1.0 + omc * fact * (tk / dzp + fsno[None, :] * tkm / dzm) - (1.0 - fsno[None, :]) * fact * dhsdT[None, :]
        lo_soil1 = -fsno[None, :] * omc * fact * tkm / dzm
        d_bot = 1.0 + omc * fact * tkm / dzm
    rt = np.where(is_top, rt_top, np.where(lev <= 0, rt_snow_int,
                  np.where(lev == 1, rt_soil1, np.where(lev == NLEVGRND, rt_soil_bot, rt_soil_int))))
    diag = np.where(is_top, d_top, np.where(lev == 1, d_soil1, np.where(lev == NLEVGRND, d_bot, d_int)))
    lower = np.where(is_top, 0.0, np.where(lev == 1, lo_soil1, lo_))
    upper = np.where(lev == NLEVGRND, 0.0, up)
    # standing surface water row and its couplings to soil layer 1
    dzm_s = R4(0.5) * dz_h2osfc + z[o + 1]
    fn_h2osfc = tk_h2osfc * (t[o + 1] - t_h2osfc) / dzm_s
    rt_ssw = t_h2osfc + (dtime / c_h2osfc) * (hs_h2osfc - dhsdT * t_h2osfc + CNFAC * fn_h2osfc)
    wet = fh2o != 0.0
    f1 = fact[o + 1]
    rt[o + 1] = np.where(wet, rt[o + 1] - fh2o * f1 * ((hs_soil - dhsdT * t[o + 1]) + CNFAC * fn_h2osfc), rt[o + 1])
    diag[o + 1] = np.where(wet, diag[o + 1] + fh2o * (omc * f1 * tk_h2osfc / dzm_s + f1 * dhsdT), diag[o + 1])
    d_ssw = 1.0 + omc * (dtime / c_h2osfc) * tk_h2osfc / dzm_s - (dtime / c_h2osfc) * dhsdT
    up_ssw = -omc * (dtime / c_h2osfc) * tk_h2osfc / dzm_s                 # ssw row, coefficient of soil layer 1
    lo_soil_ssw = np.where(wet, -fh2o * omc * f1 * tk_h2osfc / dzm_s, 0.0)  # soil-1 row, coefficient of ssw

    # ---------------- BandDiagonal -> dgbsv, one column at a time :167-219
    tnew = t.copy()
    t_ssw_new = np.zeros(n)
    for i in range(n):
        s = int(snl[i])
        nrow = NLEVGRND - s + 1                       # rows s .. 25
        ab = np.zeros((7, nrow), order="F")
        rhs = np.zeros(nrow)
        row_of = lambda r: r - s                      # band-system row r -> 0-based position

        def put(r, c, v):                             # A(r, c) in LAPACK band storage with kl = ku = 2
            ab[2 + 2 + row_of(r) - row_of(c), row_of(c)] = v

        for r in range(s, 0):                         # snow rows: layer j = r + 1
            j = r + 1 + o
            put(r, r, diag[j, i]); rhs[row_of(r)] = rt[j, i]
            if r > s:
                put(r, r - 1, lower[j, i])
            put(r, r + 1 if r < -1 else 1, upper[j, i])          # the bottom snow layer couples to soil layer 1 (band 1)
        put(0, 0, d_ssw[i]); put(0, 1, up_ssw[i]); rhs[row_of(0)] = rt_ssw[i]
        for r in range(1, NLEVGRND + 1):
            j = r + o
            put(r, r, diag[j, i]); rhs[row_of(r)] = rt[j, i]
            if r < NLEVGRND:
                put(r, r + 1, upper[j, i])
            if r == 1:
                put(1, 0, lo_soil_ssw[i])
                if s < 0:
                    put(1, -1, lower[j, i])            # band 5: to the bottom snow layer
            else:
                put(r, r - 1, lower[j, i])
        _, _, x, info = lapack.dgbsv(2, 2, ab, rhs)
        if info != 0:
            raise FloatingPointError("BandDiagonal ERROR: dgbsv returned error code (column %d)" % (cols[i] + 1))
        for r in range(s, 0):
            tnew[r + 1 + o, i] = x[row_of(r)]
        tnew[o + 1:, i] = x[row_of(1):]
        t_ssw_new[i] = x[row_of(0)]
    t = tnew
    t_h2osfc = np.where(fh2o == 0.0, t[o + 1], t_ssw_new)

    # ---------------- fn1, eflx_fgr :453-466, :575-597
    with np.errstate(all="ignore"):
        fn1 = tk * (nxt(t) - t) / (nxt(z) - z)
    fn1 = np.where(lev == NLEVGRND, 0.0, fn1)
    eflx_fgr12 = -CNFAC * fn[o + 1] - (1.0 - CNFAC) * fn1[o + 1]
    eflx_fgr = -CNFAC * fn[o + 1:] - (1.0 - CNFAC) * fn1[o + 1:]
    eflx_fgr[-1] = 0.0

    # ---------------- PhaseChangeH2osfc :904-1130
    xmf_h2osfc = np.zeros(n); qflx_h2osfc_to_ice = np.zeros(n); eflx_h2osfc_to_snow = np.zeros(n)
    h2osno_total = noly.copy()
    for r in range(NLEVSNO):
        use = (r - o) >= (snl + 1)
        h2osno_total = np.where(use, h2osno_total + ice[r] + liq[r], h2osno_total)
    fact0 = fact[o]
    for i in np.nonzero((fh2o > 0.0) & (t_h2osfc <= TFRZ))[0]:       # scalar control flow, as written in the reference
        tinc = TFRZ - t_h2osfc[i]
        t_h2osfc[i] = TFRZ
        hm = fh2o[i] * (dhsdT[i] * tinc - tinc * c_h2osfc[i] / dtime)
        xm = hm * dtime / HFUS
        temp1 = h2osfc[i] + xm
        z_avg = fsno[i] * snow_depth[i]
        rho_avg = min(800.0, h2osno_total[i] / z_avg) if z_avg > 0.0 else 200.0
        if temp1 >= 0.0:
            int_snow[i] -= xm
            if snl[i] == 0:
                noly[i] -= xm
            else:
                ice[o, i] -= xm
            h2osno_total[i] -= xm
            h2osfc[i] += xm
            xmf_h2osfc[i] = hm
            qflx_h2osfc_to_ice[i] = -xm / dtime
            if fsno[i] > 0 and snl[i] < 0:
                snow_depth[i] = h2osno_total[i] / (rho_avg * fsno[i])
            else:
                snow_depth[i] = h2osno_total[i] / DENICE
            if snl[i] == 0:
                t[o, i] = t_h2osfc[i]
                eflx_h2osfc_to_snow[i] = 0.0
            else:
                c1 = fsno[i] * (dtime / fact0[i] - dhsdT[i] * dtime) if snl[i] == -1 else fsno[i] / fact0[i] * dtime
                c2 = (-CPLIQ * xm - fh2o[i] * dhsdT[i] * dtime) if fh2o[i] != 0.0 else 0.0
                t[o, i] = (c1 * t[o, i] + c2 * t_h2osfc[i]) / (c1 + c2)
                eflx_h2osfc_to_snow[i] = (t_h2osfc[i] - t[o, i]) * c2 / dtime
        else:
            rho_avg = (h2osno_total[i] * rho_avg + h2osfc[i] * DENICE) / (h2osno_total[i] + h2osfc[i])
            int_snow[i] += h2osfc[i]
            if snl[i] == 0:
                noly[i] += h2osfc[i]
            else:
                ice[o, i] += h2osfc[i]
            h2osno_total[i] += h2osfc[i]
            qflx_h2osfc_to_ice[i] = h2osfc[i] / dtime
            t_h2osfc[i] = t_h2osfc[i] - temp1 * HFUS / (dtime * dhsdT[i] - c_h2osfc[i])
            xmf_h2osfc[i] = hm - fh2o[i] * temp1 * HFUS / dtime
            if snl[i] == 0:
                t[o, i] = t_h2osfc[i]
            else:
                c1 = fsno[i] * (dtime / fact0[i] - dhsdT[i] * dtime) if snl[i] == -1 else fsno[i] / fact0[i] * dtime
                c2 = fh2o[i] * (c_h2osfc[i] - dtime * dhsdT[i]) if fh2o[i] != 0.0 else 0.0
                t[o, i] = (c1 * t[o, i] + c2 * t_h2osfc[i]) / (c1 + c2)
                t_h2osfc[i] = t[o, i]
            h2osfc[i] = 0.0
            if fsno[i] > 0 and snl[i] < 0:
                snow_depth[i] = h2osno_total[i] / (rho_avg * fsno[i])
            else:
                snow_depth[i] = h2osno_total[i] / DENICE

    # ---------------- Phasechange :1133-1540
    xmf = np.zeros(n); qflx_snomelt = np.zeros(n); qflx_snofrz = np.zeros(n); qflx_snow_drain = np.zeros(n)
    snomelt_accum = S["snomelt_accum"][cols].copy()
    imelt = np.zeros((NL, n), dtype=np.int32)
    qsm_lyr = np.zeros((NLEVSNO, n)); qsf_lyr = np.zeros((NLEVSNO, n))
    wice0, wliq0 = ice.copy(), liq.copy()
    wmass0 = ice + liq
    tinc = np.zeros((NL, n))
    # which layers sit at the freezing point with phase to change :1245-1313
    snow_lay = (lev <= 0) & active
    m1 = snow_lay & (ice > 0.0) & (t > TFRZ)
    tinc = np.where(m1, TFRZ - t, tinc); imelt = np.where(m1, 1, imelt); t = np.where(m1, TFRZ, t)
    m2 = snow_lay & (liq > 0.0) & (t < TFRZ)
    tinc = np.where(m2, TFRZ - t, tinc); imelt = np.where(m2, 2, imelt); t = np.where(m2, TFRZ, t)
    m1 = issoil & (ice > 0.0) & (t > TFRZ)
    tinc = np.where(m1, TFRZ - t, tinc); imelt = np.where(m1, 1, imelt); t = np.where(m1, TFRZ, t)
    with np.errstate(all="ignore"):
        smp = HFUS * (TFRZ - t) / (GRAV * t) * 1000.0
        sc = watsat * np.power(smp / sucsat, -1.0 / bsw)
        sc = sc * dz * 1000.0
    supercool = np.where(issoil & soilcrop & (t < TFRZ), sc, 0.0)
    m2 = issoil & (liq > supercool) & (t < TFRZ)
    tinc = np.where(m2, TFRZ - t, tinc); imelt = np.where(m2, 2, imelt); t = np.where(m2, TFRZ, t)
    m1 = (lev == 1) & (noly[None, :] > 0.0) & (t > TFRZ)
    tinc = np.where(m1, TFRZ - t, tinc); imelt = np.where(m1, 1, imelt); t = np.where(m1, TFRZ, t)
    # energy residual and the phase change itself, level by level :1318-1517 (scalar control flow per layer)
    for i in range(n):
        s = int(snl[i])
        for j in range(s + 1, NLEVGRND + 1):
            r = j + o
            hm = 0.0
            if imelt[r, i] > 0:
                if j == s + 1:
                    if j > 0:
                        hm = dhsdT[i] * tinc[r, i] - tinc[r, i] / fact[r, i]
                    else:
                        hm = fsno[i] * (dhsdT[i] * tinc[r, i] - tinc[r, i] / fact[r, i])
                    if j == 1 and fh2o[i] != 0.0:
                        hm = hm - fh2o[i] * (dhsdT[i] * tinc[r, i])
                elif j == 1:
                    hm = (1.0 - fsno[i] - fh2o[i]) * dhsdT[i] * tinc[r, i] - tinc[r, i] / fact[r, i]
                elif j < 1:
                    hm = -fsno[i] * (tinc[r, i] / fact[r, i])
                else:
                    hm = -tinc[r, i] / fact[r, i]
            if imelt[r, i] == 1 and hm < 0.0:
                hm = 0.0; imelt[r, i] = 0
            if imelt[r, i] == 2 and hm > 0.0:
                hm = 0.0; imelt[r, i] = 0
            if not (imelt[r, i] > 0 and abs(hm) > 0.0):
                continue
            xm = hm * dtime / HFUS
            if j == 1 and noly[i] > 0.0 and xm > 0.0:
                temp1 = noly[i]
                noly[i] = max(0.0, temp1 - xm)
                propor = noly[i] / temp1
                snow_depth[i] = propor * snow_depth[i]
                heatr = hm - HFUS * (temp1 - noly[i]) / dtime
                if heatr > 0.0:
                    xm = heatr * dtime / HFUS; hm = heatr
                else:
                    xm = 0.0; hm = 0.0
                qflx_snomelt[i] = max(0.0, (temp1 - noly[i])) / dtime
                xmf[i] = HFUS * qflx_snomelt[i]
                qflx_snow_drain[i] = qflx_snomelt[i]
            heatr = 0.0
            if xm > 0.0:
                ice[r, i] = max(0.0, wice0[r, i] - xm)
                heatr = hm - HFUS * (wice0[r, i] - ice[r, i]) / dtime if j < 1 else \
                    hm - HFUS * (0.0 - 0.0 + wice0[r, i] - ice[r, i]) / dtime
            elif xm < 0.0:
                if j <= 0:
                    ice[r, i] = min(wmass0[r, i], wice0[r, i] - xm)
                elif wmass0[r, i] - 0.0 < supercool[r, i]:
                    ice[r, i] = 0.0
                else:
                    ice[r, i] = min(wmass0[r, i] - 0.0 - supercool[r, i], wice0[r, i] - xm)
                heatr = hm - HFUS * (wice0[r, i] - ice[r, i]) / dtime
            liq[r, i] = max(0.0, wmass0[r, i] - ice[r, i] - 0.0)
            if abs(heatr) > 0.0:
                if j == s + 1:
                    if j == 1:
                        t[r, i] = t[r, i] + fact[r, i] * heatr / (1.0 - (1.0 - fh2o[i]) * fact[r, i] * dhsdT[i])
                    else:
                        t[r, i] = t[r, i] + (fact[r, i] / fsno[i]) * heatr / (1.0 - fact[r, i] * dhsdT[i])
                elif j == 1:
                    t[r, i] = t[r, i] + fact[r, i] * heatr / (1.0 - (1.0 - fsno[i] - fh2o[i]) * fact[r, i] * dhsdT[i])
                elif j > 0:
                    t[r, i] = t[r, i] + fact[r, i] * heatr
                elif fsno[i] > 0.0:
                    t[r, i] = t[r, i] + (fact[r, i] / fsno[i]) * heatr
                if j <= 0 and liq[r, i] * ice[r, i] > 0.0:
                    t[r, i] = TFRZ
            if j >= 1:
                xmf[i] = xmf[i] + HFUS * (wice0[r, i] - ice[r, i]) / dtime + HFUS * (0.0 - 0.0) / dtime
            else:
                xmf[i] = xmf[i] + HFUS * (wice0[r, i] - ice[r, i]) / dtime
            if imelt[r, i] == 1 and j < 1:
                qsm_lyr[r, i] = max(0.0, (wice0[r, i] - ice[r, i])) / dtime
                qflx_snomelt[i] = qflx_snomelt[i] + qsm_lyr[r, i]
                snomelt_accum[i] = snomelt_accum[i] + qsm_lyr[r, i] * dtime * 1.0e-3
            if imelt[r, i] == 2 and j < 1:
                qsf_lyr[r, i] = max(0.0, (ice[r, i] - wice0[r, i])) / dtime
                qflx_snofrz[i] = qflx_snofrz[i] + qsf_lyr[r, i]
    eflx_snomelt = qflx_snomelt * HFUS

    # ---------------- t_grnd :543-561
    ar = np.arange(n)
    t_top_new = t[snl + 1 + o, ar]
    t1 = t[o + 1]
    tg_snow = np.where(fh2o != 0.0, fsno * t_top_new + (1.0 - fsno - fh2o) * t1 + fh2o * t_h2osfc,
                       fsno * t_top_new + (1.0 - fsno) * t1)
    tg_bare = np.where(fh2o != 0.0, (1.0 - fh2o) * t1 + fh2o * t_h2osfc, t1)
    t_grnd = np.where(snl < 0, tg_snow, tg_bare)

    # ---------------- write back (only what the reference writes: active layers / filter columns)
    def put37(name, val, mask):
        a = S[name]
        a[:, cols] = np.where(mask, val, a[:, cols])

    put37("t_soisno", t, active)
    # t_soisno(c,0) is also written by PhaseChangeH2osfc when there is no snow layer (:1028, :1078)
    S["t_soisno"][o, cols] = t[o]
    put37("h2osoi_liq", liq, active)
    put37("h2osoi_ice", ice, active | (lev == 0))
    put37("imelt", imelt, active)
    put37("fact", fact, active)
    put37("thk", thk, (issoil | issnow))
    a = S["bw"]; a[:, cols] = np.where(issnow[:NLEVSNO], bw_all[:NLEVSNO], a[:, cols])
    S["qflx_snomelt_lyr"][:, cols] = qsm_lyr
    S["qflx_snofrz_lyr"][:, cols] = qsf_lyr
    ef = S["eflx_fgr"]
    ef[:NLEVGRND, cols] = np.where(soilcrop, eflx_fgr, ef[:NLEVGRND][:, cols])
    for name, val in (("t_grnd", t_grnd), ("t_h2osfc", t_h2osfc), ("h2osfc", h2osfc), ("h2osno_no_layers", noly),
                      ("int_snow", int_snow), ("snow_depth", snow_depth), ("snomelt_accum", snomelt_accum),
                      ("c_h2osfc", c_h2osfc), ("xmf", xmf), ("xmf_h2osfc", xmf_h2osfc), ("eflx_fgr12", eflx_fgr12),
                      ("qflx_h2osfc_to_ice", qflx_h2osfc_to_ice), ("eflx_h2osfc_to_snow", eflx_h2osfc_to_snow),
                      ("qflx_snow_drain", qflx_snow_drain), ("qflx_snofrz", qflx_snofrz), ("qflx_snomelt", qflx_snomelt),
                      ("eflx_snomelt", eflx_snomelt)):
        S[name][cols] = val
    S["eflx_snomelt_r"][cols] = np.where(soilcrop[0], eflx_snomelt, S["eflx_snomelt_r"][cols])
    S["eflx_gnet"][pats] = eflx_gnet
    S["dgnetdT"][pats] = dgnetdT
    S["sabg_chk"][pats] = sabg_chk
