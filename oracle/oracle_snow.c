/* oracle_snow.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the snow routines HydrologyNoDrainage calls, one loop per reference loop,
 * in the reference's order (src/biogeophys/SnowHydrologyMod.F90 unless another file is named):
 *   BuildSnowFilter                         :3975-4010
 *   SnowWater                               :1015-1165  = UpdateState_TopLayerFluxes :1168-1290 (truncate_small_values_one_lev,
 *                                           NumericsMod.F90:101-160), BulkFlux_SnowPercolation :1293-1382,
 *                                           UpdateState_SnowPercolation :1453-1496, CalcAndApplyAerosolFluxes :1499-1708 with
 *                                           AerosolFluxes AerosolMod.F90:668-801, PostPercolation_AdjustLayerThicknesses :1711-1750,
 *                                           BulkDiag_SnowWaterAccumulatedSnow :1753-1813, SumFlux_AddSnowPercolation :1816-1867
 *   SnowCompaction                          :1870-2080 with OverburdenCompactionAnderson1976 :3766, ...Vionnet2012 :3794,
 *                                           WindDriftCompaction :3835, FracSnowDuringMelt
 *                                           SnowCoverFractionSwensonLawrence2012Mod.F90:244-268
 *   CombineSnowLayers                       :2083-2507 with Combo :3902-3946
 *   DivideSnowLayers (is_lake = .false.)    :2510-2895 with MassWeightedSnowRadius :3949-3972
 *   ZeroEmptySnowLayers                     :2898-2952
 * Configuration: bulk water only (no water tracers), non-lake and non-urban columns.
 * Parity is pinned by tests/test_oracle_snow.py: an independent Python restatement written from the Fortran, plus the
 * conservation laws the routines imply (water, enthalpy, aerosol mass, layer-thickness limits).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define NSNO CTSM_NLEVSNO
#define SNO_LO (-CTSM_NLEVSNO + 1)
static const double denice = 0.917e3, denh2o = 1.000e3, tfrz = 273.15, cpice = 2.11727e3, cpliq = 4.188e3, hfus = 3.337e5;
static const double rpi = 3.14159265358979323846, snw_rds_max = 1500.0;
static const double scvng_fct_mlt_ocphi = 0.20, scvng_fct_mlt_ocpho = 0.03;      /* SnowHydrologyMod.F90:131-132 */

/* InitSnowLayers :2985-3002: the layer-thickness limits derived from the six namelist values */
void oracle_snow_dz_limits(const ctsm_params_t* prm, double* dzmin, double* dzmax_l, double* dzmax_u) {
  dzmin[0] = prm->snow_dzmin_1; dzmax_l[0] = prm->snow_dzmax_l_1; dzmax_u[0] = prm->snow_dzmax_u_1;
  dzmin[1] = prm->snow_dzmin_2; dzmax_l[1] = prm->snow_dzmax_l_2; dzmax_u[1] = prm->snow_dzmax_u_2;
  for (int j = 2; j < NSNO; ++j) {
    dzmin[j] = dzmax_u[j - 1] * 0.5;
    dzmax_u[j] = 2.0 * dzmax_u[j - 1] + 0.01;
    dzmax_l[j] = dzmax_u[j] + dzmax_l[j - 1];
    if (j == NSNO - 1) { dzmax_u[j] = 1.79769313486231571e308; dzmax_l[j] = 1.79769313486231571e308; }
  }
}

/* BuildSnowFilter :3998-4009 */
void oracle_build_snow_filter(int num_nolakec, const int32_t* filter_nolakec, const int32_t* snl, int begc0,
                              int32_t* filter_snowc, int32_t* num_snowc, int32_t* filter_nosnowc, int32_t* num_nosnowc) {
  int ns = 0, nn = 0;
  for (int fc = 0; fc < num_nolakec; ++fc) {
    const int c = filter_nolakec[fc];
    if (snl[c - begc0] < 0) filter_snowc[ns++] = c;
    else filter_nosnowc[nn++] = c;
  }
  *num_snowc = ns; *num_nosnowc = nn;
}

#define CC(name, c) f->name[(c) - begc0]
#define CS(name, c, j) f->name[(size_t)((j) - SNO_LO) * ldc + ((c) - begc0)]         /* SNO and SNOSOI arrays start at -nlevsno+1 */
#define CZ(name, c, j) f->name[(size_t)((j) + NSNO) * ldc + ((c) - begc0)]           /* SNOSOI0 (zi) starts at -nlevsno */

static int snow_fail(ctsm_status_t* st, int c, double v, const char* msg) {
  if (st) {
    st->code = CTSM_ERR_SNOW_NEGATIVE; st->subgrid_level = CTSM_SUBGRID_COLUMN; st->subgrid_index = c; st->value = v;
    snprintf(st->msg, sizeof st->msg, "%s", msg);
  }
  return CTSM_ERR_SNOW_NEGATIVE;
}

int oracle_snow_water(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_snowc, const int32_t* filter_snowc,
                      int num_nosnowc, const int32_t* filter_nosnowc, const ctsm_snowwater_fields_t* f, ctsm_status_t* st) {
  if (st) memset(st, 0, sizeof *st);
  const int begc0 = f->alloc.begc, begg0 = f->alloc.begg;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1), ldg = (size_t)(f->alloc.endg - f->alloc.begg + 1);
  const double dtime = prm->dtime;
  const int nb = bounds->endc - bounds->begc + 1;
  if (nb <= 0) return 0;
  double* ice0 = (double*)calloc((size_t)nb, sizeof(double));
  double* liq0 = (double*)calloc((size_t)nb, sizeof(double));
  double* vol_liq = (double*)calloc((size_t)nb * NSNO, sizeof(double));
  double* vol_ice = (double*)calloc((size_t)nb * NSNO, sizeof(double));
  double* eff_por = (double*)calloc((size_t)nb * NSNO, sizeof(double));
  double* qin = (double*)calloc((size_t)nb * 8, sizeof(double));
#define LOC(a, c, j) a[(size_t)((j) - SNO_LO) * nb + ((c) - bounds->begc)]
  int rc = 0;

  /* UpdateState_TopLayerFluxes :1210-1225 */
  for (int fc = 0; fc < num_snowc; ++fc) {
    const int c = filter_snowc[fc], top = CC(snl, c) + 1;
    ice0[c - bounds->begc] = CS(h2osoi_ice, c, top);
    liq0[c - bounds->begc] = CS(h2osoi_liq, c, top);
    CS(h2osoi_ice, c, top) = CS(h2osoi_ice, c, top)
        + CC(frac_sno_eff, c) * (CC(qflx_soliddew_to_top_layer, c) - CC(qflx_solidevap_from_top_layer, c)) * dtime;
    CS(h2osoi_liq, c, top) = CS(h2osoi_liq, c, top)
        + CC(frac_sno_eff, c) * (CC(qflx_liq_grnd, c) + CC(qflx_liqdew_to_top_layer, c) - CC(qflx_liqevap_from_top_layer, c)) * dtime;
  }
  /* truncate_small_values_one_lev (custom_rel_epsilon = 1e-12) :1231-1251 */
  for (int fc = 0; fc < num_snowc; ++fc) {
    const int c = filter_snowc[fc], top = CC(snl, c) + 1;
    if (fabs(CS(h2osoi_ice, c, top)) < 1.e-12 * fabs(ice0[c - bounds->begc])) CS(h2osoi_ice, c, top) = 0.0;
  }
  for (int fc = 0; fc < num_snowc; ++fc) {
    const int c = filter_snowc[fc], top = CC(snl, c) + 1;
    if (fabs(CS(h2osoi_liq, c, top)) < 1.e-12 * fabs(liq0[c - bounds->begc])) CS(h2osoi_liq, c, top) = 0.0;
  }
  for (int fc = 0; fc < num_snowc; ++fc) {                                        /* :1256-1287 */
    const int c = filter_snowc[fc], top = CC(snl, c) + 1;
    if (CS(h2osoi_ice, c, top) < 0.0) {
      rc = snow_fail(st, c, CS(h2osoi_ice, c, top), "In UpdateState_TopLayerFluxes, h2osoi_ice has gone significantly negative");
      goto done;
    }
    if (CS(h2osoi_liq, c, top) < 0.0) {
      rc = snow_fail(st, c, CS(h2osoi_liq, c, top), "In UpdateState_TopLayerFluxes, h2osoi_liq has gone significantly negative");
      goto done;
    }
  }
  /* BulkFlux_SnowPercolation :1331-1380 */
  for (int j = SNO_LO; j <= 0; ++j)
    for (int fc = 0; fc < num_snowc; ++fc) {
      const int c = filter_snowc[fc];
      if (j >= CC(snl, c) + 1) {
        LOC(vol_ice, c, j) = fmin(1.0, CS(h2osoi_ice, c, j) / (CS(dz, c, j) * CC(frac_sno_eff, c) * denice));
        LOC(eff_por, c, j) = 1.0 - LOC(vol_ice, c, j);
        LOC(vol_liq, c, j) = fmin(LOC(eff_por, c, j), CS(h2osoi_liq, c, j) / (CS(dz, c, j) * CC(frac_sno_eff, c) * denh2o));
      }
    }
  for (int j = SNO_LO; j <= 0; ++j)
    for (int fc = 0; fc < num_snowc; ++fc) {
      const int c = filter_snowc[fc];
      if (j >= CC(snl, c) + 1) {
        double q;
        if (j <= -1) {
          if (LOC(eff_por, c, j) < prm->wimp || LOC(eff_por, c, j + 1) < prm->wimp) {
            q = 0.0;
          } else {
            q = fmax(0.0, (LOC(vol_liq, c, j) - prm->ssi * LOC(eff_por, c, j)) * CS(dz, c, j) * CC(frac_sno_eff, c));
            q = fmin(q, (1.0 - LOC(vol_ice, c, j + 1) - LOC(vol_liq, c, j + 1)) * CS(dz, c, j + 1) * CC(frac_sno_eff, c));
          }
        } else {
          q = fmax(0.0, (LOC(vol_liq, c, j) - prm->ssi * LOC(eff_por, c, j)) * CS(dz, c, j) * CC(frac_sno_eff, c));
        }
        CS(qflx_snow_percolation, c, j) = (q * 1000.0) / dtime;
      }
    }
  /* UpdateState_SnowPercolation :1483-1494 */
  for (int j = SNO_LO; j <= 0; ++j)
    for (int fc = 0; fc < num_snowc; ++fc) {
      const int c = filter_snowc[fc];
      if (j >= CC(snl, c) + 1) {
        if (j >= CC(snl, c) + 2) CS(h2osoi_liq, c, j) = CS(h2osoi_liq, c, j) + CS(qflx_snow_percolation, c, j - 1) * dtime;
        CS(h2osoi_liq, c, j) = CS(h2osoi_liq, c, j) - CS(qflx_snow_percolation, c, j) * dtime;
      }
    }
  /* CalcAndApplyAerosolFluxes :1563-1700: the eight species in the reference's order */
  {
    double* mss[8] = {f->mss_bcphi, f->mss_bcpho, f->mss_ocphi, f->mss_ocpho, f->mss_dst1, f->mss_dst2, f->mss_dst3, f->mss_dst4};
    const double scv[8] = {prm->scvng_fct_mlt_bcphi, prm->scvng_fct_mlt_bcpho, scvng_fct_mlt_ocphi, scvng_fct_mlt_ocpho,
                           prm->scvng_fct_mlt_dst1, prm->scvng_fct_mlt_dst2, prm->scvng_fct_mlt_dst3, prm->scvng_fct_mlt_dst4};
    for (int j = SNO_LO; j <= 0; ++j)
      for (int fc = 0; fc < num_snowc; ++fc) {
        const int c = filter_snowc[fc];
        if (j >= CC(snl, c) + 1) {
          const size_t o = (size_t)(j - SNO_LO) * ldc + (c - begc0);
          for (int k = 0; k < 8; ++k) mss[k][o] = mss[k][o] + qin[(size_t)k * nb + (c - bounds->begc)] * dtime;
          double mss_liqice = CS(h2osoi_liq, c, j) + CS(h2osoi_ice, c, j);
          if (mss_liqice < 1e-30) mss_liqice = 1e-30;
          for (int k = 0; k < 8; ++k) {
            /* the OC factors are module parameters: scvng_fct_mlt_sf * factor is evaluated left to right in both forms */
            double qout = CS(qflx_snow_percolation, c, j) * prm->scvng_fct_mlt_sf * scv[k] * (mss[k][o] / mss_liqice);
            if (qout * dtime > mss[k][o]) {
              qout = mss[k][o] / dtime;
              mss[k][o] = 0.0;
            } else {
              mss[k][o] = mss[k][o] - qout * dtime;
            }
            qin[(size_t)k * nb + (c - bounds->begc)] = qout;
          }
        }
      }
    /* AerosolFluxes, AerosolMod.F90:725-798 */
#define AER(g, k) f->forc_aer[(size_t)((k) - 1) * ldg + ((g) - begg0)]
    for (int c = bounds->begc; c <= bounds->endc; ++c) {
      const int g = CC(col_gridcell, c);
      const int on = prm->snicar_use_aerosol != 0;
      CC(flx_bc_dep_dry, c) = on ? AER(g, 1) + AER(g, 2) : 0.0;
      CC(flx_bc_dep_wet, c) = on ? AER(g, 3) : 0.0;
      CC(flx_bc_dep_phi, c) = on ? AER(g, 1) + AER(g, 3) : 0.0;
      CC(flx_bc_dep_pho, c) = on ? AER(g, 2) : 0.0;
      CC(flx_bc_dep, c) = on ? AER(g, 1) + AER(g, 2) + AER(g, 3) : 0.0;
      CC(flx_oc_dep_dry, c) = on ? AER(g, 4) + AER(g, 5) : 0.0;
      CC(flx_oc_dep_wet, c) = on ? AER(g, 6) : 0.0;
      CC(flx_oc_dep_phi, c) = on ? AER(g, 4) + AER(g, 6) : 0.0;
      CC(flx_oc_dep_pho, c) = on ? AER(g, 5) : 0.0;
      CC(flx_oc_dep, c) = on ? AER(g, 4) + AER(g, 5) + AER(g, 6) : 0.0;
      CC(flx_dst_dep_wet1, c) = on ? AER(g, 7) : 0.0;
      CC(flx_dst_dep_dry1, c) = on ? AER(g, 8) : 0.0;
      CC(flx_dst_dep_wet2, c) = on ? AER(g, 9) : 0.0;
      CC(flx_dst_dep_dry2, c) = on ? AER(g, 10) : 0.0;
      CC(flx_dst_dep_wet3, c) = on ? AER(g, 11) : 0.0;
      CC(flx_dst_dep_dry3, c) = on ? AER(g, 12) : 0.0;
      CC(flx_dst_dep_wet4, c) = on ? AER(g, 13) : 0.0;
      CC(flx_dst_dep_dry4, c) = on ? AER(g, 14) : 0.0;
      CC(flx_dst_dep, c) = on ? AER(g, 7) + AER(g, 8) + AER(g, 9) + AER(g, 10) + AER(g, 11) + AER(g, 12) + AER(g, 13) + AER(g, 14) : 0.0;
    }
    for (int fc = 0; fc < num_snowc; ++fc) {
      const int c = filter_snowc[fc], top = CC(snl, c) + 1;
      CS(mss_bcphi, c, top) = CS(mss_bcphi, c, top) + (CC(flx_bc_dep_phi, c) * dtime);
      CS(mss_bcpho, c, top) = CS(mss_bcpho, c, top) + (CC(flx_bc_dep_pho, c) * dtime);
      CS(mss_ocphi, c, top) = CS(mss_ocphi, c, top) + (CC(flx_oc_dep_phi, c) * dtime);
      CS(mss_ocpho, c, top) = CS(mss_ocpho, c, top) + (CC(flx_oc_dep_pho, c) * dtime);
      CS(mss_dst1, c, top) = CS(mss_dst1, c, top) + (CC(flx_dst_dep_dry1, c) + CC(flx_dst_dep_wet1, c)) * dtime;
      CS(mss_dst2, c, top) = CS(mss_dst2, c, top) + (CC(flx_dst_dep_dry2, c) + CC(flx_dst_dep_wet2, c)) * dtime;
      CS(mss_dst3, c, top) = CS(mss_dst3, c, top) + (CC(flx_dst_dep_dry3, c) + CC(flx_dst_dep_wet3, c)) * dtime;
      CS(mss_dst4, c, top) = CS(mss_dst4, c, top) + (CC(flx_dst_dep_dry4, c) + CC(flx_dst_dep_wet4, c)) * dtime;
    }
#undef AER
  }
  /* PostPercolation_AdjustLayerThicknesses :1740-1748 */
  for (int j = SNO_LO; j <= 0; ++j)
    for (int fc = 0; fc < num_snowc; ++fc) {
      const int c = filter_snowc[fc];
      if (j >= CC(snl, c) + 1) CS(dz, c, j) = fmax(CS(dz, c, j), CS(h2osoi_liq, c, j) / denh2o + CS(h2osoi_ice, c, j) / denice);
    }
  /* BulkDiag_SnowWaterAccumulatedSnow :1794-1811 */
  for (int fc = 0; fc < num_snowc; ++fc) {
    const int c = filter_snowc[fc];
    CC(int_snow, c) = CC(int_snow, c)
        + CC(frac_sno_eff, c) * (CC(qflx_soliddew_to_top_layer, c) + CC(qflx_liqdew_to_top_layer, c) + CC(qflx_liq_grnd, c)) * dtime;
  }
  for (int fc = 0; fc < num_nosnowc; ++fc) {
    const int c = filter_nosnowc[fc];
    if (CC(h2osno_no_layers, c) <= 0.0) { CC(int_snow, c) = 0.0; CC(frac_sno, c) = 0.0; CC(snow_depth, c) = 0.0; }
  }
  /* SumFlux_AddSnowPercolation :1851-1865 */
  for (int fc = 0; fc < num_snowc; ++fc) {
    const int c = filter_snowc[fc];
    CC(qflx_snow_drain, c) = CC(qflx_snow_drain, c) + CS(qflx_snow_percolation, c, 0);
    CC(qflx_rain_plus_snomelt, c) = CS(qflx_snow_percolation, c, 0) + (1.0 - CC(frac_sno_eff, c)) * CC(qflx_liq_grnd, c);
  }
  for (int fc = 0; fc < num_nosnowc; ++fc) {
    const int c = filter_nosnowc[fc];
    CC(qflx_snow_drain, c) = CC(qflx_snomelt, c);
    CC(qflx_rain_plus_snomelt, c) = CC(qflx_liq_grnd, c) + CC(qflx_snomelt, c);
  }
done:
  free(ice0); free(liq0); free(vol_liq); free(vol_ice); free(eff_por); free(qin);
  return rc;
#undef LOC
}

/* Combo :3902-3946 */
static void combo(double* dz, double* wliq, double* wice, double* t, double dz2, double wliq2, double wice2, double t2) {
  const double dzc = *dz + dz2;
  const double wicec = (*wice + wice2);
  const double wliqc = (*wliq + wliq2);
  const double h = (cpice * *wice + cpliq * *wliq) * (*t - tfrz) + hfus * *wliq;
  const double h2 = (cpice * wice2 + cpliq * wliq2) * (t2 - tfrz) + hfus * wliq2;
  const double hc = h + h2;
  const double tc = tfrz + (hc - hfus * wliqc) / (cpice * wicec + cpliq * wliqc);
  *dz = dzc; *wice = wicec; *wliq = wliqc; *t = tc;
}

/* MassWeightedSnowRadius :3949-3972 */
static double mass_weighted_snow_radius(const ctsm_params_t* prm, double rds1, double rds2, double swtot, double zwtot) {
  double r = (rds2 * swtot + rds1 * zwtot) / (swtot + zwtot);
  if (r > snw_rds_max) r = snw_rds_max;
  else if (r < prm->snw_rds_min) r = prm->snw_rds_min;
  return r;
}

int oracle_snow_layers(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_snowc, const int32_t* filter_snowc,
                       const ctsm_snowlayers_fields_t* f, ctsm_status_t* st) {
  (void)bounds;
  if (st) memset(st, 0, sizeof *st);
  const int begc0 = f->alloc.begc, begg0 = f->alloc.begg;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1);
  const double dtime = prm->dtime;
  double dzmin[NSNO], dzmax_l[NSNO], dzmax_u[NSNO];
  oracle_snow_dz_limits(prm, dzmin, dzmax_l, dzmax_u);
  double* mss[8] = {f->mss_bcphi, f->mss_bcpho, f->mss_ocphi, f->mss_ocpho, f->mss_dst1, f->mss_dst2, f->mss_dst3, f->mss_dst4};
#define MS(k, c, j) mss[k][(size_t)((j) - SNO_LO) * ldc + ((c) - begc0)]
  for (int fc = 0; fc < num_snowc; ++fc) {
    const int c = filter_snowc[fc], lt = CC(lun_itype, c);
    if ((lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) || lt == CTSM_ISTDLAK) {
      const int code = lt == CTSM_ISTDLAK ? CTSM_ERR_BAD_ARG : CTSM_ERR_URBAN;
      if (st) { st->code = code; st->subgrid_index = c; }
      return code;
    }
  }

  /* ---------------- SnowCompaction :1947-2077 ---------------- */
  {
    const double c3 = 2.777e-6, c4 = 0.04, c5 = 2.0;
    for (int fc = 0; fc < num_snowc; ++fc) {
      const int c = filter_snowc[fc];
      const int g = CC(col_gridcell, c);
      double burden = 0.0, zpseudo = 0.0;
      int mobile = 1;
      const double frac_sno = CC(frac_sno_eff, c);
      for (int j = SNO_LO; j <= 0; ++j) {
        if (j < CC(snl, c) + 1) continue;
        const double ice = CS(h2osoi_ice, c, j), liq = CS(h2osoi_liq, c, j);
        const double wx = (ice + liq);
        const double voidf = 1.0 - (ice / denice + liq / denh2o) / (frac_sno * CS(dz, c, j));
        if (voidf > 0.001 && ice > .1) {
          const double bi = ice / (frac_sno * CS(dz, c, j));
          const double fi = ice / wx;
          const double td = tfrz - CS(t_soisno, c, j);
          const double dexpf = exp(-c4 * td);
          double ddz1 = -c3 * dexpf;
          if (bi > prm->upplim_destruct_metamorph) ddz1 = ddz1 * exp(-46.0e-3 * (bi - prm->upplim_destruct_metamorph));
          if (liq > 0.01 * CS(dz, c, j) * frac_sno) ddz1 = ddz1 * c5;
          double ddz2;
          if (prm->snow_overburden_compaction_method == 1) {                                  /* Anderson1976 :3784-3789 */
            const double c2 = 23.e-3;
            ddz2 = -(burden + wx / 2.0) * exp(-prm->overburden_compress_Tfactor * td - c2 * bi) / prm->eta0_anderson;
          } else {                                                                            /* Vionnet2012 :3825-3830 */
            const double aeta = 0.1, beta = 0.023;
            const double f1 = 1.0 / (1.0 + 60.0 * liq / (denh2o * CS(dz, c, j)));
            const double f2 = 4.0;
            const double eta = f1 * f2 * (bi / prm->ceta) * exp(aeta * td + beta * bi) * prm->eta0_vionnet;
            ddz2 = -(burden + wx / 2.0) / eta;
          }
          double ddz3;
          if (CS(imelt, c, j) == 1) {
            if (prm->use_subgrid_fluxes) {                                                    /* (never lake / urban here) */
              ddz3 = fmax(0.0, fmin(1.0, (CS(swe_old, c, j) - wx) / wx));
              if ((CS(swe_old, c, j) - wx) > 0.0) {
                double wsum = 0.0;
                for (int jj = CC(snl, c) + 1; jj <= 0; ++jj) wsum += CS(h2osoi_liq, c, jj) + CS(h2osoi_ice, c, jj);
                /* FracSnowDuringMelt, SnowCoverFractionSwensonLawrence2012Mod.F90:263-266 */
                const double int_snow_limited = fmin(CC(int_snow, c), prm->int_snow_max);
                const double smr = fmin(1.0, wsum / int_snow_limited);
                double fsno_melt = 1. - pow(acos(fmin(1.0, (2. * smr - 1.0))) / rpi, CC(n_melt, c));
                if ((fsno_melt + CC(frac_h2osfc, c)) > 1.0) fsno_melt = 1.0 - CC(frac_h2osfc, c);
                ddz3 = ddz3 - fmax(0.0, (fsno_melt - frac_sno) / frac_sno);
              }
              ddz3 = -1.0 / dtime * ddz3;
            } else {
              ddz3 = -1.0 / dtime * fmax(0.0, (CS(frac_iceold, c, j) - fi) / CS(frac_iceold, c, j));
            }
          } else {
            ddz3 = 0.0;
          }
          double ddz4;
          if (prm->wind_dependent_snow_density) {                                             /* WindDriftCompaction :3872-3897 */
            const double rho_min = 50.0, drift_sph = 1.0;
            if (mobile) {
              const double Frho = 1.25 - 0.0042 * (fmax(rho_min, bi) - rho_min);
              const double MO = 0.34 * (-0.583 * prm->drift_gs - 0.833 * drift_sph + 0.833) + 0.66 * Frho;
              double SI = -2.868 * exp(-0.085 * f->forc_wind[g - begg0]) + 1.0 + MO;
              if (SI > 0.0) {
                SI = fmin(SI, 3.25);
                zpseudo = zpseudo + 0.5 * CS(dz, c, j) * (3.25 - SI);
                const double gamma_drift = SI * exp(-zpseudo / 0.1);
                const double tau_inverse = gamma_drift / prm->tau_ref;
                ddz4 = -fmax(0.0, prm->rho_max - bi) * tau_inverse;
                zpseudo = zpseudo + 0.5 * CS(dz, c, j) * (3.25 - SI);
              } else {
                mobile = 0;
                ddz4 = 0.0;
              }
            } else {
              ddz4 = 0.0;
            }
          } else {
            ddz4 = 0.0;
          }
          const double pdzdtc = ddz1 + ddz2 + ddz3 + ddz4;
          CS(dz, c, j) = fmax(CS(dz, c, j) * (1.0 + pdzdtc * dtime), (ice / denice + liq / denh2o) / frac_sno);
        } else {
          mobile = 0;
        }
        burden = burden + wx;
      }
    }
  }

  /* ---------------- CombineSnowLayers :2167-2503 (non-lake: dzminloc = dzmin) ---------------- */
  for (int fc = 0; fc < num_snowc; ++fc) CC(qflx_sl_top_soil, filter_snowc[fc]) = 0.0;
  for (int fc = 0; fc < num_snowc; ++fc) {                                                     /* :2198-2284 */
    const int c = filter_snowc[fc], lt = CC(lun_itype, c);
    const int msn_old = CC(snl, c);
    const int soil = (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP);
    for (int j = msn_old + 1; j <= 0; ++j) {
      if (CS(h2osoi_ice, c, j) <= .01) {
        if (j < 0 || soil) {
          CS(h2osoi_liq, c, j + 1) = CS(h2osoi_liq, c, j + 1) + CS(h2osoi_liq, c, j);
          CS(h2osoi_ice, c, j + 1) = CS(h2osoi_ice, c, j + 1) + CS(h2osoi_ice, c, j);
        }
        if (j < 0) {
          CS(dz, c, j + 1) = CS(dz, c, j + 1) + CS(dz, c, j);
          for (int k = 0; k < 8; ++k) MS(k, c, j + 1) = MS(k, c, j + 1) + MS(k, c, j);
        }
        if (j == 0) CC(qflx_sl_top_soil, c) = (CS(h2osoi_liq, c, j) + CS(h2osoi_ice, c, j)) / dtime;
        if (j > CC(snl, c) + 1 && CC(snl, c) < -1) {
          for (int i = j; i >= CC(snl, c) + 2; --i) {
            CS(h2osoi_liq, c, i) = CS(h2osoi_liq, c, i - 1);
            CS(h2osoi_ice, c, i) = CS(h2osoi_ice, c, i - 1);
            CS(t_soisno, c, i) = CS(t_soisno, c, i - 1);
            for (int k = 0; k < 8; ++k) MS(k, c, i) = MS(k, c, i - 1);
            CS(snw_rds, c, i) = CS(snw_rds, c, i - 1);
            CS(dz, c, i) = CS(dz, c, i - 1);
          }
        }
        CC(snl, c) = CC(snl, c) + 1;
      }
    }
  }
  for (int fc = 0; fc < num_snowc; ++fc) {                                                     /* :2286-2380 */
    const int c = filter_snowc[fc], lt = CC(lun_itype, c);
    double zwice = 0.0, zwliq = 0.0, h2osno_total = 0.0;
    CC(snow_depth, c) = 0.0;
    for (int j = SNO_LO; j <= 0; ++j)
      if (j >= CC(snl, c) + 1) {
        zwice = zwice + CS(h2osoi_ice, c, j);
        zwliq = zwliq + CS(h2osoi_liq, c, j);
        CC(snow_depth, c) = CC(snow_depth, c) + CS(dz, c, j);
        h2osno_total = h2osno_total + CS(h2osoi_ice, c, j) + CS(h2osoi_liq, c, j);
      }
    if (CC(snow_depth, c) > 0.0) {
      if ((CC(frac_sno_eff, c) * CC(snow_depth, c) < dzmin[0]) ||
          (h2osno_total / (CC(frac_sno_eff, c) * CC(snow_depth, c)) < 50.0)) {
        CC(h2osno_no_layers, c) = zwice;
        if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP) CS(h2osoi_liq, c, 1) = CS(h2osoi_liq, c, 1) + zwliq;
        CC(snl, c) = 0;
        h2osno_total = CC(h2osno_no_layers, c);
        for (int k = 0; k < 8; ++k)
          for (int j = SNO_LO; j <= 0; ++j) MS(k, c, j) = 0.0;
        if (CC(h2osno_no_layers, c) <= 0.0) CC(snow_depth, c) = 0.0;
      }
    }
    if (h2osno_total <= 0.0) {
      CC(snow_depth, c) = 0.0; CC(frac_sno, c) = 0.0; CC(frac_sno_eff, c) = 0.0; CC(int_snow, c) = 0.0;
    }
  }
  for (int fc = 0; fc < num_snowc; ++fc) {                                                     /* :2382-2489 */
    const int c = filter_snowc[fc];
    if (CC(snl, c) < -1) {
      const int msn_old = CC(snl, c);
      int mssi = 1;
      for (int i = msn_old + 1; i <= 0; ++i) {
        if ((CC(frac_sno_eff, c) * CS(dz, c, i) < dzmin[mssi - 1]) ||
            ((CS(h2osoi_ice, c, i) + CS(h2osoi_liq, c, i)) / (CC(frac_sno_eff, c) * CS(dz, c, i)) < 50.0)) {
          int neibor;
          if (i == CC(snl, c) + 1) neibor = i + 1;
          else if (i == 0) neibor = i - 1;
          else {
            neibor = i + 1;
            if ((CS(dz, c, i - 1) + CS(dz, c, i)) < (CS(dz, c, i + 1) + CS(dz, c, i))) neibor = i - 1;
          }
          int j, l;
          if (neibor > i) { j = neibor; l = i; } else { j = i; l = neibor; }
          for (int k = 0; k < 8; ++k) MS(k, c, j) = MS(k, c, j) + MS(k, c, l);
          CS(snw_rds, c, j) = (CS(snw_rds, c, j) * (CS(h2osoi_liq, c, j) + CS(h2osoi_ice, c, j)) +
                               CS(snw_rds, c, l) * (CS(h2osoi_liq, c, l) + CS(h2osoi_ice, c, l))) /
                              (CS(h2osoi_liq, c, j) + CS(h2osoi_ice, c, j) + CS(h2osoi_liq, c, l) + CS(h2osoi_ice, c, l));
          combo(&CS(dz, c, j), &CS(h2osoi_liq, c, j), &CS(h2osoi_ice, c, j), &CS(t_soisno, c, j), CS(dz, c, l),
                CS(h2osoi_liq, c, l), CS(h2osoi_ice, c, l), CS(t_soisno, c, l));
          if (j - 1 > CC(snl, c) + 1) {
            for (int k = j - 1; k >= CC(snl, c) + 2; --k) {
              CS(h2osoi_ice, c, k) = CS(h2osoi_ice, c, k - 1);
              CS(h2osoi_liq, c, k) = CS(h2osoi_liq, c, k - 1);
              CS(t_soisno, c, k) = CS(t_soisno, c, k - 1);
              for (int a = 0; a < 8; ++a) MS(a, c, k) = MS(a, c, k - 1);
              CS(snw_rds, c, k) = CS(snw_rds, c, k - 1);
              CS(dz, c, k) = CS(dz, c, k - 1);
            }
          }
          CC(snl, c) = CC(snl, c) + 1;
          if (CC(snl, c) >= -1) break;
        } else {
          mssi = mssi + 1;
        }
      }
    }
  }
  for (int j = 0; j >= SNO_LO; --j)                                                            /* :2493-2501 */
    for (int fc = 0; fc < num_snowc; ++fc) {
      const int c = filter_snowc[fc];
      if (j >= CC(snl, c) + 1) {
        CS(z, c, j) = CZ(zi, c, j) - 0.5 * CS(dz, c, j);
        CZ(zi, c, j - 1) = CZ(zi, c, j) - CS(dz, c, j);
      }
    }

  /* ---------------- DivideSnowLayers :2620-2890, is_lake = .false. ---------------- */
  for (int fc = 0; fc < num_snowc; ++fc) {
    const int c = filter_snowc[fc];
    double dzsno[NSNO + 1], swice[NSNO + 1], swliq[NSNO + 1], tsno[NSNO + 1], rds[NSNO + 1], ma[8][NSNO + 1];
    const double frac_sno = CC(frac_sno_eff, c);
    const int snl0 = CC(snl, c);
    memset(dzsno, 0, sizeof dzsno); memset(swice, 0, sizeof swice); memset(swliq, 0, sizeof swliq);
    memset(tsno, 0, sizeof tsno); memset(rds, 0, sizeof rds); memset(ma, 0, sizeof ma);
    for (int j = 1; j <= NSNO; ++j)                                                            /* 1-based as in the reference */
      if (j <= abs(snl0)) {
        dzsno[j] = frac_sno * CS(dz, c, j + snl0);
        swice[j] = CS(h2osoi_ice, c, j + snl0);
        swliq[j] = CS(h2osoi_liq, c, j + snl0);
        tsno[j] = CS(t_soisno, c, j + snl0);
        for (int k = 0; k < 8; ++k) ma[k][j] = MS(k, c, j + snl0);
        rds[j] = CS(snw_rds, c, j + snl0);
      }
    int msno = abs(snl0);
    int k = 1;
    while (k <= msno && k < NSNO) {
      if (k == msno) {
        if (dzsno[k] > dzmax_l[k - 1]) {
          msno = msno + 1;
          dzsno[k] = dzsno[k] / 2.0;
          dzsno[k + 1] = dzsno[k];
          swice[k] = swice[k] / 2.0; swice[k + 1] = swice[k];
          swliq[k] = swliq[k] / 2.0; swliq[k + 1] = swliq[k];
          if (k == 1) {
            tsno[k + 1] = tsno[k];
          } else {
            const double dtdz = (tsno[k - 1] - tsno[k]) / ((dzsno[k - 1] + 2 * dzsno[k]) / 2.0);
            tsno[k + 1] = tsno[k] - dtdz * dzsno[k] / 2.0;
            if (tsno[k + 1] >= tfrz) tsno[k + 1] = tsno[k];
            else tsno[k] = tsno[k] + dtdz * dzsno[k] / 2.0;
          }
          for (int a = 0; a < 8; ++a) { ma[a][k] = ma[a][k] / 2.0; ma[a][k + 1] = ma[a][k]; }
          rds[k + 1] = rds[k];
        }
      }
      if (k < msno) {
        if (dzsno[k] > dzmax_u[k - 1]) {
          const double drr = dzsno[k] - dzmax_u[k - 1] - 0.0;
          double propor = drr / dzsno[k];
          double zwice = propor * swice[k], zwliq = propor * swliq[k];
          double zm[8];
          for (int a = 0; a < 8; ++a) zm[a] = propor * ma[a][k];
          propor = (dzmax_u[k - 1] + 0.0) / dzsno[k];
          swice[k] = propor * swice[k];
          swliq[k] = propor * swliq[k];
          for (int a = 0; a < 8; ++a) ma[a][k] = propor * ma[a][k];
          dzsno[k] = dzmax_u[k - 1] + 0.0;
          for (int a = 0; a < 8; ++a) ma[a][k + 1] = ma[a][k + 1] + zm[a];
          rds[k + 1] = mass_weighted_snow_radius(prm, rds[k], rds[k + 1], (swliq[k + 1] + swice[k + 1]), (zwliq + zwice));
          combo(&dzsno[k + 1], &swliq[k + 1], &swice[k + 1], &tsno[k + 1], drr, zwliq, zwice, tsno[k]);
        }
      }
      k = k + 1;
    }
    CC(snl, c) = -msno;
    for (int j = SNO_LO; j <= 0; ++j)
      if (j >= CC(snl, c) + 1) {
        const int jj = j - CC(snl, c);
        CS(dz, c, j) = dzsno[jj] / frac_sno;
        CS(h2osoi_ice, c, j) = swice[jj];
        CS(h2osoi_liq, c, j) = swliq[jj];
        CS(t_soisno, c, j) = tsno[jj];
        for (int a = 0; a < 8; ++a) MS(a, c, j) = ma[a][jj];
        CS(snw_rds, c, j) = rds[jj];
      }
  }
  for (int j = 0; j >= SNO_LO; --j)                                                            /* :2883-2891 */
    for (int fc = 0; fc < num_snowc; ++fc) {
      const int c = filter_snowc[fc];
      if (j >= CC(snl, c) + 1) {
        CS(z, c, j) = CZ(zi, c, j) - 0.5 * CS(dz, c, j);
        CZ(zi, c, j - 1) = CZ(zi, c, j) - CS(dz, c, j);
      }
    }

  /* ---------------- ZeroEmptySnowLayers :2933-2948 ---------------- */
  for (int j = SNO_LO; j <= 0; ++j)
    for (int fc = 0; fc < num_snowc; ++fc) {
      const int c = filter_snowc[fc];
      if (j <= CC(snl, c) && CC(snl, c) > -NSNO) {
        CS(h2osoi_ice, c, j) = 0.0;
        CS(h2osoi_liq, c, j) = 0.0;
        CS(t_soisno, c, j) = 0.0;
        CS(dz, c, j) = 0.0;
        CS(z, c, j) = 0.0;
        CZ(zi, c, j - 1) = 0.0;
      }
    }
  return 0;
#undef MS
}

/* SnowCapping :3121-3247 with InitFlux_SnowCapping :3273-3283, CalculateTotalH2osno WaterStateType.F90:887-897,
 * SnowCappingExcess :3425-3472, BulkFlux_SnowCappingFluxes :3344-3392, UpdateState_RemoveSnowCappingFluxes :3601-3618,
 * SnowCappingUpdateDzAndAerosols :3668-3691.  Bulk water only. */
int oracle_snow_capping(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_initc, const int32_t* filter_initc,
                        int num_snowc, const int32_t* filter_snowc, const ctsm_snowcapping_fields_t* f, int nstep, ctsm_status_t* st) {
  (void)bounds;
  if (st) memset(st, 0, sizeof *st);
  const int begc0 = f->alloc.begc;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1);
  const double dtime = prm->dtime;
  const double reset_snow_h2osno = 35.0, min_snow_to_keep = 1.e-3;
  const int reset_snow_timesteps_per_layer = 4;
  double* mss[8] = {f->mss_bcphi, f->mss_bcpho, f->mss_ocphi, f->mss_ocpho, f->mss_dst1, f->mss_dst2, f->mss_dst3, f->mss_dst4};
  for (int fc = 0; fc < num_initc; ++fc) {
    const int c = filter_initc[fc];
    CC(qflx_snwcp_ice, c) = 0.0; CC(qflx_snwcp_liq, c) = 0.0; CC(qflx_snwcp_discarded_ice, c) = 0.0; CC(qflx_snwcp_discarded_liq, c) = 0.0;
  }
  int is_reset_snow_active = 0;
  if (prm->reset_snow || prm->reset_snow_glc) {
    const int reset_snow_timesteps = reset_snow_timesteps_per_layer * NSNO;
    if (nstep <= reset_snow_timesteps) is_reset_snow_active = 1;
  }
  for (int fc = 0; fc < num_snowc; ++fc) {
    const int c = filter_snowc[fc], lt = CC(lun_itype, c);
    double h2osno = CC(h2osno_no_layers, c);
    for (int j = CC(snl, c) + 1; j <= 0; ++j) h2osno = h2osno + CS(h2osoi_ice, c, j) + CS(h2osoi_liq, c, j);
    double excess = 0.0;
    int apply_runoff = 0;
    if (h2osno > prm->h2osno_max) { excess = h2osno - prm->h2osno_max; apply_runoff = 1; }
    if (is_reset_snow_active) {
      if ((lt != CTSM_ISTICE) && prm->reset_snow && (h2osno > reset_snow_h2osno)) {
        excess = h2osno - reset_snow_h2osno; apply_runoff = 0;
      } else if ((lt == CTSM_ISTICE) && prm->reset_snow_glc && (h2osno > reset_snow_h2osno) && (CC(topo, c) <= prm->reset_snow_glc_ela)) {
        excess = h2osno - reset_snow_h2osno; apply_runoff = 0;
      }
    }
    if (!(excess > 0.0)) continue;
    const double rho_orig_bottom = CS(h2osoi_ice, c, 0) / CS(dz, c, 0);
    const double mss_snow_bottom_lyr = CS(h2osoi_ice, c, 0) + CS(h2osoi_liq, c, 0);
    const double mss_snwcp_tot = fmin(excess, mss_snow_bottom_lyr * (1.0 - min_snow_to_keep));
    const double icefrac = CS(h2osoi_ice, c, 0) / mss_snow_bottom_lyr;
    const double snwcp_flux_ice = mss_snwcp_tot / dtime * icefrac;
    const double snwcp_flux_liq = mss_snwcp_tot / dtime * (1.0 - icefrac);
    if (apply_runoff) { CC(qflx_snwcp_ice, c) = snwcp_flux_ice; CC(qflx_snwcp_liq, c) = snwcp_flux_liq; }
    else { CC(qflx_snwcp_discarded_ice, c) = snwcp_flux_ice; CC(qflx_snwcp_discarded_liq, c) = snwcp_flux_liq; }
    const double frac_adjust = (mss_snow_bottom_lyr - mss_snwcp_tot) / mss_snow_bottom_lyr;
    CS(h2osoi_ice, c, 0) = CS(h2osoi_ice, c, 0) - (CC(qflx_snwcp_ice, c) + CC(qflx_snwcp_discarded_ice, c)) * dtime;
    CS(h2osoi_liq, c, 0) = CS(h2osoi_liq, c, 0) - (CC(qflx_snwcp_liq, c) + CC(qflx_snwcp_discarded_liq, c)) * dtime;
    if (CS(h2osoi_ice, c, 0) < 0.0 || CS(h2osoi_liq, c, 0) < 0.0) {
      if (st) {
        st->code = CTSM_ERR_SNOW_NEGATIVE; st->subgrid_level = CTSM_SUBGRID_COLUMN; st->subgrid_index = c; st->info = 3;
        snprintf(st->msg, sizeof st->msg, "ERROR: capping procedure failed (negative mass remaining)");
      }
      return CTSM_ERR_SNOW_NEGATIVE;
    }
    if (rho_orig_bottom > 1.0) CS(dz, c, 0) = CS(h2osoi_ice, c, 0) / rho_orig_bottom;
    for (int k = 0; k < 8; ++k) mss[k][(size_t)(0 - SNO_LO) * ldc + (c - begc0)] = mss[k][(size_t)(0 - SNO_LO) * ldc + (c - begc0)] * frac_adjust;
  }
  return 0;
}
