/* oracle_hydrology.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the surface-water / infiltration chain HydrologyNoDrainage runs between SnowWater and the root-water
 * sink (HydrologyNoDrainageMod.F90:297-337; first part of SURVEY.md section 8f rank 3):
 *   SetSoilWaterFractions      SoilHydrologyMod.F90:202-256      SetFloodc                :259-295
 *   SaturatedExcessRunoff      SaturatedExcessRunoffMod.F90:203-312 with ComputeFsatTopmodel :315-360
 *   SetQflxInputs              SoilHydrologyMod.F90:298-366
 *   InfiltrationExcessRunoff   InfiltrationExcessRunoffMod.F90:196-263 with ComputeQinmaxHksat :266-304
 *   RouteInfiltrationExcess    SoilHydrologyMod.F90:369-423
 *   UpdateH2osfc               SurfaceWaterMod.F90:345-433 with QflxH2osfcSurf :436-505, QflxH2osfcDrain :508-556,
 *                              truncate_small_values NumericsMod.F90:50
 *   Infiltration               SoilHydrologyMod.F90:426-457      TotalSurfaceRunoff       :460-550
 * Configuration: non-urban columns, use_excess_ice = use_vichydro = .false., no hillslope columns, fsat_method = TOPModel,
 * qinmax_method = hksat.  One loop per reference loop, in the reference's order.
 * Further down: oracle_water_table (PerchedWaterTable SoilHydrologyMod.F90:1525, ThetaBasedWaterTable :1933, RenewCondensation :2569)
 * and oracle_hydrology_diagnostics (the inline tail of HydrologyNoDrainage, HydrologyNoDrainageMod.F90:420-757).
 * Pinned by the NumPy restatements in tests/test_oracle_hydrology.py.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "oracle_pert.h"

static const double denice = 0.917e3, rpi = 3.14159265358979323846;

int oracle_hydrology_infiltration(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_nolakec,
                                  const int32_t* filter_nolakec, int num_hydrologyc, const int32_t* filter_hydrologyc,
                                  int num_urbanc, const ctsm_infiltration_fields_t* f, ctsm_status_t* st) {
  if (st) memset(st, 0, sizeof *st);
  const int begc0 = f->alloc.begc, begg0 = f->alloc.begg;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1);
  const double dtime = prm->dtime;
#define CC(name, c) f->name[(c) - begc0]
#define C2(name, c, j, lo) f->name[(size_t)((j) - (lo)) * ldc + ((c) - begc0)]
#define SNO_LO (-CTSM_NLEVSNO + 1)
  if (num_urbanc != 0) { if (st) st->code = CTSM_ERR_URBAN; return CTSM_ERR_URBAN; }
  for (int pass = 0; pass < 2; ++pass) {
    const int n = pass ? num_hydrologyc : num_nolakec;
    const int32_t* flt = pass ? filter_hydrologyc : filter_nolakec;
    for (int fc = 0; fc < n; ++fc) {
      const int lt = CC(lun_itype, flt[fc]);
      if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) {
        if (st) { st->code = CTSM_ERR_URBAN; st->subgrid_index = flt[fc]; }
        return CTSM_ERR_URBAN;
      }
    }
  }
  const int nb = bounds->endc - bounds->begc + 1;
  double* h2osfc_partial = (double*)calloc((size_t)(nb > 0 ? nb : 1), sizeof(double));
#define HP(c) h2osfc_partial[(c) - bounds->begc]

  /* SetSoilWaterFractions :239-252 (excess_ice = 0) */
  for (int j = 1; j <= CTSM_NLEVSOI; ++j)
    for (int fc = 0; fc < num_hydrologyc; ++fc) {
      const int c = filter_hydrologyc[fc];
      const double dz_ext = C2(dz, c, j, SNO_LO) + 0.0 / denice;
      const double vol_ice = fmin(C2(watsat, c, j, 1), (C2(h2osoi_ice, c, j, SNO_LO) + 0.0) / (dz_ext * denice));
      C2(eff_porosity, c, j, 1) = fmax(0.01, C2(watsat, c, j, 1) - vol_ice);
      C2(icefrac, c, j, 1) = fmin(1.0, vol_ice / C2(watsat, c, j, 1));
    }
  /* SetFloodc :282-291 */
  for (int fc = 0; fc < num_nolakec; ++fc) {
    const int c = filter_nolakec[fc];
    CC(qflx_floodc, c) = f->forc_flood[CC(col_gridcell, c) - begg0];
  }
  /* SaturatedExcessRunoff: ComputeFsatTopmodel :344-356, crop switch :254-260, :277-281 */
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc];
    if (CC(frost_table, c) > CC(zwt_perched, c) && CC(frost_table, c) <= CC(zwt, c))
      CC(fsat, c) = CC(wtfact, c) * exp(-0.5 * prm->fff * CC(zwt_perched, c));
    else
      CC(fsat, c) = CC(wtfact, c) * exp(-0.5 * prm->fff * CC(zwt, c));
  }
  if (prm->crop_fsat_equals_zero)
    for (int fc = 0; fc < num_hydrologyc; ++fc) {
      const int c = filter_hydrologyc[fc];
      if (CC(lun_itype, c) == CTSM_ISTCROP) CC(fsat, c) = 0.0;
    }
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc];
    CC(qflx_sat_excess_surf, c) = CC(fsat, c) * CC(qflx_rain_plus_snomelt, c);
    CC(fcov, c) = CC(fsat, c);
  }
  /* SetQflxInputs :339-362 */
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc];
    CC(qflx_top_soil, c) = CC(qflx_rain_plus_snomelt, c) + CC(qflx_snow_h2osfc, c) + CC(qflx_floodc, c);
    double fsno, qflx_evap;
    if (CC(snl, c) >= 0) { fsno = 0.0; qflx_evap = CC(qflx_liqevap_from_top_layer, c); }
    else { fsno = CC(frac_sno_eff, c); qflx_evap = CC(qflx_ev_soil_col, c); }
    CC(qflx_in_soil, c) = (1.0 - CC(frac_h2osfc, c)) * (CC(qflx_top_soil, c) - CC(qflx_sat_excess_surf, c));
    CC(qflx_top_soil_to_h2osfc, c) = CC(frac_h2osfc, c) * (CC(qflx_top_soil, c) - CC(qflx_sat_excess_surf, c));
    CC(qflx_in_soil, c) = CC(qflx_in_soil, c) - (1.0 - fsno - CC(frac_h2osfc, c)) * qflx_evap;
    CC(qflx_top_soil_to_h2osfc, c) = CC(qflx_top_soil_to_h2osfc, c) - CC(frac_h2osfc, c) * CC(qflx_ev_h2osfc_col, c);
  }
  /* InfiltrationExcessRunoff: ComputeQinmaxHksat :296-300 (minval over levels 1..3), :253-259 */
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc];
    double q = pow(10.0, -prm->e_ice * (C2(icefrac, c, 1, 1))) * C2(hksat, c, 1, 1);
    for (int j = 2; j <= 3; ++j) {
      const double v = pow(10.0, -prm->e_ice * (C2(icefrac, c, j, 1))) * C2(hksat, c, j, 1);
      if (v < q) q = v;
    }
    CC(qinmax, c) = (1.0 - CC(fsat, c)) * q;
    CC(qflx_infl_excess, c) = fmax(0.0, (CC(qflx_in_soil, c) - (1.0 - CC(frac_h2osfc, c)) * CC(qinmax, c)));
  }
  /* RouteInfiltrationExcess :399-419 */
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc], lt = CC(lun_itype, c);
    if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP) {
      CC(qflx_in_soil_limited, c) = CC(qflx_in_soil, c) - CC(qflx_infl_excess, c);
      if (prm->h2osfcflag != 0) {
        CC(qflx_in_h2osfc, c) = CC(qflx_top_soil_to_h2osfc, c) + CC(qflx_infl_excess, c);
        CC(qflx_infl_excess_surf, c) = 0.0;
      } else {
        CC(qflx_in_h2osfc, c) = CC(qflx_top_soil_to_h2osfc, c);
        CC(qflx_infl_excess_surf, c) = CC(qflx_infl_excess, c);
      }
    } else {
      CC(qflx_in_soil_limited, c) = CC(qflx_in_soil, c);
      CC(qflx_in_h2osfc, c) = 0.0;
      CC(qflx_infl_excess_surf, c) = 0.0;
    }
  }
  /* UpdateH2osfc: QflxH2osfcSurf SurfaceWaterMod.F90:468-501 */
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc];
    double frac_infclust = 0.0;
    if (prm->h2osfcflag == 1) {
      if (CC(frac_h2osfc_nosnow, c) <= prm->pc) frac_infclust = 0.0;
      else frac_infclust = pow(CC(frac_h2osfc_nosnow, c) - prm->pc, prm->mu);
    }
    if (CC(h2osfc, c) > CC(h2osfc_thresh, c) && prm->h2osfcflag != 0) {
      const double k_wet = 1.0e-4 * sin((rpi / 180.0) * CC(topo_slope, c));
      CC(qflx_h2osfc_surf, c) = k_wet * frac_infclust * (CC(h2osfc, c) - CC(h2osfc_thresh, c));
      CC(qflx_h2osfc_surf, c) = fmin(CC(qflx_h2osfc_surf, c), (CC(h2osfc, c) - CC(h2osfc_thresh, c)) / dtime);
    } else {
      CC(qflx_h2osfc_surf, c) = 0.0;
    }
    if (CC(qflx_h2osfc_surf, c) < (double)1.0e-8f) CC(qflx_h2osfc_surf, c) = 0.0;     /* :499, a default-kind literal: REAL(4) 1.0e-8 */
  }
  for (int fc = 0; fc < num_hydrologyc; ++fc) {                       /* :394-397 */
    const int c = filter_hydrologyc[fc];
    HP(c) = CC(h2osfc, c) + (CC(qflx_in_h2osfc, c) - CC(qflx_h2osfc_surf, c)) * dtime;
  }
  oracle_truncate_small_values(num_hydrologyc, filter_hydrologyc, bounds->begc, &CC(h2osfc, bounds->begc), h2osfc_partial, 1.e-13);
  for (int fc = 0; fc < num_hydrologyc; ++fc) {                       /* QflxH2osfcDrain :541-552 */
    const int c = filter_hydrologyc[fc];
    if (HP(c) < 0.0) {
      CC(qflx_h2osfc_drain, c) = HP(c) / dtime;
    } else {
      CC(qflx_h2osfc_drain, c) = fmin(CC(frac_h2osfc, c) * CC(qinmax, c), HP(c) / dtime);
      if (prm->h2osfcflag == 0) CC(qflx_h2osfc_drain, c) = fmax(0.0, HP(c) / dtime);
    }
  }
  {
    double* base = (double*)malloc(sizeof(double) * (size_t)(nb > 0 ? nb : 1));
    memcpy(base, h2osfc_partial, sizeof(double) * (size_t)(nb > 0 ? nb : 1));
    for (int fc = 0; fc < num_hydrologyc; ++fc) {                     /* :421-424 */
      const int c = filter_hydrologyc[fc];
      CC(h2osfc, c) = HP(c) - CC(qflx_h2osfc_drain, c) * dtime;
    }
    oracle_truncate_small_values(num_hydrologyc, filter_hydrologyc, bounds->begc, base, &CC(h2osfc, bounds->begc), 1.e-13);
    free(base);
  }
  /* Infiltration :450-453, TotalSurfaceRunoff :511-515 */
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc];
    CC(qflx_infl, c) = CC(qflx_in_soil_limited, c) + CC(qflx_h2osfc_drain, c);
  }
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc];
    CC(qflx_surf, c) = CC(qflx_sat_excess_surf, c) + CC(qflx_infl_excess_surf, c) + CC(qflx_h2osfc_surf, c);
  }
  free(h2osfc_partial);
  return 0;
#undef CC
#undef C2
#undef HP
}

/* ---------------------------------------------------------------------------------------------------------------------
 * PerchedWaterTable SoilHydrologyMod.F90:1525-1641, ThetaBasedWaterTable :1933-2025, RenewCondensation :2569-2678
 * (HydrologyNoDrainageMod.F90:359-373, use_aquifer_layer = .false.).  Both water-table routines compare against
 * sat_lev = 0.9 written as a default-kind literal: REAL(4) 0.9 promoted to double, kept. */
int oracle_water_table(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_hydrologyc, const int32_t* filter_hydrologyc,
                       int num_urbanc, const ctsm_watertable_fields_t* f, ctsm_status_t* st) {
  (void)bounds;
  if (st) memset(st, 0, sizeof *st);
  const int begc0 = f->alloc.begc;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1);
  const double dtime = prm->dtime, tfrz = 273.15, denh2o = 1.000e3;
  const double sat_lev = (double)0.9f;
#define CC(name, c) f->name[(c) - begc0]
#define C2(name, c, j, lo) f->name[(size_t)((j) - (lo)) * ldc + ((c) - begc0)]
#define SNO_LO (-CTSM_NLEVSNO + 1)
#define ZI(c, j) C2(zi, c, j, -CTSM_NLEVSNO)
#define LIQ(c, j) C2(h2osoi_liq, c, j, SNO_LO)
#define ICE(c, j) C2(h2osoi_ice, c, j, SNO_LO)
#define DZ(c, j) C2(dz, c, j, SNO_LO)
#define ZZ(c, j) C2(z, c, j, SNO_LO)
#define TS(c, j) C2(t_soisno, c, j, SNO_LO)
  if (num_urbanc != 0) { if (st) st->code = CTSM_ERR_URBAN; return CTSM_ERR_URBAN; }
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int lt = CC(lun_itype, filter_hydrologyc[fc]);
    if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) {
      if (st) { st->code = CTSM_ERR_URBAN; st->subgrid_index = filter_hydrologyc[fc]; }
      return CTSM_ERR_URBAN;
    }
  }
  /* PerchedWaterTable :1581-1637 */
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc];
    int k_frz, k_perch;
    if (TS(c, 1) > tfrz) k_frz = CTSM_NLEVSOI; else k_frz = 1;
    for (int k = 2; k <= CTSM_NLEVSOI; ++k)
      if (TS(c, k - 1) > tfrz && TS(c, k) <= tfrz) { k_frz = k; break; }
    CC(frost_table, c) = ZI(c, k_frz - 1);
    CC(zwt_perched, c) = CC(frost_table, c);
    if (CC(zwt, c) < CC(frost_table, c) && TS(c, k_frz) <= tfrz) {
      /* water table above the frost table: nothing */
    } else if (k_frz > 1) {
      k_perch = 1;
      for (int k = k_frz; k >= 1; --k) {
        C2(h2osoi_vol, c, k, 1) = LIQ(c, k) / (DZ(c, k) * denh2o) + ICE(c, k) / (DZ(c, k) * denice);
        if (C2(h2osoi_vol, c, k, 1) / C2(watsat, c, k, 1) <= sat_lev) { k_perch = k; break; }
      }
      if (TS(c, k_frz) > tfrz) k_perch = k_frz;
      if (k_frz > k_perch) {
        const double s1 = (LIQ(c, k_perch) / (DZ(c, k_perch) * denh2o) + ICE(c, k_perch) / (DZ(c, k_perch) * denice)) / C2(watsat, c, k_perch, 1);
        const double s2 = (LIQ(c, k_perch + 1) / (DZ(c, k_perch + 1) * denh2o) + ICE(c, k_perch + 1) / (DZ(c, k_perch + 1) * denice)) /
                          C2(watsat, c, k_perch + 1, 1);
        if (s1 > s2) {
          CC(zwt_perched, c) = ZI(c, k_perch - 1);
        } else {
          const double m = (ZZ(c, k_perch + 1) - ZZ(c, k_perch)) / (s2 - s1);
          const double b = ZZ(c, k_perch + 1) - m * s2;
          CC(zwt_perched, c) = fmax(0.0, m * sat_lev + b);
        }
      }
    }
  }
  /* ThetaBasedWaterTable :1974-2021 */
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc];
    const int nb = CC(nbedrock, c);
    CC(zwt, c) = ZI(c, CTSM_NLEVSOI);
    int k_zwt = nb, sat_flag = 1;
    for (int k = nb; k >= 1; --k) {
      C2(h2osoi_vol, c, k, 1) = LIQ(c, k) / (DZ(c, k) * denh2o) + ICE(c, k) / (DZ(c, k) * denice);
      if (C2(h2osoi_vol, c, k, 1) / C2(watsat, c, k, 1) <= sat_lev) { k_zwt = k; sat_flag = 0; break; }
    }
    if (sat_flag == 1) k_zwt = 1;
    if (k_zwt == 1) {
      CC(zwt, c) = ZI(c, 1);
    } else if (k_zwt < nb) {
      const double s1 = (LIQ(c, k_zwt) / (DZ(c, k_zwt) * denh2o) + ICE(c, k_zwt) / (DZ(c, k_zwt) * denice)) / C2(watsat, c, k_zwt, 1);
      const double s2 = (LIQ(c, k_zwt + 1) / (DZ(c, k_zwt + 1) * denh2o) + ICE(c, k_zwt + 1) / (DZ(c, k_zwt + 1) * denice)) /
                        C2(watsat, c, k_zwt + 1, 1);
      const double m = (ZZ(c, k_zwt + 1) - ZZ(c, k_zwt)) / (s2 - s1);
      const double b = ZZ(c, k_zwt + 1) - m * s2;
      CC(zwt, c) = fmax(0.0, m * sat_lev + b);
    } else {
      CC(zwt, c) = ZI(c, nb);
    }
  }
  /* RenewCondensation :2612-2674 (tolerance = 1e-12, :71) */
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc];
    if (CC(snl, c) + 1 >= 1) {
      LIQ(c, 1) = LIQ(c, 1) + (1.0 - CC(frac_h2osfc, c)) * CC(qflx_liqdew_to_top_layer, c) * dtime;
      ICE(c, 1) = ICE(c, 1) + (1.0 - CC(frac_h2osfc, c)) * CC(qflx_soliddew_to_top_layer, c) * dtime;
      const double before = ICE(c, 1);
      ICE(c, 1) = ICE(c, 1) - (1.0 - CC(frac_h2osfc, c)) * CC(qflx_solidevap_from_top_layer, c) * dtime;
      if (fabs(ICE(c, 1)) < 1.e-12 * fabs(before)) ICE(c, 1) = 0.0;
    }
  }
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    const int c = filter_hydrologyc[fc];
    if (CC(snl, c) + 1 >= 1 && ICE(c, 1) < 0.0) {
      if (st) {
        st->code = CTSM_ERR_SNOW_NEGATIVE; st->subgrid_level = CTSM_SUBGRID_COLUMN; st->subgrid_index = c; st->info = 2;
        st->value = ICE(c, 1);
        snprintf(st->msg, sizeof st->msg, "In RenewCondensation, h2osoi_ice has gone significantly negative");
      }
      return CTSM_ERR_SNOW_NEGATIVE;
    }
  }
  return 0;
}

/* The inline tail of HydrologyNoDrainage, HydrologyNoDrainageMod.F90:420-757 */
int oracle_hydrology_diagnostics(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                                 int num_snowc, const int32_t* filter_snowc, int num_nosnowc, const int32_t* filter_nosnowc,
                                 int num_hydrologyc, const int32_t* filter_hydrologyc, int num_urbanc,
                                 const ctsm_hydrodiag_fields_t* f, ctsm_status_t* st) {
  if (st) memset(st, 0, sizeof *st);
  const int begc0 = f->alloc.begc;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1);
  const double dtime = prm->dtime, tfrz = 273.15, denh2o = 1.000e3, spval = 1.e36;
  if (num_urbanc != 0) { if (st) st->code = CTSM_ERR_URBAN; return CTSM_ERR_URBAN; }
  for (int fc = 0; fc < num_nolakec; ++fc) {
    const int lt = CC(lun_itype, filter_nolakec[fc]);
    if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) {
      if (st) { st->code = CTSM_ERR_URBAN; st->subgrid_index = filter_nolakec[fc]; }
      return CTSM_ERR_URBAN;
    }
  }
  for (int fc = 0; fc < num_snowc; ++fc) { const int c = filter_snowc[fc]; CC(snow_persistence, c) = CC(snow_persistence, c) + dtime; }   /* :420-427 */
  for (int fc = 0; fc < num_nosnowc; ++fc) CC(snow_persistence, filter_nosnowc[fc]) = 0.0;
  for (int fc = 0; fc < num_nolakec; ++fc) { const int c = filter_nolakec[fc]; CC(snowice, c) = 0.0; CC(snowliq, c) = 0.0; }           /* :432-447 */
  for (int j = SNO_LO; j <= 0; ++j)
    for (int fc = 0; fc < num_snowc; ++fc) {
      const int c = filter_snowc[fc];
      if (j >= CC(snl, c) + 1) { CC(snowice, c) = CC(snowice, c) + ICE(c, j); CC(snowliq, c) = CC(snowliq, c) + LIQ(c, j); }
    }
  for (int c = bounds->begc; c <= bounds->endc; ++c) CC(snowdp, c) = CC(snow_depth, c) * CC(frac_sno_eff, c);                       /* :452 */
  for (int fc = 0; fc < num_nolakec; ++fc) CC(t_sno_mul_mss, filter_nolakec[fc]) = 0.0;                                           /* :458-472 */
  for (int j = SNO_LO; j <= 0; ++j)
    for (int fc = 0; fc < num_snowc; ++fc) {
      const int c = filter_snowc[fc];
      if (j >= CC(snl, c) + 1) {
        CC(t_sno_mul_mss, c) = CC(t_sno_mul_mss, c) + ICE(c, j) * TS(c, j);
        CC(t_sno_mul_mss, c) = CC(t_sno_mul_mss, c) + LIQ(c, j) * tfrz;
      }
    }
  for (int fc = 0; fc < num_nolakec; ++fc) { const int c = filter_nolakec[fc]; CC(t_soi10cm, c) = 0.0; CC(t_soi17cm, c) = 0.0; }    /* :477-522 */
  for (int j = 1; j <= CTSM_NLEVSOI; ++j)
    for (int fc = 0; fc < num_nolakec; ++fc) {
      const int c = filter_nolakec[fc];
      if (j == 1) CC(tsl, c) = TS(c, j);
      if (ZI(c, j) <= 0.17) {
        const double fracl = 1.0;
        CC(t_soi17cm, c) = CC(t_soi17cm, c) + TS(c, j) * DZ(c, j) * fracl;
      } else if (ZI(c, j) > 0.17 && ZI(c, j - 1) < 0.17) {
        const double fracl = (0.17 - ZI(c, j - 1)) / DZ(c, j);
        CC(t_soi17cm, c) = CC(t_soi17cm, c) + TS(c, j) * DZ(c, j) * fracl;
      }
      if (ZI(c, j) <= 0.1) {
        const double fracl = 1.0;
        CC(t_soi10cm, c) = CC(t_soi10cm, c) + TS(c, j) * DZ(c, j) * fracl;
      } else if (ZI(c, j) > 0.1 && ZI(c, j - 1) < 0.1) {
        const double fracl = (0.1 - ZI(c, j - 1)) / DZ(c, j);
        CC(t_soi10cm, c) = CC(t_soi10cm, c) + TS(c, j) * DZ(c, j) * fracl;
      }
    }
  for (int fc = 0; fc < num_nolakec; ++fc) {                                                                                     /* :524-553 */
    const int c = filter_nolakec[fc], lt = CC(lun_itype, c);
    if (CC(snl, c) < 0)
      CC(t_grnd, c) = CC(frac_sno_eff, c) * TS(c, CC(snl, c) + 1) + (1.0 - CC(frac_sno_eff, c) - CC(frac_h2osfc, c)) * TS(c, 1) +
                      CC(frac_h2osfc, c) * CC(t_h2osfc, c);
    else
      CC(t_grnd, c) = (1.0 - CC(frac_h2osfc, c)) * TS(c, 1) + CC(frac_h2osfc, c) * CC(t_h2osfc, c);
    CC(t_soi10cm, c) = CC(t_soi10cm, c) / 0.1;
    CC(t_soi17cm, c) = CC(t_soi17cm, c) / 0.17;
    if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP) CC(t_grnd_r, c) = TS(c, CC(snl, c) + 1);
  }
  for (int j = 1; j <= CTSM_NLEVGRND; ++j)                                                                                       /* :561-570 */
    for (int fc = 0; fc < num_nolakec; ++fc) {
      const int c = filter_nolakec[fc];
      C2(h2osoi_vol, c, j, 1) = LIQ(c, j) / (DZ(c, j) * denh2o) + ICE(c, j) / (DZ(c, j) * denice);
    }
  for (int j = 1; j <= CTSM_NLEVGRND; ++j)                                                                                       /* :598-617 */
    for (int fc = 0; fc < num_hydrologyc; ++fc) {
      const int c = filter_hydrologyc[fc];
      if (LIQ(c, j) > 0.0) {
        const double vwc = LIQ(c, j) / (DZ(c, j) * denh2o);
        const double fsattmp = fmax(vwc / C2(watsat, c, j, 1), 0.001);
        const double psi = C2(sucsat, c, j, 1) * (-9.8e-6) * pow(fsattmp, -C2(bsw, c, j, 1));
        C2(soilpsi, c, j, 1) = fmin(fmax(psi, -15.0), 0.0);
      } else {
        C2(soilpsi, c, j, 1) = -15.0;
      }
    }
  for (int j = 1; j <= CTSM_NLEVGRND; ++j)                                                                                       /* :624-634 */
    for (int fc = 0; fc < num_hydrologyc; ++fc) {
      const int c = filter_hydrologyc[fc];
      double s_node = fmax(C2(h2osoi_vol, c, j, 1) / C2(watsat, c, j, 1), 0.01);
      s_node = fmin(1.0, s_node);
      C2(smp_l, c, j, 1) = -C2(sucsat, c, j, 1) * pow(s_node, -C2(bsw, c, j, 1));
      C2(smp_l, c, j, 1) = fmax(CC(smpmin, c), C2(smp_l, c, j, 1));
    }
  {                                                                                    /* wf, wf2 :641-733: rwat / swat / rz are NOT reset between the two */
    const int nb = bounds->endc - bounds->begc + 1;
    double* rwat = (double*)calloc((size_t)(nb > 0 ? nb : 1) * 3, sizeof(double));
    double *swat = rwat + (nb > 0 ? nb : 1), *rz = swat + (nb > 0 ? nb : 1);
    for (int pass = 0; pass < 2; ++pass) {
      const double depth = pass ? 0.17 : 0.05;
      for (int j = 1; j <= CTSM_NLEVGRND; ++j)
        for (int fc = 0; fc < num_hydrologyc; ++fc) {
          const int c = filter_hydrologyc[fc], k = c - bounds->begc;
          if (ZZ(c, j) + 0.5 * DZ(c, j) <= depth) {
            const double watdry = C2(watsat, c, j, 1) * pow(316230.0 / C2(sucsat, c, j, 1), -1.0 / C2(bsw, c, j, 1));
            rwat[k] = rwat[k] + (C2(h2osoi_vol, c, j, 1) - watdry) * DZ(c, j);
            swat[k] = swat[k] + (C2(watsat, c, j, 1) - watdry) * DZ(c, j);
            rz[k] = rz[k] + DZ(c, j);
          }
        }
      for (int fc = 0; fc < num_hydrologyc; ++fc) {
        const int c = filter_hydrologyc[fc], k = c - bounds->begc;
        double tsw, stsw;
        if (rz[k] != 0.0) {
          tsw = rwat[k] / rz[k];
          stsw = swat[k] / rz[k];
        } else {
          const double watdry = C2(watsat, c, 1, 1) * pow(316230.0 / C2(sucsat, c, 1, 1), -1.0 / C2(bsw, c, 1, 1));
          tsw = C2(h2osoi_vol, c, 1, 1) - watdry;
          stsw = C2(watsat, c, 1, 1) - watdry;
        }
        if (pass) CC(wf2, c) = tsw / stsw; else CC(wf, c) = tsw / stsw;
      }
    }
    free(rwat);
  }
  for (int fc = 0; fc < num_snowc; ++fc) {                                                                                       /* :738-754 */
    const int c = filter_snowc[fc];
    CC(h2osno_top, c) = ICE(c, CC(snl, c) + 1) + LIQ(c, CC(snl, c) + 1);
  }
  for (int fc = 0; fc < num_nosnowc; ++fc) {
    const int c = filter_nosnowc[fc];
    CC(h2osno_top, c) = 0.0;
    for (int j = SNO_LO; j <= 0; ++j) C2(snw_rds, c, j, SNO_LO) = 0.0;
    CC(snot_top, c) = spval; CC(dTdz_top, c) = spval; CC(snw_rds_top, c) = spval; CC(sno_liq_top, c) = spval;
  }
  return 0;
#undef CC
#undef C2
}
