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
 */
#include <math.h>
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
    if (CC(qflx_h2osfc_surf, c) < 1.0e-8) CC(qflx_h2osfc_surf, c) = 0.0;
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
