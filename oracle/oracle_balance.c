/* oracle_balance.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of
 *   BalanceCheck            src/biogeophys/BalanceCheckMod.F90:445-857
 *   EnergyBalanceCheck      :859-1119
 *   c2g_1d ('urbanf','unity' on non-urban landunits)   src/main/subgridAveMod.F90:761-818
 *   CalculateTotalH2osno    src/biogeophys/WaterStateType.F90:887-896
 *   Compute_EffecRootFrac_And_VertTranSink_HydStress   src/biogeophys/SoilWaterPlantSinkMod.F90:236-328
 * for bulk water, non-urban landunits, use_hillslope_routing = .false..  The skip-step rule
 * (BalanceCheckInit :74-95) is pinned by the reference's test_Balance.pf (tests/test_oracle_golden.py);
 * the residual formulas are PARITY UNPINNED by the reference's tests.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define NLEVSNO CTSM_NLEVSNO
#define NLEVSOI CTSM_NLEVSOI
#define SNOSOI_LO (-NLEVSNO + 1)
static const double spval = 1.e36;
static const double error_thresh = 1.e-5, h2o_warning_thresh = 1.e-9, energy_warning_thresh = 1.e-7;

/* maxval / maxloc over [lo,hi] with optional mask (mask==NULL: all); arrays are offset by base */
static void maxabs(const double* a, int base, int lo, int hi, const int* mask_ok, double* mx, int* loc) {
  *mx = 0.0; *loc = 0;
  int any = 0;
  for (int i = lo; i <= hi; ++i) {
    if (mask_ok && !mask_ok[i - base]) continue;
    const double v = fabs(a[i - base]);
    if (!any || v > *mx) { *mx = v; *loc = i; any = 1; }
  }
  if (!any) { *mx = 0.0; *loc = 0; }
}

int oracle_balancecheck(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_allc, const int32_t* filter_allc,
                        const ctsm_balancecheck_fields_t* f, int DAnstep, ctsm_balance_report_t* rep, ctsm_status_t* st) {
  const int begc0 = f->alloc.begc, begp0 = f->alloc.begp, begg0 = f->alloc.begg;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1), ldg = (size_t)(f->alloc.endg - f->alloc.begg + 1);
  const double dtime = prm->dtime;
  const int skip_steps = prm->balance_skip_steps;
  memset(rep, 0, sizeof *rep);
  if (st) memset(st, 0, sizeof *st);
  rep->abort_kind = -1;
  rep->skip_steps = skip_steps;
  if (skip_steps <= 0) return CTSM_ERR_BAD_ARG;
#define CC(name, c) (f->name[(c) - begc0])
#define GG(name, g) (f->name[(g) - begg0])
#define PP(name, p) (f->name[(p) - begp0])
  const int nc = bounds->endc - bounds->begc + 1, ng = bounds->endg - bounds->begg + 1, np = bounds->endp - bounds->begp + 1;
  double* h2osno_total = (double*)calloc((size_t)(nc > 0 ? nc : 1), sizeof(double));
  double* gdyn = (double*)calloc((size_t)(ng > 0 ? ng : 1), sizeof(double));
  double* gliq = (double*)calloc((size_t)(ng > 0 ? ng : 1), sizeof(double));
  double* gice = (double*)calloc((size_t)(ng > 0 ? ng : 1), sizeof(double));
  int* okmask = (int*)calloc((size_t)((np > nc ? np : nc) + 1), sizeof(int));
  int abort_kind = -1, abort_index = 0, abort_level = 0;

  for (int c = bounds->begc; c <= bounds->endc; ++c) {              /* :592-608 */
    if (CC(col_active, c)) {
      CC(errh2o, c) = CC(endwb, c) - CC(begwb, c)
          - (CC(forc_rain, c) + CC(forc_snow, c) + CC(qflx_flood, c) + CC(qflx_sfc_irrig, c) + CC(qflx_glcice_dyn_water_flux, c)
             - CC(qflx_evap_tot, c) - CC(qflx_surf, c) - CC(qflx_qrgwl, c) - CC(qflx_drain, c) - CC(qflx_drain_perched, c)
             - CC(qflx_ice_runoff, c) - CC(qflx_snwcp_discarded_liq, c) - CC(qflx_snwcp_discarded_ice, c)) * dtime;
    } else {
      CC(errh2o, c) = 0.0;
    }
  }
  maxabs(f->errh2o, begc0, bounds->begc, bounds->endc, NULL, &rep->max_abs[CTSM_BAL_H2O_COL], &rep->index[CTSM_BAL_H2O_COL]);
  if (rep->max_abs[CTSM_BAL_H2O_COL] > h2o_warning_thresh) {         /* :612-660 */
    rep->warn[CTSM_BAL_H2O_COL] = 1;
    if (rep->max_abs[CTSM_BAL_H2O_COL] > error_thresh && DAnstep > skip_steps && abort_kind < 0) {
      abort_kind = CTSM_BAL_H2O_COL; abort_index = rep->index[CTSM_BAL_H2O_COL]; abort_level = CTSM_SUBGRID_COLUMN;
    }
  }
  /* c2g of three column fluxes, subgridAveMod.F90:791-812 (unity scales) */
  const double* carr[3] = {f->qflx_glcice_dyn_water_flux, f->qflx_snwcp_discarded_liq, f->qflx_snwcp_discarded_ice};
  double* garr[3] = {gdyn, gliq, gice};
  for (int k = 0; k < 3; ++k) {
    for (int g = bounds->begg; g <= bounds->endg; ++g) {
      double sumwt = 0.0, acc = spval;
      for (int c = GG(grc_coli, g); c <= GG(grc_colf, g); ++c) {
        if (CC(col_active, c) && CC(wtgcell, c) != 0.0) {
          if (carr[k][c - begc0] != spval) {
            if (sumwt == 0.0) acc = 0.0;
            acc = acc + carr[k][c - begc0] * 1.0 * 1.0 * CC(wtgcell, c);
            sumwt = sumwt + CC(wtgcell, c);
          }
        }
      }
      if (!(sumwt > 1.0 + 1.e-6) && sumwt != 0.0) acc = acc / sumwt;
      garr[k][g - bounds->begg] = acc;
    }
  }
  for (int g = bounds->begg; g <= bounds->endg; ++g) {              /* :679-694 */
    const int i = g - bounds->begg;
    GG(errh2o_grc, g) = GG(endwb_grc, g) - GG(begwb_grc, g)
        - (GG(forc_rain_grc, g) + GG(forc_snow_grc, g) + GG(forc_flood_grc, g) + GG(qflx_sfc_irrig_grc, g) + gdyn[i]
           - GG(qflx_evap_tot_grc, g) - GG(qflx_surf_grc, g) - GG(qflx_qrgwl_grc, g) - GG(qflx_drain_grc, g)
           - GG(qflx_drain_perched_grc, g) - GG(qflx_ice_runoff_grc, g) - gliq[i] - gice[i]) * dtime;
  }
  maxabs(f->errh2o_grc, begg0, bounds->begg, bounds->endg, NULL, &rep->max_abs[CTSM_BAL_H2O_GRC], &rep->index[CTSM_BAL_H2O_GRC]);
  if (rep->max_abs[CTSM_BAL_H2O_GRC] > h2o_warning_thresh) {         /* :709-748 */
    rep->warn[CTSM_BAL_H2O_GRC] = 1;
    if (rep->max_abs[CTSM_BAL_H2O_GRC] > error_thresh && DAnstep > skip_steps && abort_kind < 0) {
      abort_kind = CTSM_BAL_H2O_GRC; abort_index = rep->index[CTSM_BAL_H2O_GRC]; abort_level = CTSM_SUBGRID_GRIDCELL;
    }
  }
  /* CalculateTotalH2osno over filter_allc, WaterStateType.F90:887-896 */
  for (int fc = 0; fc < num_allc; ++fc) {
    const int c = filter_allc[fc];
    double t = CC(h2osno_no_layers, c);
    for (int j = CC(snl, c) + 1; j <= 0; ++j)
      t = t + f->h2osoi_ice[(size_t)(j - SNOSOI_LO) * ldc + (c - begc0)] + f->h2osoi_liq[(size_t)(j - SNOSOI_LO) * ldc + (c - begc0)];
    h2osno_total[c - bounds->begc] = t;
  }
  for (int c = bounds->begc; c <= bounds->endc; ++c) {              /* :754-803 */
    if (CC(col_active, c)) {
      const int lt = CC(lun_itype, c);
      if (CC(snl, c) < 0) {
        double src = CC(qflx_prec_grnd, c) + CC(qflx_soliddew_to_top_layer, c) + CC(qflx_liqdew_to_top_layer, c);
        double snk = CC(qflx_solidevap_from_top_layer, c) + CC(qflx_liqevap_from_top_layer, c) + CC(qflx_snow_drain, c)
                     + CC(qflx_snwcp_ice, c) + CC(qflx_snwcp_liq, c) + CC(qflx_snwcp_discarded_ice, c)
                     + CC(qflx_snwcp_discarded_liq, c) + CC(qflx_sl_top_soil, c);
        if (lt == CTSM_ISTDLAK) {
          src = CC(qflx_snow_grnd, c) + CC(frac_sno_eff, c) * (CC(qflx_liq_grnd, c) + CC(qflx_soliddew_to_top_layer, c)
                                                                + CC(qflx_liqdew_to_top_layer, c));
          snk = CC(frac_sno_eff, c) * (CC(qflx_solidevap_from_top_layer, c) + CC(qflx_liqevap_from_top_layer, c))
                + CC(qflx_snwcp_ice, c) + CC(qflx_snwcp_liq, c) + CC(qflx_snwcp_discarded_ice, c)
                + CC(qflx_snwcp_discarded_liq, c) + CC(qflx_snow_drain, c) + CC(qflx_sl_top_soil, c);
        }
        if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP || lt == CTSM_ISTWET || lt == CTSM_ISTICE) {
          src = (CC(qflx_snow_grnd, c) - CC(qflx_snow_h2osfc, c))
                + CC(frac_sno_eff, c) * (CC(qflx_liq_grnd, c) + CC(qflx_soliddew_to_top_layer, c) + CC(qflx_liqdew_to_top_layer, c))
                + CC(qflx_h2osfc_to_ice, c);
          snk = CC(frac_sno_eff, c) * (CC(qflx_solidevap_from_top_layer, c) + CC(qflx_liqevap_from_top_layer, c))
                + CC(qflx_snwcp_ice, c) + CC(qflx_snwcp_liq, c) + CC(qflx_snwcp_discarded_ice, c)
                + CC(qflx_snwcp_discarded_liq, c) + CC(qflx_snow_drain, c) + CC(qflx_sl_top_soil, c);
        }
        CC(snow_sources, c) = src;
        CC(snow_sinks, c) = snk;
        CC(errh2osno, c) = (h2osno_total[c - bounds->begc] - CC(h2osno_old, c)) - (src - snk) * dtime;
      } else {
        CC(snow_sources, c) = 0.0;
        CC(snow_sinks, c) = 0.0;
        CC(errh2osno, c) = 0.0;
      }
    } else {
      CC(errh2osno, c) = 0.0;
    }
  }
  maxabs(f->errh2osno, begc0, bounds->begc, bounds->endc, NULL, &rep->max_abs[CTSM_BAL_H2OSNO], &rep->index[CTSM_BAL_H2OSNO]);
  if (rep->max_abs[CTSM_BAL_H2OSNO] > h2o_warning_thresh) {          /* :808-847 */
    rep->warn[CTSM_BAL_H2OSNO] = 1;
    if (rep->max_abs[CTSM_BAL_H2OSNO] > error_thresh && DAnstep > skip_steps && abort_kind < 0) {
      abort_kind = CTSM_BAL_H2OSNO; abort_index = rep->index[CTSM_BAL_H2OSNO]; abort_level = CTSM_SUBGRID_COLUMN;
    }
  }
  /* EnergyBalanceCheck :962-1005 (non-urban) */
  for (int p = bounds->begp; p <= bounds->endp; ++p) {
    if (PP(patch_active, p)) {
      const int c = PP(column, p), g = PP(gridcell, p);
      PP(errsol, p) = PP(fsa, p) + PP(fsr, p)
          - (f->forc_solad[c - begc0] + f->forc_solad[ldc + (c - begc0)] + f->forc_solai[g - begg0] + f->forc_solai[ldg + (g - begg0)]);
      PP(errlon, p) = PP(eflx_lwrad_out, p) - PP(eflx_lwrad_net, p) - CC(forc_lwrad, c);
      PP(errseb, p) = PP(sabv, p) + PP(sabg_chk, p) + CC(forc_lwrad, c) - PP(eflx_lwrad_out, p) - PP(eflx_sh_tot, p)
                      - PP(eflx_lh_tot, p) - PP(eflx_soil_grnd, p) - PP(dhsdt_canopy, p);
      PP(netrad, p) = PP(fsa, p) - PP(eflx_lwrad_net, p);
    } else {
      PP(errsol, p) = 0.0; PP(errlon, p) = 0.0; PP(errseb, p) = 0.0;
    }
  }
  const int kinds[3] = {CTSM_BAL_SOL, CTSM_BAL_LON, CTSM_BAL_SEB};
  const double* earr[3] = {f->errsol, f->errlon, f->errseb};
  for (int k = 0; k < 3; ++k) {                                     /* :1009-1097 */
    const int kd = kinds[k];
    int* mask = NULL;
    if (kd != CTSM_BAL_SEB) {
      for (int p = bounds->begp; p <= bounds->endp; ++p) okmask[p - bounds->begp] = (earr[k][p - begp0] != spval);
      mask = okmask;
    }
    /* maxabs takes base-relative mask: shift by using a temporary view */
    double mx = 0.0; int loc = 0, any = 0;
    for (int p = bounds->begp; p <= bounds->endp; ++p) {
      if (mask && !mask[p - bounds->begp]) continue;
      const double v = fabs(earr[k][p - begp0]);
      if (!any || v > mx) { mx = v; loc = p; any = 1; }
    }
    rep->max_abs[kd] = mx; rep->index[kd] = loc;
    if (mx > energy_warning_thresh && DAnstep > skip_steps) {
      rep->warn[kd] = 1;
      if (mx > error_thresh && abort_kind < 0) { abort_kind = kd; abort_index = loc; abort_level = CTSM_SUBGRID_PATCH; }
    }
  }
  {                                                                 /* :1101-1114 */
    double mx = 0.0; int loc = 0, any = 0;
    for (int c = bounds->begc; c <= bounds->endc; ++c) {
      if (!CC(col_active, c)) continue;
      const double v = fabs(CC(errsoi_col, c));
      if (!any || v > mx) { mx = v; loc = c; any = 1; }
    }
    rep->max_abs[CTSM_BAL_SOI] = mx; rep->index[CTSM_BAL_SOI] = loc;
    if (mx > 1.0e-5) {
      rep->warn[CTSM_BAL_SOI] = 1;
      if (mx > 1.e-4 && DAnstep > skip_steps && abort_kind < 0) { abort_kind = CTSM_BAL_SOI; abort_index = loc; abort_level = CTSM_SUBGRID_COLUMN; }
    }
  }
  rep->abort_kind = abort_kind;
  free(h2osno_total); free(gdyn); free(gliq); free(gice); free(okmask);
  if (abort_kind >= 0) {
    if (st) { st->code = CTSM_ERR_BALANCE; st->subgrid_index = abort_index; st->subgrid_level = abort_level; st->info = abort_kind;
              st->value = rep->max_abs[abort_kind]; }
    return CTSM_ERR_BALANCE;
  }
  return 0;
#undef CC
#undef GG
#undef PP
}

/* SoilWaterPlantSinkMod.F90:236-328 */
int oracle_vert_tran_sink_hydstress(const ctsm_bounds_t* bounds, int num_filterc, const int32_t* filterc,
                                    const ctsm_plantsink_fields_t* f) {
  (void)bounds;
  const int begc0 = f->alloc.begc, begp0 = f->alloc.begp;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1), ldp = (size_t)(f->alloc.endp - f->alloc.begp + 1);
  for (int fc = 0; fc < num_filterc; ++fc) {
    const int c = filterc[fc];
    f->qflx_phs_neg[c - begc0] = 0.0;
    for (int j = 1; j <= NLEVSOI; ++j) {
      const double grav2 = f->z[(size_t)(j - SNOSOI_LO) * ldc + (c - begc0)] * 1000.0;
      double temp = 0.0;
      for (int p = f->patchi[c - begc0]; p <= f->patchi[c - begc0] + f->npatches[c - begc0] - 1; ++p) {
        if (j == 1) f->qflx_hydr_redist[p - begp0] = 0.0;
        if (f->patch_active[p - begp0] && f->frac_veg_nosno[p - begp0] > 0) {
          if (f->wtcol[p - begp0] > 0.0) {
            const double patchflux = f->k_soil_root[(size_t)(j - 1) * ldp + (p - begp0)] *
                                     (f->smp_l[(size_t)(j - 1) * ldc + (c - begc0)] - f->vegwp[(size_t)3 * ldp + (p - begp0)] - grav2);
            if (patchflux < 0) f->qflx_hydr_redist[p - begp0] = f->qflx_hydr_redist[p - begp0] + patchflux;
            temp = temp + patchflux * f->wtcol[p - begp0];
          }
        }
      }
      f->qflx_rootsoi[(size_t)(j - 1) * ldc + (c - begc0)] = temp;
      if (temp < 0.0) f->qflx_phs_neg[c - begc0] = f->qflx_phs_neg[c - begc0] + temp;
    }
  }
  return 0;
}

/* SoilWaterPlantSinkMod.F90:332-424 (Compute_EffecRootFrac_And_VertTranSink_Default), loops as in the Fortran */
int oracle_vert_tran_sink_default(const ctsm_bounds_t* bounds, int num_filterc, const int32_t* filterc,
                                  const ctsm_plantsinkdefault_fields_t* f) {
  const int begc0 = f->alloc.begc, begp0 = f->alloc.begp;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1), ldp = (size_t)(f->alloc.endp - f->alloc.begp + 1);
  const int nb = bounds->endc - bounds->begc + 1;
  double* temp = (double*)calloc((size_t)(nb > 0 ? nb : 1), sizeof(double));      /* temp(bounds%begc:bounds%endc) = 0 */
  for (int j = 1; j <= NLEVSOI; ++j)
    for (int fc = 0; fc < num_filterc; ++fc) f->rootr_col[(size_t)(j - 1) * ldc + (filterc[fc] - begc0)] = 0.0;
  for (int j = 1; j <= NLEVSOI; ++j)
    for (int fc = 0; fc < num_filterc; ++fc) {
      const int c = filterc[fc];
      for (int p = f->patchi[c - begc0]; p <= f->patchi[c - begc0] + f->npatches[c - begc0] - 1; ++p)
        if (f->patch_active[p - begp0])
          f->rootr_col[(size_t)(j - 1) * ldc + (c - begc0)] = f->rootr_col[(size_t)(j - 1) * ldc + (c - begc0)] +
              f->rootr[(size_t)(j - 1) * ldp + (p - begp0)] * f->qflx_tran_veg[p - begp0] * f->wtcol[p - begp0];
    }
  for (int fc = 0; fc < num_filterc; ++fc) {
    const int c = filterc[fc];
    for (int p = f->patchi[c - begc0]; p <= f->patchi[c - begc0] + f->npatches[c - begc0] - 1; ++p)
      if (f->patch_active[p - begp0]) temp[c - bounds->begc] = temp[c - bounds->begc] + f->qflx_tran_veg[p - begp0] * f->wtcol[p - begp0];
  }
  for (int j = 1; j <= NLEVSOI; ++j)
    for (int fc = 0; fc < num_filterc; ++fc) {
      const int c = filterc[fc];
      double* rc = &f->rootr_col[(size_t)(j - 1) * ldc + (c - begc0)];
      if (temp[c - bounds->begc] != 0.0) *rc = *rc / temp[c - bounds->begc];
      f->qflx_rootsoi[(size_t)(j - 1) * ldc + (c - begc0)] = *rc * f->qflx_tran_veg_col[c - begc0];
    }
  free(temp);
  return 0;
}

/* ComputeLiqIceMassNonLake (TotalWaterAndHeatMod.F90:200-326) + AccumulateSoilLiqIceMassNonLake (:329-393, level-outer /
 * column-inner, i.e. ascending levels per column) for one non-urban column; canopy water by p2c (subgridAveMod.F90:312-318). */
#define WB_FIELDS(F) \
  const int32_t *snl_ = (F)->snl, *hyd_ = (F)->col_hydrologically_active, *patchi_ = (F)->patchi, *patchf_ = (F)->patchf, \
                *pact_ = (F)->patch_active; \
  const double *liq_ = (F)->h2osoi_liq, *ice_ = (F)->h2osoi_ice, *exice_ = (F)->excess_ice, *h2osfc_ = (F)->h2osfc, \
               *noly_ = (F)->h2osno_no_layers, *wa_ = (F)->wa, *tps_ = (F)->total_plant_stored_h2o, *wtcol_ = (F)->wtcol, \
               *liqcan_ = (F)->liqcan, *snocan_ = (F)->snocan
static void water_mass_nonlake(int cc, int begp0, size_t ldc, const int32_t* snl_, const int32_t* hyd_, const int32_t* patchi_,
                               const int32_t* patchf_, const int32_t* pact_, const double* liq_, const double* ice_,
                               const double* exice_, const double* h2osfc_, const double* noly_, const double* wa_,
                               const double* tps_, const double* wtcol_, const double* liqcan_, const double* snocan_,
                               double aquifer_water_baseline, double* liquid_out, double* ice_out) {
  double liqcan_col = 0.0, snocan_col = 0.0;
  for (int p = patchi_[cc]; p <= patchf_[cc]; ++p)
    if (pact_[p - begp0]) liqcan_col = liqcan_col + liqcan_[p - begp0] * wtcol_[p - begp0];
  for (int p = patchi_[cc]; p <= patchf_[cc]; ++p)
    if (pact_[p - begp0]) snocan_col = snocan_col + snocan_[p - begp0] * wtcol_[p - begp0];
  double liquid_mass = 0.0, ice_mass = 0.0;
  liquid_mass = liquid_mass + liqcan_col + tps_[cc];
  ice_mass = ice_mass + snocan_col;
  ice_mass = ice_mass + noly_[cc];
  const int snl = snl_[cc];
  for (int j = snl + 1; j <= 0; ++j) {
    liquid_mass = liquid_mass + liq_[(size_t)(j - SNOSOI_LO) * ldc + cc];
    ice_mass = ice_mass + ice_[(size_t)(j - SNOSOI_LO) * ldc + cc];
  }
  if (hyd_[cc]) liquid_mass = liquid_mass + (wa_[cc] - aquifer_water_baseline);
  liquid_mass = liquid_mass + h2osfc_[cc];
  for (int j = 1; j <= CTSM_NLEVGRND; ++j) {
    liquid_mass = liquid_mass + liq_[(size_t)(j - SNOSOI_LO) * ldc + cc];
    ice_mass = ice_mass + ice_[(size_t)(j - SNOSOI_LO) * ldc + cc] + exice_[(size_t)(j - 1) * ldc + cc];
  }
  *liquid_out = liquid_mass; *ice_out = ice_mass;
}
/* snow and soil layers of a lake column: ComputeLiqIceMassLake :449-462 */
static void water_mass_lake_layers(int cc, size_t ldc, const int32_t* snl_, const double* liq_, const double* ice_,
                                   const double* noly_, double* liquid_mass, double* ice_mass) {
  *ice_mass = *ice_mass + noly_[cc];
  for (int j = snl_[cc] + 1; j <= 0; ++j) {
    *liquid_mass = *liquid_mass + liq_[(size_t)(j - SNOSOI_LO) * ldc + cc];
    *ice_mass = *ice_mass + ice_[(size_t)(j - SNOSOI_LO) * ldc + cc];
  }
  for (int j = 1; j <= CTSM_NLEVGRND; ++j) {
    *liquid_mass = *liquid_mass + liq_[(size_t)(j - SNOSOI_LO) * ldc + cc];
    *ice_mass = *ice_mass + ice_[(size_t)(j - SNOSOI_LO) * ldc + cc];
  }
}
static double total_h2osno(int cc, size_t ldc, const int32_t* snl_, const double* liq_, const double* ice_, const double* noly_) {
  double t = noly_[cc];                                                  /* CalculateTotalH2osno, WaterStateType.F90:887-896 */
  for (int j = snl_[cc] + 1; j <= 0; ++j) t = t + ice_[(size_t)(j - SNOSOI_LO) * ldc + cc] + liq_[(size_t)(j - SNOSOI_LO) * ldc + cc];
  return t;
}

/* BeginWaterColumnBalanceSingle (BalanceCheckMod.F90:353-442, use_aquifer_layer = .false.): ComputeWaterMassNonLake
 * (subtract_dynbal_baselines = .false.), ComputeWaterMassLake (add_lake_water_and_subtract_dynbal_baselines = .false.),
 * CalculateTotalH2osno over both filters.  Non-urban columns. */
int oracle_begin_water_column_balance(const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                                      int num_lakec, const int32_t* filter_lakec,
                                      const ctsm_waterbalance_fields_t* f, double aquifer_water_baseline, ctsm_status_t* st) {
  (void)bounds;
  const int begc0 = f->alloc.begc, begp0 = f->alloc.begp;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1);
  if (st) memset(st, 0, sizeof *st);
  WB_FIELDS(f);
  for (int fc = 0; fc < num_nolakec; ++fc) {
    const int c = filter_nolakec[fc], cc = c - begc0;
    const int lt = f->lun_itype[cc];
    if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) {
      if (st) { st->code = CTSM_ERR_URBAN; st->subgrid_level = CTSM_SUBGRID_COLUMN; st->subgrid_index = c; }
      return CTSM_ERR_URBAN;
    }
    double liquid_mass, ice_mass;
    water_mass_nonlake(cc, begp0, ldc, snl_, hyd_, patchi_, patchf_, pact_, liq_, ice_, exice_, h2osfc_, noly_, wa_, tps_, wtcol_,
                       liqcan_, snocan_, aquifer_water_baseline, &liquid_mass, &ice_mass);
    f->begwb[cc] = liquid_mass + ice_mass;
    f->h2osno_old[cc] = total_h2osno(cc, ldc, snl_, liq_, ice_, noly_);
  }
  for (int fc = 0; fc < num_lakec; ++fc) {
    const int cc = filter_lakec[fc] - begc0;
    double liquid_mass = 0.0, ice_mass = 0.0;
    water_mass_lake_layers(cc, ldc, snl_, liq_, ice_, noly_, &liquid_mass, &ice_mass);
    f->begwb[cc] = liquid_mass + ice_mass;
    f->h2osno_old[cc] = total_h2osno(cc, ldc, snl_, liq_, ice_, noly_);
  }
  return 0;
}

/* WaterGridcellBalanceSingle (BalanceCheckMod.F90:212-350): bulk water, use_aquifer_layer = .false., no hillslope routing */
int oracle_water_gridcell_balance(const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec, int num_lakec,
                                  const int32_t* filter_lakec, const ctsm_watergridbalance_fields_t* f,
                                  double aquifer_water_baseline, int flag_endwb, ctsm_status_t* st) {
  const int begc0 = f->alloc.begc, begp0 = f->alloc.begp, begg0 = f->alloc.begg;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1);
  if (st) memset(st, 0, sizeof *st);
  WB_FIELDS(f);
  const int ncb = bounds->endc - bounds->begc + 1;
  double* wb_col = (double*)malloc(sizeof(double) * (size_t)(ncb > 0 ? ncb : 1));
  for (int i = 0; i < ncb; ++i) wb_col[i] = 0.0;         /* the reference leaves columns outside both filters undefined */
  for (int fc = 0; fc < num_nolakec; ++fc) {
    const int c = filter_nolakec[fc], cc = c - begc0;
    const int lt = f->lun_itype[cc];
    if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) {
      if (st) { st->code = CTSM_ERR_URBAN; st->subgrid_level = CTSM_SUBGRID_COLUMN; st->subgrid_index = c; }
      free(wb_col);
      return CTSM_ERR_URBAN;
    }
    double liquid_mass, ice_mass;
    water_mass_nonlake(cc, begp0, ldc, snl_, hyd_, patchi_, patchf_, pact_, liq_, ice_, exice_, h2osfc_, noly_, wa_, tps_, wtcol_,
                       liqcan_, snocan_, aquifer_water_baseline, &liquid_mass, &ice_mass);
    liquid_mass = liquid_mass - f->dynbal_baseline_liq[cc];             /* subtract_dynbal_baselines :316-322 */
    ice_mass = ice_mass - f->dynbal_baseline_ice[cc];
    wb_col[c - bounds->begc] = liquid_mass + ice_mass;
  }
  for (int fc = 0; fc < num_lakec; ++fc) {
    const int c = filter_lakec[fc], cc = c - begc0;
    double liquid_mass = 0.0, ice_mass = 0.0;
    liquid_mass = liquid_mass - f->dynbal_baseline_liq[cc];             /* ComputeLiqIceMassLake :434-447 */
    ice_mass = ice_mass - f->dynbal_baseline_ice[cc];
    for (int j = 1; j <= CTSM_NLEVLAK; ++j) {                           /* AccumulateLiqIceMassLake :536-544, tracer_ratio = 1 */
      const double dzl = f->dz_lake[(size_t)(j - 1) * ldc + cc], fr = f->lake_icefrac[(size_t)(j - 1) * ldc + cc];
      const double h2olak_liq = dzl * 1.000e3 * (1 - fr) * 1.0;
      const double h2olak_ice = dzl * 1.000e3 * fr * 1.0;
      liquid_mass = liquid_mass + h2olak_liq;
      ice_mass = ice_mass + h2olak_ice;
    }
    water_mass_lake_layers(cc, ldc, snl_, liq_, ice_, noly_, &liquid_mass, &ice_mass);
    wb_col[c - bounds->begc] = liquid_mass + ice_mass;
  }
  /* c2g, subgridAveMod.F90:791-816 (scale factors 1 off the urban landunits) */
  int rc = 0;
  for (int g = bounds->begg; g <= bounds->endg; ++g) {
    const int gg = g - begg0;
    double garr = 1.0e36, sumwt = 0.0;
    for (int c = f->grc_coli[gg]; c <= f->grc_colf[gg]; ++c) {
      const int cc = c - begc0;
      if (c < bounds->begc || c > bounds->endc) continue;
      if (f->col_active[cc] && f->wtgcell[cc] != 0.0) {
        const double v = wb_col[c - bounds->begc];
        if (v != 1.0e36) {
          if (sumwt == 0.0) garr = 0.0;
          garr = garr + v * 1.0 * 1.0 * f->wtgcell[cc];
          sumwt = sumwt + f->wtgcell[cc];
        }
      }
    }
    if (sumwt > 1.0 + 1.e-6) {
      rc = CTSM_ERR_BALANCE;
      if (st) { st->code = rc; st->subgrid_level = CTSM_SUBGRID_GRIDCELL; st->subgrid_index = g; }   /* the reference reports the last such g */
    } else if (sumwt != 0.0) {
      garr = garr / sumwt;
    }
    double wb = garr - f->qflx_liq_dynbal_left_to_dribble[gg] - f->qflx_ice_dynbal_left_to_dribble[gg];
    if (flag_endwb) f->endwb_grc[gg] = wb - 0.0;                         /* wa_reset_nonconservation_gain_grc = 0 */
    else f->begwb_grc[gg] = wb;
  }
  free(wb_col);
  return rc;
}
