/* oracle_soilfluxes.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of
 *   SoilFluxes      src/biogeophys/SoilFluxesMod.F90:37-521
 *   p2c_1d_filter   src/main/subgridAveMod.F90:292-320
 * for non-urban landunits (lun%urbpoi = .false.: the urban branches of :196-200, :254-272, :310-322, :370-383,
 * :417-419, :485-490, :502-509 are not restated; an urban column in the filter is an error here as it is in the
 * product).  Loop for loop as the reference: filter loops over clump-sized scratch arrays (tinc, t_grnd0).
 * PARITY UNPINNED by the reference's own tests (no unit test covers SoilFluxes).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define NLEVSNO CTSM_NLEVSNO
#define NLEVGRND CTSM_NLEVGRND
#define SNOSOI_LO (-NLEVSNO + 1)
static const double hvap = 2.501e6, tfrz = 273.15, sb = 5.67e-8;

static double pow4(double t) { const double t2 = t * t; return t2 * t2; }   /* x**4 as gfortran expands it */
static double pow3(double t) { return (t * t) * t; }

int oracle_soilfluxes(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                      int num_nolakep, const int32_t* filter_nolakep, const ctsm_soilfluxes_fields_t* f, ctsm_status_t* st) {
  const int begc0 = f->alloc.begc, begp0 = f->alloc.begp;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1);
  const double dtime = prm->dtime;
  if (st) memset(st, 0, sizeof *st);
#define C1(name, c) f->name[(c) - begc0]
#define C2(name, c, j) f->name[(size_t)((j) - SNOSOI_LO) * ldc + (size_t)((c) - begc0)]
#define P1(name, p) f->name[(p) - begp0]
  const int nc = bounds->endc - bounds->begc + 1;
  double* tinc = (double*)calloc((size_t)(nc > 0 ? nc : 1), sizeof(double));
  double* t_grnd0 = (double*)calloc((size_t)(nc > 0 ? nc : 1), sizeof(double));
#define TINC(c) tinc[(c) - bounds->begc]
#define TG0(c) t_grnd0[(c) - bounds->begc]

  for (int fc = 0; fc < num_nolakec; ++fc) {                  /* :167-184 */
    const int c = filter_nolakec[fc];
    const int lt = C1(lun_itype, c);
    if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) {
      if (st) { st->code = CTSM_ERR_URBAN; st->subgrid_level = CTSM_SUBGRID_COLUMN; st->subgrid_index = c; }
      free(tinc); free(t_grnd0);
      return CTSM_ERR_URBAN;
    }
    const int snl = C1(snl, c);
    if (snl < 0) {
      TG0(c) = C1(frac_sno_eff, c) * C2(t_ssbef, c, snl + 1) + (1 - C1(frac_sno_eff, c) - C1(frac_h2osfc, c)) * C2(t_ssbef, c, 1)
               + C1(frac_h2osfc, c) * C1(t_h2osfc_bef, c);
    } else {
      TG0(c) = (1 - C1(frac_h2osfc, c)) * C2(t_ssbef, c, 1) + C1(frac_h2osfc, c) * C1(t_h2osfc_bef, c);
    }
    TINC(c) = C1(t_grnd, c) - TG0(c);
  }

  for (int fp = 0; fp < num_nolakep; ++fp) {                  /* :188-206 */
    const int p = filter_nolakep[fp];
    const int c = P1(column, p);
    P1(eflx_sh_grnd, p) = P1(eflx_sh_grnd, p) + TINC(c) * P1(cgrnds, p);
    P1(qflx_evap_soi, p) = P1(qflx_evap_soi, p) + TINC(c) * P1(cgrndl, p);
    P1(qflx_ev_snow, p) = P1(qflx_ev_snow, p) + TINC(c) * P1(cgrndl, p);
    P1(qflx_ev_soil, p) = P1(qflx_ev_soil, p) + TINC(c) * P1(cgrndl, p);
    P1(qflx_ev_h2osfc, p) = P1(qflx_ev_h2osfc, p) + TINC(c) * P1(cgrndl, p);
  }

  for (int fp = 0; fp < num_nolakep; ++fp) {                  /* :209-274 partition evaporation */
    const int p = filter_nolakep[fp];
    const int c = P1(column, p);
    const int j = C1(snl, c) + 1;
    P1(qflx_liqevap_from_top_layer_patch, p) = 0.0;
    P1(qflx_solidevap_from_top_layer_patch, p) = 0.0;
    P1(qflx_soliddew_to_top_layer_patch, p) = 0.0;
    P1(qflx_liqdew_to_top_layer_patch, p) = 0.0;
    if (P1(qflx_ev_snow, p) >= 0.0) {
      if ((C2(h2osoi_liq, c, j) + C2(h2osoi_ice, c, j)) > 0.0) {
        P1(qflx_liqevap_from_top_layer_patch, p) =
            fmax(P1(qflx_ev_snow, p) * (C2(h2osoi_liq, c, j) / (C2(h2osoi_liq, c, j) + C2(h2osoi_ice, c, j))), 0.0);
      } else {
        P1(qflx_liqevap_from_top_layer_patch, p) = 0.0;
      }
      P1(qflx_solidevap_from_top_layer_patch, p) = P1(qflx_ev_snow, p) - P1(qflx_liqevap_from_top_layer_patch, p);
    } else {
      if (C1(t_grnd, c) < tfrz) P1(qflx_soliddew_to_top_layer_patch, p) = fabs(P1(qflx_ev_snow, p));
      else P1(qflx_liqdew_to_top_layer_patch, p) = fabs(P1(qflx_ev_snow, p));
    }
  }

  for (int fp = 0; fp < num_nolakep; ++fp) {                  /* :277-331 limit evaporation to the available moisture */
    const int p = filter_nolakep[fp];
    const int c = P1(column, p);
    const int j = C1(snl, c) + 1;
    if (j < 1) {
      const double evaporation_limit = (C2(h2osoi_ice, c, j) + C2(h2osoi_liq, c, j)) / (C1(frac_sno_eff, c) * dtime);
      if (P1(qflx_ev_snow, p) > evaporation_limit) {
        const double evaporation_demand = P1(qflx_ev_snow, p);
        P1(qflx_ev_snow, p) = evaporation_limit;
        P1(qflx_evap_soi, p) = P1(qflx_evap_soi, p) - C1(frac_sno_eff, c) * (evaporation_demand - evaporation_limit);
        P1(qflx_liqevap_from_top_layer_patch, p) = fmax(C2(h2osoi_liq, c, j) / (C1(frac_sno_eff, c) * dtime), 0.0);
        P1(qflx_solidevap_from_top_layer_patch, p) = fmax(C2(h2osoi_ice, c, j) / (C1(frac_sno_eff, c) * dtime), 0.0);
        P1(eflx_sh_grnd, p) = P1(eflx_sh_grnd, p) + C1(frac_sno_eff, c) * (evaporation_demand - evaporation_limit) * C1(htvp, c);
      }
    }
    if (j == 1 && C1(frac_h2osfc, c) < 1.0) {
      const double evaporation_limit = C2(h2osoi_ice, c, j) / (dtime * (1.0 - C1(frac_h2osfc, c)));
      if (P1(qflx_solidevap_from_top_layer_patch, p) >= evaporation_limit) {
        const double evaporation_demand = P1(qflx_solidevap_from_top_layer_patch, p);
        P1(qflx_solidevap_from_top_layer_patch, p) = evaporation_limit;
        P1(qflx_liqevap_from_top_layer_patch, p) = P1(qflx_liqevap_from_top_layer_patch, p) + (evaporation_demand - evaporation_limit);
      }
    }
  }

  for (int fp = 0; fp < num_nolakep; ++fp) {                  /* :338-400 ground heat flux, totals */
    const int p = filter_nolakep[fp];
    const int c = P1(column, p);
    const int lt = C1(lun_itype, c);
    const int snl = C1(snl, c);
    const double lw_grnd = (C1(frac_sno_eff, c) * pow4(C2(t_ssbef, c, snl + 1))
                            + (1.0 - C1(frac_sno_eff, c) - C1(frac_h2osfc, c)) * pow4(C2(t_ssbef, c, 1))
                            + C1(frac_h2osfc, c) * pow4(C1(t_h2osfc_bef, c)));
    P1(eflx_soil_grnd, p) = ((1.0 - C1(frac_sno_eff, c)) * P1(sabg_soil, p) + C1(frac_sno_eff, c) * P1(sabg_snow, p)) + P1(dlrad, p)
                            + (1 - P1(frac_veg_nosno, p)) * C1(emg, c) * C1(forc_lwrad, c)
                            - C1(emg, c) * sb * lw_grnd - C1(emg, c) * sb * pow3(TG0(c)) * (4.0 * TINC(c))
                            - (P1(eflx_sh_grnd, p) + P1(qflx_evap_soi, p) * C1(htvp, c));
    if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP) P1(eflx_soil_grnd_r, p) = P1(eflx_soil_grnd, p);
    P1(eflx_sh_tot, p) = P1(eflx_sh_veg, p) + P1(eflx_sh_grnd, p);
    P1(eflx_sh_tot, p) = P1(eflx_sh_tot, p) + P1(eflx_sh_stem, p);
    P1(qflx_evap_tot_patch, p) = P1(qflx_evap_veg, p) + P1(qflx_evap_soi, p);
    P1(eflx_lh_tot, p) = hvap * P1(qflx_evap_veg, p) + C1(htvp, c) * P1(qflx_evap_soi, p);
    if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP) {
      P1(eflx_lh_tot_r, p) = P1(eflx_lh_tot, p);
      P1(eflx_sh_tot_r, p) = P1(eflx_sh_tot, p);
    }
    P1(qflx_evap_can, p) = P1(qflx_evap_veg, p) - P1(qflx_tran_veg, p);
    P1(eflx_lh_vege, p) = (P1(qflx_evap_veg, p) - P1(qflx_tran_veg, p)) * hvap;
    P1(eflx_lh_vegt, p) = P1(qflx_tran_veg, p) * hvap;
    P1(eflx_lh_grnd, p) = P1(qflx_evap_soi, p) * C1(htvp, c);
  }

  for (int fp = 0; fp < num_nolakep; ++fp) {                  /* :406-419 soil energy balance check */
    const int p = filter_nolakep[fp];
    const int c = P1(column, p);
    P1(errsoi_patch, p) = P1(eflx_soil_grnd, p) - C1(xmf, c) - C1(xmf_h2osfc, c)
                          - C1(frac_h2osfc, c) * (C1(t_h2osfc, c) - C1(t_h2osfc_bef, c)) * (C1(c_h2osfc, c) / dtime);
    P1(errsoi_patch, p) = P1(errsoi_patch, p) + C1(eflx_h2osfc_to_snow, c);
  }
  for (int j = -NLEVSNO + 1; j <= NLEVGRND; ++j) {            /* :420-436 level-outer, filter-inner */
    for (int fp = 0; fp < num_nolakep; ++fp) {
      const int p = filter_nolakep[fp];
      const int c = P1(column, p);
      if (j >= C1(snl, c) + 1 && j < 1)
        P1(errsoi_patch, p) = P1(errsoi_patch, p) - C1(frac_sno_eff, c) * (C2(t_soisno, c, j) - C2(t_ssbef, c, j)) / C2(fact, c, j);
      if (j >= 1) P1(errsoi_patch, p) = P1(errsoi_patch, p) - (C2(t_soisno, c, j) - C2(t_ssbef, c, j)) / C2(fact, c, j);
    }
  }

  for (int fp = 0; fp < num_nolakep; ++fp) {                  /* :463-500 outgoing longwave, skin temperature */
    const int p = filter_nolakep[fp];
    const int c = P1(column, p);
    const int lt = C1(lun_itype, c);
    const int snl = C1(snl, c);
    const double lw_grnd = (C1(frac_sno_eff, c) * pow4(C2(t_ssbef, c, snl + 1))
                            + (1.0 - C1(frac_sno_eff, c) - C1(frac_h2osfc, c)) * pow4(C2(t_ssbef, c, 1))
                            + C1(frac_h2osfc, c) * pow4(C1(t_h2osfc_bef, c)));
    P1(eflx_lwrad_out, p) = P1(ulrad, p) + (1 - P1(frac_veg_nosno, p)) * (1. - C1(emg, c)) * C1(forc_lwrad, c)
                            + (1 - P1(frac_veg_nosno, p)) * C1(emg, c) * sb * lw_grnd
                            + 4.0 * C1(emg, c) * sb * pow3(TG0(c)) * TINC(c);
    if (P1(frac_veg_nosno, p) == 0) P1(t_skin, p) = sqrt(sqrt(lw_grnd));
    P1(eflx_lwrad_net, p) = P1(eflx_lwrad_out, p) - C1(forc_lwrad, c);
    if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP) {
      P1(eflx_lwrad_net_r, p) = P1(eflx_lwrad_out, p) - C1(forc_lwrad, c);
      P1(eflx_lwrad_out_r, p) = P1(eflx_lwrad_out, p);
    }
  }

  for (int fc = 0; fc < num_nolakec; ++fc) {                  /* p2c :312-318 */
    const int c = filter_nolakec[fc];
    C1(errsoi_col, c) = 0.0;
    for (int p = C1(patchi, c); p <= C1(patchf, c); ++p)
      if (P1(patch_active, p)) C1(errsoi_col, c) = C1(errsoi_col, c) + P1(errsoi_patch, p) * P1(wtcol, p);
  }
  free(tinc); free(t_grnd0);
  return 0;
#undef C1
#undef C2
#undef P1
#undef TINC
#undef TG0
}

/* clm_drv_patch2col: src/main/clm_driver.F90:1655-1739 with p2c_1d_filter (subgridAveMod.F90:312-318) */
static void p2c_filter(const ctsm_patch2col_fields_t* f, int numfc, const int32_t* filterc, const double* parr, double* carr) {
  const int begc0 = f->alloc.begc, begp0 = f->alloc.begp;
  for (int fc = 0; fc < numfc; ++fc) {
    const int c = filterc[fc];
    carr[c - begc0] = 0.0;
    for (int p = f->patchi[c - begc0]; p <= f->patchf[c - begc0]; ++p)
      if (f->patch_active[p - begp0]) carr[c - begc0] = carr[c - begc0] + parr[p - begp0] * f->wtcol[p - begp0];
  }
}

int oracle_patch2col(const ctsm_bounds_t* bounds, int num_allc, const int32_t* filter_allc, int num_nolakec,
                     const int32_t* filter_nolakec, const ctsm_patch2col_fields_t* f) {
  (void)bounds;
  p2c_filter(f, num_nolakec, filter_nolakec, f->qflx_ev_snow, f->qflx_ev_snow_col);
  p2c_filter(f, num_nolakec, filter_nolakec, f->qflx_ev_soil, f->qflx_ev_soil_col);
  p2c_filter(f, num_nolakec, filter_nolakec, f->qflx_ev_h2osfc, f->qflx_ev_h2osfc_col);
  p2c_filter(f, num_nolakec, filter_nolakec, f->qflx_evap_soi, f->qflx_evap_soi_col);
  p2c_filter(f, num_nolakec, filter_nolakec, f->qflx_evap_tot_patch, f->qflx_evap_tot);
  p2c_filter(f, num_nolakec, filter_nolakec, f->qflx_tran_veg, f->qflx_tran_veg_col);
  p2c_filter(f, num_nolakec, filter_nolakec, f->qflx_liqevap_from_top_layer_patch, f->qflx_liqevap_from_top_layer);
  p2c_filter(f, num_allc, filter_allc, f->qflx_evap_soi, f->qflx_evap_soi_col);
  p2c_filter(f, num_nolakec, filter_nolakec, f->qflx_liqdew_to_top_layer_patch, f->qflx_liqdew_to_top_layer);
  p2c_filter(f, num_nolakec, filter_nolakec, f->qflx_solidevap_from_top_layer_patch, f->qflx_solidevap_from_top_layer);
  p2c_filter(f, num_nolakec, filter_nolakec, f->qflx_soliddew_to_top_layer_patch, f->qflx_soliddew_to_top_layer);
  return 0;
}

