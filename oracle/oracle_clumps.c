/* oracle_clumps.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The reference's CPU parallelism for the hot path: an OpenMP PARALLEL DO over
 * clumps (src/main/clm_driver.F90:525-527, closes :1345), each thread calling
 * the physics routines with its clump's bounds and filters.  This driver runs
 * the oracle's routines the same way and is what bench.py times as the
 * cpu_baseline ("port": C restatement, not gfortran — SURVEY.md F10).
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "../include/ctsm_b200_defaults.h"

/* the clm6_0 defaults of ctsm_params_t (shared header), so that the CPU arm never needs the CUDA library */
void oracle_default_params(ctsm_params_t* p) { ctsm_default_params_fill(p); }

int oracle_step_clumps(const ctsm_params_t* prm, int nclumps, const oracle_clump_t* clumps,
                       const ctsm_soiltemperature_fields_t* ft, const ctsm_soilwater_fields_t* fw,
                       const ctsm_canopyfluxes_fields_t* fc, int which) {
  int rc_all = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int nc = 0; nc < nclumps; ++nc) {
    const oracle_clump_t* k = &clumps[nc];
    ctsm_status_t st;
    int rc = 0;
    if ((which & 4) && fc)
      rc = oracle_canopyfluxes(prm, &k->bounds, k->num_exposedvegp, k->filter_exposedvegp, fc, &st);
    if (!rc && (which & 1) && ft)
      rc = oracle_soiltemperature(prm, &k->bounds, k->num_nolakep, k->filter_nolakep, k->num_nolakec,
                                  k->filter_nolakec, ft, &st);
    if (!rc && (which & 2) && fw)
      rc = oracle_soilwater(prm, &k->bounds, k->num_hydrologyc, k->filter_hydrologyc, fw, &st);
    if (rc) {
#pragma omp critical
      rc_all = rc;
    }
  }
  return rc_all;
}

/* The full biogeophysics step of BASELINE.json config 4 in clm_drv order (clm_driver.F90:766, :900, :950 ->
 * HydrologyNoDrainageMod.F90:339,346, :1422): CanopyFluxes -> SoilTemperature -> root-water sink -> SoilWater ->
 * BalanceCheck over all columns of the clump (SoilFluxes, :921, between SoilTemperature and the sink).  `which` bits:
 * 1 SoilTemperature, 2 SoilWater, 4 CanopyFluxes, 8 plant sink, 16 BalanceCheck, 32 SoilFluxes, 64 clm_drv_patch2col. */
static const ctsm_plantsinkdefault_fields_t* g_sink_default = NULL;
void oracle_set_plantsink_default(const ctsm_plantsinkdefault_fields_t* f) { g_sink_default = f; }

int oracle_fullstep_clumps(const ctsm_params_t* prm, int nclumps, const oracle_clump_t* clumps,
                           const ctsm_soiltemperature_fields_t* ft, const ctsm_soilwater_fields_t* fw,
                           const ctsm_canopyfluxes_fields_t* fc, const ctsm_plantsink_fields_t* fs,
                           const ctsm_balancecheck_fields_t* fb, const ctsm_soilfluxes_fields_t* fx,
                           const ctsm_patch2col_fields_t* f2c, int DAnstep, int which) {
  int rc_all = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int nc = 0; nc < nclumps; ++nc) {
    const oracle_clump_t* k = &clumps[nc];
    ctsm_status_t st;
    int rc = 0;
    if ((which & 4) && fc)
      rc = oracle_canopyfluxes(prm, &k->bounds, k->num_exposedvegp, k->filter_exposedvegp, fc, &st);
    if (!rc && (which & 1) && ft)
      rc = oracle_soiltemperature(prm, &k->bounds, k->num_nolakep, k->filter_nolakep, k->num_nolakec,
                                  k->filter_nolakec, ft, &st);
    if (!rc && (which & 32) && fx)
      rc = oracle_soilfluxes(prm, &k->bounds, k->num_nolakec, k->filter_nolakec, k->num_nolakep, k->filter_nolakep, fx, &st);
    if (!rc && (which & 64) && f2c) {                      /* clm_drv_patch2col, clm_driver.F90:936 */
      const int n = k->bounds.endc - k->bounds.begc + 1;
      int32_t* allc = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
      for (int i = 0; i < n; ++i) allc[i] = k->bounds.begc + i;
      rc = oracle_patch2col(&k->bounds, n, allc, k->num_nolakec, k->filter_nolakec, f2c);
      free(allc);
    }
    if (!rc && (which & 8) && !prm->use_hydrstress) {
      if (!g_sink_default) rc = CTSM_ERR_BAD_ARG;
      else rc = oracle_vert_tran_sink_default(&k->bounds, k->num_hydrologyc, k->filter_hydrologyc, g_sink_default);
    } else if (!rc && (which & 8) && fs)
      rc = oracle_vert_tran_sink_hydstress(&k->bounds, k->num_hydrologyc, k->filter_hydrologyc, fs);
    if (!rc && (which & 2) && fw)
      rc = oracle_soilwater(prm, &k->bounds, k->num_hydrologyc, k->filter_hydrologyc, fw, &st);
    if (!rc && (which & 16) && fb) {
      const int n = k->bounds.endc - k->bounds.begc + 1;
      int32_t* allc = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
      for (int i = 0; i < n; ++i) allc[i] = k->bounds.begc + i;
      ctsm_balance_report_t rep;
      rc = oracle_balancecheck(prm, &k->bounds, n, allc, fb, DAnstep, &rep, &st);
      free(allc);
    }
    if (rc) {
#pragma omp critical
      rc_all = rc;
    }
  }
  return rc_all;
}

