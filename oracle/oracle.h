/* oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, gcc -O2 -ffp-contract=off like the reference's
 * gnu.cmake:51-53 build) of the CTSM biogeophysics hot path.  Each function
 * follows the Fortran routine it cites loop for loop (level-outer /
 * filter-inner, full-size temporaries, one LAPACK call per column).
 *
 * Parity status: the reference cannot be compiled in this environment (no
 * Fortran compiler, un-vendored submodules), and its own unit
 * tests pin only plc/d1plc, quadratic, truncate_small_values and the
 * BalanceCheck skip steps.  Those known answers are checked in
 * tests/test_oracle_golden.py.  Everything else here is "PARITY UNPINNED": a
 * restatement checked by analytic invariants (tests/test_oracle_*.py), not by
 * reference outputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
 * reference) may call into this library.  The product (libctsm_b200.so) never
 * links or loads it.
 */
#ifndef CTSM_ORACLE_H
#define CTSM_ORACLE_H

#include "../include/ctsm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* LAPACK restatements (oracle_lapack.c) */
void oracle_dgbsv(int n, int kl, int ku, int nrhs, double* ab, int ldab, int* ipiv, double* b, int ldb, int* info);
void oracle_dgtsv(int n, int nrhs, double* dl, double* d, double* du, double* b, int ldb, int* info);

/* TridiagonalMod.F90:23-91 */
void oracle_tridiagonal(const ctsm_bounds_t* bounds, int lbj, int ubj, const int32_t* jtop, int numf,
                        const int32_t* filter, const double* a, const double* b, const double* c,
                        const double* r, double* u);
/* BandDiagonalMod.F90:29-221; returns 0 or the first failing column's dgbsv info in st */
int oracle_banddiagonal(const ctsm_bounds_t* bounds, int lbj, int ubj, const int32_t* jtop,
                        const int32_t* jbot, int numf, const int32_t* filter, int nband,
                        const double* b, const double* r, double* u, ctsm_status_t* st);
/* SoilWaterMovementMod.F90:1279-1299 call shape */
int oracle_dgtsv_batch(const ctsm_bounds_t* bounds, int nlev, const int32_t* nlayers, int numf,
                       const int32_t* filter, const double* amx, const double* bmx, const double* cmx,
                       const double* rmx, double* x, ctsm_status_t* st);

/* SoilWaterMovementMod.F90:240 -> :976 */
int oracle_soilwater(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_hydrologyc,
                     const int32_t* filter_hydrologyc, const ctsm_soilwater_fields_t* f, ctsm_status_t* st);

/* SoilTemperatureMod.F90:92 */
int oracle_soiltemperature(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_nolakep,
                           const int32_t* filter_nolakep, int num_nolakec, const int32_t* filter_nolakec,
                           const ctsm_soiltemperature_fields_t* f, ctsm_status_t* st);

/* CanopyFluxesMod.F90:191 (use_hydrstress) */
int oracle_canopyfluxes(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_exposedvegp,
                        const int32_t* filter_exposedvegp, const ctsm_canopyfluxes_fields_t* f, ctsm_status_t* st);
/* BalanceCheckMod.F90:445-857 + EnergyBalanceCheck :859-1119 (prm->balance_skip_steps from BalanceCheckInit) */
int oracle_balancecheck(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_allc, const int32_t* filter_allc,
                        const ctsm_balancecheck_fields_t* f, int DAnstep, ctsm_balance_report_t* rep, ctsm_status_t* st);
int oracle_balancecheck_skip_steps(double dtime);
/* SoilWaterPlantSinkMod.F90:236-328 */
int oracle_vert_tran_sink_hydstress(const ctsm_bounds_t* bounds, int num_filterc, const int32_t* filterc,
                                    const ctsm_plantsink_fields_t* f);
/* oracle_preflux.c: the three routines before CanopyFluxes */
int oracle_biogeophys_pre_flux_calcs(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_nolakec,
                                     const int32_t* filter_nolakec, int num_nolakep, const int32_t* filter_nolakep,
                                     int num_urbanc, int time_flags, const ctsm_preflux_fields_t* f, ctsm_status_t* st);
int oracle_calculate_surface_humidity(const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                                      const ctsm_surfacehumidity_fields_t* f, ctsm_status_t* st);
int oracle_bare_ground_fluxes(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_noexposedvegp,
                              const int32_t* filter_noexposedvegp, const ctsm_baregroundfluxes_fields_t* f, ctsm_status_t* st);
/* oracle_hydrology.c: the infiltration chain of HydrologyNoDrainage (first part) */
int oracle_hydrology_infiltration(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_nolakec,
                                  const int32_t* filter_nolakec, int num_hydrologyc, const int32_t* filter_hydrologyc,
                                  int num_urbanc, const ctsm_infiltration_fields_t* f, ctsm_status_t* st);
int oracle_water_table(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_hydrologyc, const int32_t* filter_hydrologyc,
                       int num_urbanc, const ctsm_watertable_fields_t* f, ctsm_status_t* st);
int oracle_hydrology_diagnostics(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                                 int num_snowc, const int32_t* filter_snowc, int num_nosnowc, const int32_t* filter_nosnowc,
                                 int num_hydrologyc, const int32_t* filter_hydrologyc, int num_urbanc,
                                 const ctsm_hydrodiag_fields_t* f, ctsm_status_t* st);
/* oracle_ozone.c: CalcOzoneUptake / CalcOzoneStress */
int oracle_calc_ozone_uptake(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_exposedvegp,
                             const int32_t* filter_exposedvegp, const ctsm_ozone_fields_t* f, ctsm_status_t* st);
int oracle_calc_ozone_stress(const ctsm_bounds_t* bounds, int num_exposedvegp, const int32_t* filter_exposedvegp, int num_noexposedvegp,
                             const int32_t* filter_noexposedvegp, int stress_method, int is_time_to_run_luna,
                             const ctsm_ozone_fields_t* f, ctsm_status_t* st);
/* oracle_snow.c: the snow routines of HydrologyNoDrainage */
void oracle_snow_dz_limits(const ctsm_params_t* prm, double* dzmin, double* dzmax_l, double* dzmax_u);
void oracle_build_snow_filter(int num_nolakec, const int32_t* filter_nolakec, const int32_t* snl, int begc0,
                              int32_t* filter_snowc, int32_t* num_snowc, int32_t* filter_nosnowc, int32_t* num_nosnowc);
int oracle_snow_water(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_snowc, const int32_t* filter_snowc,
                      int num_nosnowc, const int32_t* filter_nosnowc, const ctsm_snowwater_fields_t* f, ctsm_status_t* st);
int oracle_snow_capping(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_initc, const int32_t* filter_initc,
                        int num_snowc, const int32_t* filter_snowc, const ctsm_snowcapping_fields_t* f, int nstep, ctsm_status_t* st);
int oracle_snow_layers(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_snowc, const int32_t* filter_snowc,
                       const ctsm_snowlayers_fields_t* f, ctsm_status_t* st);
int oracle_vert_tran_sink_default(const ctsm_bounds_t* bounds, int num_filterc, const int32_t* filterc,
                                  const ctsm_plantsinkdefault_fields_t* f);
/* the field struct oracle_fullstep_clumps hands to the default sink when prm->use_hydrstress == 0 (its own plant-sink
 * argument is the PHS struct); NULL = none */
void oracle_set_plantsink_default(const ctsm_plantsinkdefault_fields_t* f);
void oracle_truncate_small_values(int num_f, const int32_t* filter_f, int lb, const double* data_baseline, double* data,
                                  double rel_epsilon);
/* filterMod.F90:595-648 */
void oracle_set_exposedvegp_filter(const ctsm_bounds_t* bounds, int num_nolakeurbanp, const int32_t* nolakeurbanp,
                                   const int32_t* frac_veg_nosno, int32_t* exposedvegp, int32_t* num_exposedvegp,
                                   int32_t* noexposedvegp, int32_t* num_noexposedvegp);
/* scalar pieces pinned by the reference's own unit tests (tests/test_oracle_golden.py) */
void oracle_qsat(double T, double p, double* qs, double* es, double* qsdT);
void oracle_moninobukini(double zetamaxstable, double ur, double thv, double dthv, double zldis, double z0m, double* um,
                         double* obu);
int oracle_quadratic(double a, double b, double c, double* r1, double* r2);
double oracle_plc(double x, double psi50, double ck);
double oracle_d1plc(double x, double psi50, double ck);

/* one clump = what one OpenMP thread of clm_drv's clump loop owns (clm_driver.F90:525-527) */
typedef struct oracle_clump_t {
  ctsm_bounds_t bounds;
  int32_t num_nolakep; const int32_t* filter_nolakep;
  int32_t num_nolakec; const int32_t* filter_nolakec;
  int32_t num_hydrologyc; const int32_t* filter_hydrologyc;
  int32_t num_exposedvegp; const int32_t* filter_exposedvegp;
} oracle_clump_t;
/* which: bit 2 CanopyFluxes, bit 0 SoilTemperature, bit 1 SoilWater; executed in clm_drv call order */
/* oracle_soilfluxes.c (SoilFluxesMod.F90:37-521 + p2c) */
int oracle_soilfluxes(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                      int num_nolakep, const int32_t* filter_nolakep, const ctsm_soilfluxes_fields_t* f, ctsm_status_t* st);
int oracle_patch2col(const ctsm_bounds_t* bounds, int num_allc, const int32_t* filter_allc, int num_nolakec,
                     const int32_t* filter_nolakec, const ctsm_patch2col_fields_t* f);
int oracle_begin_water_column_balance(const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                                      int num_lakec, const int32_t* filter_lakec,
                                      const ctsm_waterbalance_fields_t* f, double aquifer_water_baseline, ctsm_status_t* st);
int oracle_water_gridcell_balance(const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec, int num_lakec,
                                  const int32_t* filter_lakec, const ctsm_watergridbalance_fields_t* f,
                                  double aquifer_water_baseline, int flag_endwb, ctsm_status_t* st);
/* oracle_filters.c (filterMod.F90:303-592) */
int oracle_set_filters(const ctsm_bounds_t* bounds, const ctsm_filter_inputs_t* in, ctsm_filters_t* out);
void oracle_set_num_threads(int n);
void oracle_default_params(ctsm_params_t* p);
int oracle_step_clumps(const ctsm_params_t* prm, int nclumps, const oracle_clump_t* clumps,
                       const ctsm_soiltemperature_fields_t* ft, const ctsm_soilwater_fields_t* fw,
                       const ctsm_canopyfluxes_fields_t* fc, int which);
int oracle_fullstep_clumps(const ctsm_params_t* prm, int nclumps, const oracle_clump_t* clumps,
                           const ctsm_soiltemperature_fields_t* ft, const ctsm_soilwater_fields_t* fw,
                           const ctsm_canopyfluxes_fields_t* fc, const ctsm_plantsink_fields_t* fs,
                           const ctsm_balancecheck_fields_t* fb, const ctsm_soilfluxes_fields_t* fx,
                           const ctsm_patch2col_fields_t* f2c, int DAnstep, int which);

/* number of OpenMP threads the clump-loop drivers will use */
int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
