/* oracle_canopy.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Shared context + indexing macros of the CanopyFluxes / PHS restatement
 * (oracle_canopy.c, oracle_phs.c). */
#ifndef CTSM_ORACLE_CANOPY_H
#define CTSM_ORACLE_CANOPY_H
#include "oracle.h"

#define NLEVSNO CTSM_NLEVSNO
#define NLEVGRND CTSM_NLEVGRND
#define NLEVSOI CTSM_NLEVSOI
#define SNOSOI_LO (-NLEVSNO + 1)
#define NPFT (x->prm->npft_table)      /* runtime: (mxpft+1) x parameter-set members */

/* shr_const_mod.F90:16-53, clm_varcon.F90:41-125 (spelled as the reference spells them) */
static const double rpi = 3.14159265358979323846;
static const double rgas = 6.02214e26 * 1.38065e-23;    /* SHR_CONST_AVOGAD*SHR_CONST_BOLTZ */
static const double tfrz = 273.15;

typedef struct cf_ctx {
  const ctsm_canopyfluxes_fields_t* f;
  const ctsm_params_t* prm;
  int begp0, begc0, begg0;      /* allocation lower bounds */
  size_t ldp, ldc;
  int np;                       /* endp0 - begp0 + 1 */
  double *bbb, *mbb;            /* photosyns_inst%bbb_patch / mbb_patch (Ball-Berry only), (begp0:endp0) */
  int err_code, err_index;      /* first endrun site hit */
  int n_warnings;
} cf_ctx;

#define P1(name, p) (x->f->name[(p) - x->begp0])
#define P2(name, p, j, lo) (x->f->name[(size_t)((j) - (lo)) * x->ldp + ((p) - x->begp0)])
#define C1(name, c) (x->f->name[(c) - x->begc0])
#define C2(name, c, j, lo) (x->f->name[(size_t)((j) - (lo)) * x->ldc + ((c) - x->begc0)])
#define G1(name, g) (x->f->name[(g) - x->begg0])
#define PFT(name, ivt) (x->f->name[(ivt)])
#define PFTV(name, ivt, lvl) (x->f->name[(size_t)((lvl) - 1) * NPFT + (ivt)])

/* outputs of FrictionVelocity for one point (FrictionVelocityMod.F90:754-1117) */
typedef struct oracle_fricvel_t { double ustar, temp1, temp2, temp12m, temp22m, fm, vds, u10_clm, va, u10, fv; } oracle_fricvel_t;
void oracle_friction_velocity_point(double hgt_u, double hgt_t, double hgt_q, double displa, double z0m, double z0h, double z0q,
                                    double obu, int iter, double ur, double um, oracle_fricvel_t* o);
void oracle_qsat(double T, double p, double* qs, double* es, double* qsdT);
void oracle_moninobukini(double zetamaxstable, double ur, double thv, double dthv, double zldis, double z0m, double* um,
                         double* obu);
int oracle_quadratic(double a, double b, double c, double* r1, double* r2);
double oracle_plc(double x, double psi50, double ck);
double oracle_d1plc(double x, double psi50, double ck);
double oracle_wet_bulbs(double tc, double rh);
void oracle_photosyns_timestepinit(cf_ctx* x, const ctsm_bounds_t* bounds);
void oracle_photosynthesis_total(cf_ctx* x, int fn, const int32_t* filterp);
void oracle_photosynthesis_hydraulic_stress(cf_ctx* x, int fn, const int32_t* filterp, const double* esat_tv,
                                            const double* eair, const double* oair, const double* cair,
                                            const double* rb, double* bsun, double* bsha, double* btran,
                                            const double* dayl_factor, const double* qsatl, const double* qaf);
int oracle_phs_calcstress(const ctsm_params_t* prm, const ctsm_canopyfluxes_fields_t* fld, int p, double* xv4, double* bsun,
                          double* bsha, double gb_mol, double gs_mol_sun, double gs_mol_sha, double qsatl, double qaf);
int oracle_phs_standalone(const ctsm_params_t* prm, const ctsm_canopyfluxes_fields_t* fld, int fn, const int32_t* filterp,
                          const double* esat_tv, const double* eair, const double* oair, const double* cair, const double* rb,
                          double* bsun, double* bsha, double* btran, const double* dayl_factor, const double* qsatl,
                          const double* qaf);
void oracle_photosynthesis(cf_ctx* x, int fn, const int32_t* filterp, const double* esat_tv, const double* eair,
                           const double* oair, const double* cair, const double* rb, const double* btran,
                           const double* dayl_factor, int phase);
#endif
