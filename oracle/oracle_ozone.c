/* oracle_ozone.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the ozone routines (src/biogeophys/OzoneMod.F90):
 *   CalcOzoneUptake :356-426 with CalcOzoneUptakeOnePoint :429-511 (called from CanopyFluxes, CanopyFluxesMod.F90:1690)
 *   CalcOzoneStress :514-548 = CalcOzoneStressLombardozzi2015 :551-671 or CalcOzoneStressFalk :675-782 (clm_driver.F90:690)
 * Pinned by the NumPy restatement in tests/test_oracle_ozone.py.
 */
#include <math.h>
#include <string.h>
#include "oracle.h"

static const double ko3 = 1.67, lai_thresh = 0.5, o3_flux_threshold = 0.8;                     /* :95-101 */
static const double SHR_CONST_RGAS = 6.02214e26 * 1.38065e-23;

/* CalcOzoneUptakeOnePoint :470-509 */
static double uptake_one_point(double forc_ozone, double forc_pbot, double forc_th, double rs, double rb, double ram, double tlai,
                               double tlai_old, double evergreen, double leaf_long, int dtime, double o3uptake) {
  const double o3concnmolm3 = forc_ozone * 1.e9 * (forc_pbot / (forc_th * SHR_CONST_RGAS * 0.001));
  const double o3flux = o3concnmolm3 / (ko3 * rs + rb + ram);
  double o3fluxcrit;
  if (o3flux < o3_flux_threshold) o3fluxcrit = 0.0;
  else o3fluxcrit = o3flux - o3_flux_threshold;
  const double dtimeh = dtime / 3600.0;
  const double o3fluxperdt = o3fluxcrit * dtime * 0.000001;
  if (tlai > lai_thresh) {
    double heal, leafturn;
    if (tlai - tlai_old > 0) heal = fmax(0.0, (((tlai - tlai_old) / tlai) * o3fluxperdt));
    else heal = 0.0;
    if (evergreen == 1) leafturn = 1.0 / (leaf_long * 365.0 * 24.0);
    else leafturn = 0.0;
    const double decay = o3uptake * leafturn * dtimeh;
    return fmax(0.0, o3uptake + o3fluxperdt - decay - heal);
  }
  return 0.0;
}

int oracle_calc_ozone_uptake(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_exposedvegp,
                             const int32_t* filter_exposedvegp, const ctsm_ozone_fields_t* f, ctsm_status_t* st) {
  (void)bounds;
  if (st) memset(st, 0, sizeof *st);
  const int begp0 = f->alloc.begp, begc0 = f->alloc.begc, begg0 = f->alloc.begg;
  const int dtime = (int)prm->dtime;
  for (int fp = 0; fp < num_exposedvegp; ++fp) {
    const int p = filter_exposedvegp[fp], pp = p - begp0;
    const int c = f->column[pp] - begc0, g = f->gridcell[pp] - begg0, t = f->itype[pp];
    f->o3uptakesha[pp] = uptake_one_point(f->forc_o3[g], f->forc_pbot[c], f->forc_th[c], f->rssha[pp], f->rb1[pp], f->ram1[pp],
                                          f->tlai[pp], f->tlai_old[pp], f->pft_evergreen[t], f->pft_leaf_long[t], dtime, f->o3uptakesha[pp]);
    f->o3uptakesun[pp] = uptake_one_point(f->forc_o3[g], f->forc_pbot[c], f->forc_th[c], f->rssun[pp], f->rb1[pp], f->ram1[pp],
                                          f->tlai[pp], f->tlai_old[pp], f->pft_evergreen[t], f->pft_leaf_long[t], dtime, f->o3uptakesun[pp]);
    f->tlai_old[pp] = f->tlai[pp];
  }
  return 0;
}

/* the intercept / slope tables :103-133: rows needleleaf (pft_type <= 3), broadleaf (woody), nonwoody */
static const double photoInt[3] = {0.8390, 0.8752, 0.8021}, photoSlope[3] = {0.0, 0.0, -0.0009};
static const double condInt[3] = {0.7823, 0.9125, 0.7511}, condSlope[3] = {0.0048, 0.0, 0.0};
static const double jmaxInt[3] = {1.0, 1.0, 1.0}, jmaxSlope[3] = {0.0, -0.0037, 0.0};

static int plant_class(int pft_type, double woody) { return pft_type > 3 ? (woody == 0 ? 2 : 1) : 0; }

int oracle_calc_ozone_stress(const ctsm_bounds_t* bounds, int num_exposedvegp, const int32_t* filter_exposedvegp, int num_noexposedvegp,
                             const int32_t* filter_noexposedvegp, int stress_method, int is_time_to_run_luna,
                             const ctsm_ozone_fields_t* f, ctsm_status_t* st) {
  (void)bounds;
  if (st) memset(st, 0, sizeof *st);
  const int begp0 = f->alloc.begp;
  if (stress_method == 1) {                                                   /* Lombardozzi2015 :581-607, :636-669 */
    for (int fp = 0; fp < num_exposedvegp; ++fp) {
      const int pp = filter_exposedvegp[fp] - begp0, t = f->itype[pp];
      const int k = plant_class(t, f->pft_woody[t]);
      const double u[2] = {f->o3uptakesha[pp], f->o3uptakesun[pp]};
      double* v[2] = {&f->o3coefvsha[pp], &f->o3coefvsun[pp]};
      double* gq[2] = {&f->o3coefgsha[pp], &f->o3coefgsun[pp]};
      for (int i = 0; i < 2; ++i) {
        if (u[i] == 0.0) { *v[i] = 1.0; *gq[i] = 1.0; }
        else {
          *v[i] = fmax(0.0, fmin(1.0, photoInt[k] + photoSlope[k] * u[i]));
          *gq[i] = fmax(0.0, fmin(1.0, condInt[k] + condSlope[k] * u[i]));
        }
      }
    }
    for (int fp = 0; fp < num_noexposedvegp; ++fp) {
      const int pp = filter_noexposedvegp[fp] - begp0;
      f->o3coefvsha[pp] = 1.0; f->o3coefgsha[pp] = 1.0; f->o3coefvsun[pp] = 1.0; f->o3coefgsun[pp] = 1.0;
    }
  } else if (stress_method == 2) {                                            /* Falk :699-735, :759-780 */
    if (!is_time_to_run_luna) return 0;
    for (int fp = 0; fp < num_exposedvegp; ++fp) {
      const int pp = filter_exposedvegp[fp] - begp0, t = f->itype[pp];
      const int k = plant_class(t, f->pft_woody[t]);
      f->o3coefjmaxsha[pp] = f->o3uptakesha[pp] == 0.0 ? 1.0 : fmax(0.0, fmin(1.0, jmaxInt[k] + jmaxSlope[k] * f->o3uptakesha[pp]));
      f->o3coefjmaxsun[pp] = f->o3uptakesun[pp] == 0.0 ? 1.0 : fmax(0.0, fmin(1.0, jmaxInt[k] + jmaxSlope[k] * f->o3uptakesun[pp]));
    }
    for (int fp = 0; fp < num_noexposedvegp; ++fp) {
      const int pp = filter_noexposedvegp[fp] - begp0;
      f->o3coefjmaxsha[pp] = 1.0; f->o3coefjmaxsun[pp] = 1.0;
    }
  } else {
    if (st) { st->code = CTSM_ERR_BAD_ARG; }
    return CTSM_ERR_BAD_ARG;
  }
  return 0;
}
