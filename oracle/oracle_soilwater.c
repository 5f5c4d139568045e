/* oracle_soilwater.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of SoilWater -> soilwater_moisture_form and its callees
 * (src/biogeophys/SoilWaterMovementMod.F90:240-370, 976-1440, 1444-1556,
 * 1560-1866, 1871-1937, 1941-2012, 2139-2160) with the Clapp-Hornberger 1978
 * retention curve (SoilWaterRetentionCurveClappHornberg1978Mod.F90:75-79,115-120),
 * for the default configuration upper_boundary_condition = bc_flux,
 * lower_boundary_condition in {bc_zero_flux, bc_flux}, tridiagSolution = lapack,
 * use_flexibleCN = .false., no aquifer layer.
 * PARITY UNPINNED by the reference's own tests; invariants
 * (mass conservation, hydrostatic equilibrium) in tests/test_oracle_soilwater.py.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#ifdef _OPENMP
#include <omp.h>
#else
static int omp_in_parallel(void) { return 0; }
#endif

enum { BC_HEAD = 0, BC_FLUX = 1, BC_ZERO_FLUX = 2, BC_WATERTABLE = 3 };   /* :55-58 */
#define NLEVSNO CTSM_NLEVSNO
#define NLEVSOI CTSM_NLEVSOI

static const double denh2o = 1.000e3;     /* SHR_CONST_RHOFW, shr_const_mod.F90 */
static const double m_to_mm = 1.e3;       /* :65 */

typedef struct {
  const ctsm_soilwater_fields_t* f;
  const ctsm_params_t* prm;
  int begc0;
  size_t ldc;
} sw_ctx;

#define SNOSOI_LO (-NLEVSNO + 1)
#define F2(name, c, j, lo) (x->f->name[(size_t)((j) - (lo)) * x->ldc + ((c) - x->begc0)])
#define F1(name, c) (x->f->name[(c) - x->begc0])

/* SoilWaterRetentionCurveClappHornberg1978Mod.F90:75-79 soil_hk */
static void soil_hk(const sw_ctx* x, int c, int j, double s, double imped, double* hk, double* dhkds) {
  const double hksat = F2(hksat, c, j, 1), bsw = F2(bsw, c, j, 1);
  *hk = imped * hksat * pow(s, 2.0 * bsw + 3.0);
  if (dhkds) *dhkds = (2.0 * bsw + 3.0) * (*hk) / s;
}
/* :115-120 soil_suction */
static void soil_suction(const sw_ctx* x, int c, int j, double s, double* smp, double* dsmpds) {
  const double bsw = F2(bsw, c, j, 1), sucsat = F2(sucsat, c, j, 1);
  *smp = -sucsat * pow(s, -bsw);
  if (dsmpds) *dsmpds = -bsw * (*smp) / s;
}
/* SoilWaterMovementMod.F90:2139-2160 */
static void IceImpedance(const sw_ctx* x, double icefrac, double* imped) {
  *imped = pow(10.0, -x->prm->e_ice * icefrac);
}

/* :1444-1556.  1-based local vectors of length nlayers (index 0 unused). */
static void compute_hydraulic_properties(const sw_ctx* x, int c, int nlayers, const double* vwc_liq,
                                         double* hk, double* smp, double* dhkdw, double* dsmpdw, double* imped) {
  double s2[NLEVSOI + 2];
  for (int j = 1; j <= nlayers; ++j) { hk[j] = 0; smp[j] = 0; dhkdw[j] = 0; dsmpdw[j] = 0; imped[j] = 0; }
  for (int j = 1; j <= nlayers; ++j) {                      /* :1511-1516 */
    s2[j] = vwc_liq[j] / F2(watsat, c, j, 1);
    s2[j] = fmin(s2[j], 1.0);
    s2[j] = fmax(0.01, s2[j]);
  }
  for (int j = 1; j <= nlayers; ++j) {                      /* :1518-1553 */
    double s1, dhkds, dsmpds;
    if (j == nlayers) {
      s1 = s2[j];
      IceImpedance(x, F2(icefrac, c, j, 1), &imped[j]);
    } else {
      s1 = 0.5 * (s2[j] + s2[j + 1]);
      IceImpedance(x, 0.5 * (F2(icefrac, c, j, 1) + F2(icefrac, c, j + 1, 1)), &imped[j]);
    }
    s1 = fmin(s1, 1.0);
    s1 = fmax(0.01, s1);
    soil_hk(x, c, j, s1, imped[j], &hk[j], &dhkds);
    soil_suction(x, c, j, s2[j], &smp[j], &dsmpds);
    dhkdw[j] = dhkds;
    dsmpdw[j] = dsmpds / F2(watsat, c, j, 1);
    F2(smp_l, c, j, 1) = smp[j];
    F2(hk_l, c, j, 1) = hk[j];
  }
}

/* :1560-1866 for bc_flux upper and bc_flux / bc_zero_flux lower boundary. */
static int compute_moisture_fluxes_and_derivs(const sw_ctx* x, int c, int nlayers, const double* hk,
                                              const double* smp, const double* dhkdw, const double* dsmpdw,
                                              double* qin, double* qout, double* dqidw0, double* dqidw1,
                                              double* dqodw1, double* dqodw2) {
  double num, den, dhkds1, dhkds2;
  for (int j = 1; j <= nlayers; ++j) { qin[j] = 0; qout[j] = 0; dqidw0[j] = 0; dqidw1[j] = 0; dqodw1[j] = 0; dqodw2[j] = 0; }
  int j = 1;
  switch (x->prm->upper_boundary_condition) {               /* :1648-1686 */
    case BC_FLUX:
      qin[j] = F1(qflx_infl, c);
      dqidw1[j] = 0.0;
      break;
    default:
      return CTSM_ERR_BAD_ARG;   /* bc_head needs vwc_liq_ub, which the reference never sets */
  }
  dhkds1 = 0.5 * dhkdw[j] / F2(watsat, c, j, 1);            /* :1697-1698 */
  dhkds2 = 0.5 * dhkdw[j] / F2(watsat, c, j + 1, 1);
  num = (smp[j + 1] - smp[j]);                              /* :1709-1715 */
  den = m_to_mm * (F2(z, c, j + 1, SNOSOI_LO) - F2(z, c, j, SNOSOI_LO));
  qout[j] = -hk[j] * num / den + hk[j];
  dqodw1[j] = (hk[j] * dsmpdw[j] - dhkds1 * num) / den + dhkds1;
  dqodw2[j] = (-hk[j] * dsmpdw[j + 1] - dhkds2 * num) / den + dhkds2;
  for (j = 2; j <= nlayers - 1; ++j) {                      /* :1719-1751 */
    qin[j] = qout[j - 1];
    dqidw0[j] = dqodw1[j - 1];
    dqidw1[j] = dqodw2[j - 1];
    dhkds1 = 0.5 * dhkdw[j] / F2(watsat, c, j, 1);
    dhkds2 = 0.5 * dhkdw[j] / F2(watsat, c, j + 1, 1);
    num = (smp[j + 1] - smp[j]);
    den = m_to_mm * (F2(z, c, j + 1, SNOSOI_LO) - F2(z, c, j, SNOSOI_LO));
    qout[j] = -hk[j] * num / den + hk[j];
    dqodw1[j] = (hk[j] * dsmpdw[j] - dhkds1 * num) / den + dhkds1;
    dqodw2[j] = (-hk[j] * dsmpdw[j + 1] - dhkds2 * num) / den + dhkds2;
  }
  j = nlayers;                                              /* :1755-1762 */
  qin[j] = qout[j - 1];
  dqidw0[j] = dqodw1[j - 1];
  dqidw1[j] = dqodw2[j - 1];
  switch (x->prm->lower_boundary_condition) {               /* :1764-1860 */
    case BC_FLUX:
      qout[j] = hk[j];
      dqodw1[j] = dhkdw[j] / F2(watsat, c, j, 1);
      break;
    case BC_ZERO_FLUX:
      qout[j] = 0.;
      dqodw1[j] = 0.;
      break;
    default:
      return CTSM_ERR_BAD_ARG;
  }
  return 0;
}

/* :1871-1937 */
static void compute_RHS_moisture_form(int nlayers, const double* vert_trans_sink, const double* qin,
                                      const double* qout, const double* dt_dz, double* rmx) {
  for (int j = 1; j <= nlayers; ++j) {
    const double fluxNet = qin[j] - qout[j] - vert_trans_sink[j];
    rmx[j] = -fluxNet * dt_dz[j];
  }
}

/* :1941-2012 */
static void compute_LHS_moisture_form(int nlayers, const double* dt_dz, const double* dqidw0,
                                      const double* dqidw1, const double* dqodw1, const double* dqodw2,
                                      double* amx, double* bmx, double* cmx) {
  int j = 1;
  amx[j] = 0.0;
  bmx[j] = -1.0 - (-dqidw1[j] + dqodw1[j]) * dt_dz[j];
  cmx[j] = -dqodw2[j] * dt_dz[j];
  for (j = 2; j <= nlayers - 1; ++j) {
    amx[j] = dqidw0[j] * dt_dz[j];
    bmx[j] = -1.0 - (-dqidw1[j] + dqodw1[j]) * dt_dz[j];
    cmx[j] = -dqodw2[j] * dt_dz[j];
  }
  j = nlayers;
  amx[j] = dqidw0[j] * dt_dz[j];
  bmx[j] = -1.0 - (-dqidw1[j] + dqodw1[j]) * dt_dz[j];
  cmx[j] = 0.0;
}

/* One column of the main spatial loop, :1180-1420.  Returns 0 or an error code. */
static int soilwater_column(const sw_ctx* x, int c, int* lapack_err) {
  const ctsm_params_t* prm = x->prm;
  const double dtime = prm->dtime;
  enum { N = NLEVSOI + 2 };
  double hk[N], smp[N], dhkdw[N], dsmpdw[N], imped[N], vwc_liq[N], dt_dz[N];
  double qin[N], qout[N], dqidw0[N], dqidw1[N], dqodw1[N], dqodw2[N], dwat[N];
  double amx[N], bmx[N], cmx[N], rmx[N], sink[N];
  double dLow[N], dUpp[N], diag[N], rhs[N], fluxNet0[N], fluxNet1[N];

  const int nlayers = F1(nbedrock, c);                      /* :1184 */
  int nsubstep = 0;                                         /* :1187 */
  double dtsub = dtime;                                     /* :1190 */
  double dtdone = 0.0;
  F1(qcharge, c) = 0.0;                                     /* :1194 */
  for (int j = 1; j <= nlayers; ++j) sink[j] = F2(qflx_rootsoi, c, j, 1);

  for (;;) {                                                /* :1197 */
    nsubstep = nsubstep + 1;
    for (int j = 1; j <= nlayers; ++j) {                    /* :1203-1206 */
      vwc_liq[j] = fmax(F2(h2osoi_liq, c, j, SNOSOI_LO), 1.0e-6) / (F2(dz, c, j, SNOSOI_LO) * denh2o);
      dt_dz[j] = dtsub / (m_to_mm * F2(dz, c, j, SNOSOI_LO));
    }
    compute_hydraulic_properties(x, c, nlayers, vwc_liq, hk, smp, dhkdw, dsmpdw, imped);
    int rc = compute_moisture_fluxes_and_derivs(x, c, nlayers, hk, smp, dhkdw, dsmpdw, qin, qout,
                                                dqidw0, dqidw1, dqodw1, dqodw2);
    if (rc) return rc;
    compute_RHS_moisture_form(nlayers, sink, qin, qout, dt_dz, rmx);
    compute_LHS_moisture_form(nlayers, dt_dz, dqidw0, dqidw1, dqodw1, dqodw2, amx, bmx, cmx);

    /* :1279-1299 lapack solution */
    for (int j = 1; j <= nlayers - 1; ++j) dLow[j - 1] = amx[j + 1];
    for (int j = 1; j <= nlayers; ++j) diag[j - 1] = bmx[j];
    for (int j = 1; j <= nlayers - 1; ++j) dUpp[j - 1] = cmx[j];
    for (int j = 1; j <= nlayers; ++j) rhs[j - 1] = rmx[j];
    int err = 0;
    oracle_dgtsv(nlayers, 1, dLow, diag, dUpp, rhs, nlayers, &err);
    if (err != 0) { *lapack_err = err; return CTSM_ERR_DGTSV; }
    for (int j = 1; j <= nlayers; ++j) dwat[j] = rhs[j - 1];

    /* :1307-1348 error estimation */
    for (int j = 1; j <= nlayers; ++j) {
      if (prm->flux_calculation == 42) {
        double qin_test, qout_test;
        if (j == 1) qin_test = qin[j] + dqidw1[j] * dwat[j];
        else qin_test = qin[j] + dqidw0[j] * dwat[j - 1] + dqidw1[j] * dwat[j];
        if (j == nlayers) qout_test = qout[j] + dqodw1[j] * dwat[j];
        else qout_test = qout[j] + dqodw1[j] * dwat[j] + dqodw2[j] * dwat[j + 1];
        fluxNet0[j] = qin_test - qout_test - sink[j];
      } else {
        fluxNet0[j] = dwat[j] / dt_dz[j];
      }
      fluxNet1[j] = qin[j] - qout[j] - sink[j];
    }
    double errorMax = -HUGE_VAL;                            /* :1345-1346 maxval */
    for (int j = 1; j <= nlayers; ++j) {
      const double e = fabs(fluxNet1[j] - fluxNet0[j]) * dtsub * 0.5;
      if (e > errorMax) errorMax = e;
    }
    if (errorMax > prm->xTolerUpper && dtsub > prm->dtmin) {   /* :1349-1353 */
      dtsub = fmax(dtsub / 2.0, prm->dtmin);
      continue;
    }
    for (int j = 1; j <= nlayers; ++j)                      /* :1360-1362 */
      F2(h2osoi_liq, c, j, SNOSOI_LO) = F2(h2osoi_liq, c, j, SNOSOI_LO) + dwat[j] * (m_to_mm * F2(dz, c, j, SNOSOI_LO));
    double qcTemp;                                          /* :1365-1383 */
    switch (prm->lower_boundary_condition) {
      case BC_FLUX: qcTemp = hk[nlayers] + dhkdw[nlayers] * dwat[nlayers]; break;
      case BC_ZERO_FLUX: qcTemp = 0.0; break;
      default: return CTSM_ERR_BAD_ARG;
    }
    F1(qcharge, c) = F1(qcharge, c) + qcTemp * (dtsub / dtime);   /* :1386 */
    dtdone = dtdone + dtsub;                                /* :1389-1390 */
    if (fabs(dtime - dtdone) < prm->verySmall) break;
    if (errorMax < prm->xTolerLower) dtsub = dtsub * 2.0;   /* :1393-1395 */
    dtsub = fmin(dtsub, dtime - dtdone);                    /* :1398 */
  }
  F1(num_substeps, c) = (double)nsubstep;                   /* :1403 */
  for (int j = nlayers; j >= 2; --j) {                      /* :1406-1410 */
    const double cap = F2(eff_porosity, c, j, 1) * m_to_mm * F2(dz, c, j, SNOSOI_LO);
    const double over_saturation = fmax(F2(h2osoi_liq, c, j, SNOSOI_LO) - cap, 0.0);
    F2(h2osoi_liq, c, j, SNOSOI_LO) = fmin(cap, F2(h2osoi_liq, c, j, SNOSOI_LO));
    F2(h2osoi_liq, c, j - 1, SNOSOI_LO) = F2(h2osoi_liq, c, j - 1, SNOSOI_LO) + over_saturation;
  }
  for (int j = 1; j <= nlayers; ++j) {                      /* :1419-1420 */
    F2(qin, c, j, 1) = qin[j];
    F2(qout, c, j, 1) = qout[j];
  }
  return 0;
}

int oracle_soilwater(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_hydrologyc,
                     const int32_t* filter_hydrologyc, const ctsm_soilwater_fields_t* f, ctsm_status_t* st) {
  (void)bounds;
  sw_ctx xx = { f, prm, f->alloc.begc, (size_t)(f->alloc.endc - f->alloc.begc + 1) };
  int first_fc = num_hydrologyc, first_rc = 0, first_err = 0;
  if (st) memset(st, 0, sizeof *st);
  /* The reference runs this column-outer loop inside one OpenMP thread per
   * clump (clm_driver.F90:525); columns are independent, so splitting the
   * filter across threads is the same decomposition. */
#pragma omp parallel for schedule(static) if (!omp_in_parallel())
  for (int fc = 0; fc < num_hydrologyc; ++fc) {
    int lerr = 0;
    const int rc = soilwater_column(&xx, filter_hydrologyc[fc], &lerr);
    if (rc) {
#pragma omp critical
      if (fc < first_fc) { first_fc = fc; first_rc = rc; first_err = lerr; }
    }
  }
  if (first_rc && st) {
    st->code = first_rc; st->subgrid_level = CTSM_SUBGRID_COLUMN;
    st->subgrid_index = filter_hydrologyc[first_fc]; st->info = first_err;
    snprintf(st->msg, sizeof st->msg, "%s", first_rc == CTSM_ERR_DGTSV
             ? "soilwater_moisture_form:: problem with the lapack solver"
             : "soilwater_moisture_form:: unsupported boundary condition");
  }
  return first_rc;
}
