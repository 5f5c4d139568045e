/* oracle_phs.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of PhotosynthesisHydraulicStress and its callees
 * (src/biogeophys/PhotosynthesisMod.F90):
 *   PhotosynthesisHydraulicStress :2704-3811   hybrid_PHS  :3815-4064
 *   brent_PHS   :4068-4223                     ci_func_PHS :4227-4486
 *   calcstress  :4490-4710                     spacA       :4715-4893
 *   spacF       :4898-4976                     getvegwp    :4979-5077
 *   getqflx     :5080-5164                     plc :5167   d1plc :5199
 *   TimeStepInit :1143-1201                    PhotosynthesisTotal :2065-2151
 * and quadratic (src/utils/quadraticMod.F90:17-74), for the configuration
 * use_cn = use_fates = use_c13 = .false., lnc_opt = .false., vcmax_opt = 0,
 * nlevcan = 1.  Patch loops, argument order and operation order follow the
 * Fortran.  plc / d1plc / quadratic are pinned by the reference's own unit
 * tests (tests/test_oracle_golden.py); everything else is PARITY UNPINNED.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_canopy.h"
#include "oracle_pert.h"

static const double spval = 1.e36;
static const double bbbopt_c3 = 10000.0, bbbopt_c4 = 40000.0;       /* :85-86 */
static const double medlyn_rh_can_max = 50.0, medlyn_rh_can_fact = 0.001;   /* :87-88 */
static const double max_cs = 1.e-06;                                 /* :89 */

/* quadraticMod.F90:17-74.  Returns 0, or CTSM_ERR_QUADRATIC where the reference calls endrun. */
int oracle_quadratic(double a, double b, double c, double* r1, double* r2) {
  double q, root;
  if (a == 0.0) return CTSM_ERR_QUADRATIC;
  root = b * b - 4.0 * a * c;
  if (root < 0.0) {
    if (-root < 3.0 * 2.220446049250313e-16) root = 0.0;   /* 3*epsilon(b) */
    else return CTSM_ERR_QUADRATIC;
  }
  if (b >= 0.0) q = -0.5 * (b + sqrt(root));
  else q = -0.5 * (b - sqrt(root));
  *r1 = q / a;
  if (q != 0.0) *r2 = c / q;
  else *r2 = 1.e36;
  return 0;
}

/* plc :5167-5195 with explicit parameters (the unit test calls it with pft 1, root segment) */
double oracle_plc(double x, double psi50, double ck) {
  double v = pow(2.0, -pow(x / psi50, ck));
  if (v < 0.005) v = 0.0;
  return v;
}
/* d1plc :5199-5228 */
double oracle_d1plc(double x, double psi50, double ck) {
  return -ck * log(2.0) * pow(2.0, -pow(x / psi50, ck)) * pow(x / psi50, ck) / x;
}

static double plc(const cf_ctx* x, double v, int p, int level) {
  const int ivt = P1(itype, p);
  return oracle_plc(v, PFTV(pft_psi50, ivt, level), PFTV(pft_ck, ivt, level));
}
static double d1plc(const cf_ctx* x, double v, int p, int level) {
  const int ivt = P1(itype, p);
  return oracle_d1plc(v, PFTV(pft_psi50, ivt, level), PFTV(pft_ck, ivt, level));
}
static double kmax(const cf_ctx* x, int p, int level) { return PFTV(pft_kmax, P1(itype, p), level); }

static void fail(cf_ctx* x, int code, int p) {
  if (x->err_code == 0) { x->err_code = code; x->err_index = p; }
}
#define QUAD(a, b, c, r1, r2) do { if (oracle_quadratic((a), (b), (c), (r1), (r2))) { fail(x, CTSM_ERR_QUADRATIC, p); *(r1) = 0.0; *(r2) = 0.0; } } while (0)

enum { sun = 1, sha = 2, xyl = 3, root = 4 };

/* getqflx :5080-5164 */
static void getqflx(const cf_ctx* x, int p, int c, double gb_mol, double* gs_mol_sun, double* gs_mol_sha,
                    double* qflx_sun, double* qflx_sha, double qsatl, double qaf, int havegs) {
  const double cf = C1(forc_pbot, c) / (rgas * 1.e-3 * P1(thm, p)) * 1.e6;
  const double wtl = (P1(elai, p) + P1(esai, p)) * gb_mol;
  const double efpot = C1(forc_rho, c) * wtl * (qsatl - qaf);
  if (havegs) {
    if ((efpot > 0.0) && (P1(elai, p) > 0.0)) {
      if (*gs_mol_sun > 0.0) {
        const double rppdry_sun = P1(fdry, p) / gb_mol * (P1(laisun, p) / (1.0 / gb_mol + 1.0 / *gs_mol_sun)) / P1(elai, p);
        *qflx_sun = efpot * rppdry_sun / cf;
      } else *qflx_sun = 0.0;
      if (*gs_mol_sha > 0.0) {
        const double rppdry_sha = P1(fdry, p) / gb_mol * (P1(laisha, p) / (1.0 / gb_mol + 1.0 / *gs_mol_sha)) / P1(elai, p);
        *qflx_sha = efpot * rppdry_sha / cf;
      } else *qflx_sha = 0.0;
    } else {
      *qflx_sun = 0.0;
      *qflx_sha = 0.0;
    }
  } else {
    if (*qflx_sun > 0.0)
      *gs_mol_sun = gb_mol * *qflx_sun * cf * P1(elai, p) / (efpot * P1(fdry, p) * P1(laisun, p) - *qflx_sun * cf * P1(elai, p));
    else *gs_mol_sun = 0.0;
    if (*qflx_sha > 0.0)
      *gs_mol_sha = gb_mol * *qflx_sha * cf * P1(elai, p) / (efpot * P1(fdry, p) * P1(laisha, p) - *qflx_sha * cf * P1(elai, p));
    else *gs_mol_sha = 0.0;
  }
}

static double sum_ksr(const cf_ctx* x, int p) {
  double s = 0.0;
  for (int j = 1; j <= NLEVSOI; ++j) s += P2(k_soil_root, p, j, 1);
  return s;
}

/* getvegwp :4979-5077 (xv is 1-based: xv[sun..root]) */
static void getvegwp(const cf_ctx* x, int p, int c, double* xv, double gb_mol, double* gs_mol_sun, double* gs_mol_sha,
                     double qsatl, double qaf, double* soilflux) {
  double qflx_sun = 0.0, qflx_sha = 0.0, grav2[NLEVSOI + 1];
  const double grav1 = 1000.0 * P1(htop, p);
  for (int j = 1; j <= NLEVSOI; ++j) grav2[j] = 1000.0 * C2(z, c, j, SNOSOI_LO);
  getqflx(x, p, c, gb_mol, gs_mol_sun, gs_mol_sha, &qflx_sun, &qflx_sha, qsatl, qaf, 1);
  const double sk = sum_ksr(x, p);
  if (fabs(sk) == 0.0) {
    double s = 0.0;
    for (int j = 1; j <= NLEVSOI; ++j) s += C2(smp_l, c, j, 1) - grav2[j];
    xv[root] = s / NLEVSOI;
  } else {
    double s = 0.0;
    for (int j = 1; j <= NLEVSOI; ++j) s += P2(k_soil_root, p, j, 1) * (C2(smp_l, c, j, 1) - grav2[j]);
    xv[root] = (s - qflx_sun - qflx_sha) / sk;
  }
  const double fr = plc(x, xv[root], p, root);
  if ((P1(tsai, p) > 0.0) && (fr > 0.0))
    xv[xyl] = xv[root] - grav1 - (qflx_sun + qflx_sha) / (fr * kmax(x, p, root) / P1(htop, p) * P1(tsai, p));
  else
    xv[xyl] = xv[root] - grav1;
  const double fx = plc(x, xv[xyl], p, xyl);
  if ((P1(laisha, p) > 0.0) && (fx > 0.0)) xv[sha] = xv[xyl] - (qflx_sha / (fx * kmax(x, p, xyl) * P1(laisha, p)));
  else xv[sha] = xv[xyl];
  if ((P1(laisun, p) > 0.0) && (fx > 0.0)) xv[sun] = xv[xyl] - (qflx_sun / (fx * kmax(x, p, xyl) * P1(laisun, p)));
  else xv[sun] = xv[xyl];
  *soilflux = 0.0;
  for (int j = 1; j <= NLEVSOI; ++j)
    *soilflux = *soilflux + P2(k_soil_root, p, j, 1) * (C2(smp_l, c, j, 1) - xv[root] - grav2[j]);
}

static const double tol_lai = .001;

/* workload statistics for DESIGN.md / tools/phs_stats.py: [0] calcstress calls, [1] Newton iterations,
 * [2] calcstress calls that hit itmax, [3] ci_func_PHS calls, [4] brent_PHS calls, [5] hybrid outer passes */
long long oracle_phs_counters[8];
int oracle_pert_mode = 1;      /* oracle_pert.h; unused in the plain build */
void oracle_set_pert_mode(int m) { oracle_pert_mode = m; }
long long oracle_phs_newton_hist[2][64];   /* [night][Newton iterations of one calcstress call] (workload statistics) */
static long long* per_patch_newton = 0;   /* optional (begp0:endp0) accumulator */
void oracle_phs_set_patch_counter(long long* p) { per_patch_newton = p; }

/* spacF :4898-4976 */
static void spacF(const cf_ctx* x, int p, int c, const double* xv, double* f, double qflx_sun, double qflx_sha) {
  const double grav1 = P1(htop, p) * 1000.0;
  double grav2[NLEVSOI + 1];
  for (int j = 1; j <= NLEVSOI; ++j) grav2[j] = C2(z, c, j, SNOSOI_LO) * 1000.0;
  const double fsto1 = plc(x, xv[sun], p, sun);
  const double fsto2 = plc(x, xv[sha], p, sha);
  const double fx = plc(x, xv[xyl], p, xyl);
  const double fr = plc(x, xv[root], p, root);
  const double laisun = P1(laisun, p), laisha = P1(laisha, p), tsai = P1(tsai, p), htop = P1(htop, p);
  f[sun] = qflx_sun * fsto1 - laisun * kmax(x, p, sun) * fx * (xv[xyl] - xv[sun]);
  f[sha] = qflx_sha * fsto2 - laisha * kmax(x, p, sha) * fx * (xv[xyl] - xv[sha]);
  f[xyl] = laisun * kmax(x, p, sun) * fx * (xv[xyl] - xv[sun])
         + laisha * kmax(x, p, sha) * fx * (xv[xyl] - xv[sha])
         - tsai * kmax(x, p, xyl) / htop * fr * (xv[root] - xv[xyl] - grav1);
  double s1 = 0.0, s2 = 0.0;
  for (int j = 1; j <= NLEVSOI; ++j) s1 += P2(k_soil_root, p, j, 1) * (xv[root] + grav2[j]);
  for (int j = 1; j <= NLEVSOI; ++j) s2 += P2(k_soil_root, p, j, 1) * C2(smp_l, c, j, 1);
  f[root] = tsai * kmax(x, p, xyl) / htop * fr * (xv[root] - xv[xyl] - grav1) + s1 - s2;
  if (laisha < tol_lai) {
    const double temp = f[sun];
    f[sun] = f[sha];
    f[sha] = temp;
  }
}

/* spacA :4715-4893.  A, invA are 1-based [5][5]. */
static void spacA(const cf_ctx* x, int p, int c, const double* xv, double invA[5][5], double qflx_sun, double qflx_sha,
                  int* flag) {
  (void)c;
  double A[5][5];
  memset(A, 0, sizeof A);
  memset(invA, 0, 25 * sizeof(double));
  const double laisun = P1(laisun, p), laisha = P1(laisha, p), tsai = P1(tsai, p), htop = P1(htop, p);
  const double grav1 = htop * 1000.0;
  const double fx = plc(x, xv[xyl], p, xyl);
  const double fr = plc(x, xv[root], p, root);
  const double dfsto1 = d1plc(x, xv[sun], p, sun);
  const double dfsto2 = d1plc(x, xv[sha], p, sha);
  const double dfx = d1plc(x, xv[xyl], p, xyl);
  const double dfr = d1plc(x, xv[root], p, root);
  const double ksun = kmax(x, p, sun), ksha = kmax(x, p, sha), kxyl = kmax(x, p, xyl);
  A[1][1] = -laisun * ksun * fx - qflx_sun * dfsto1;
  A[1][3] = laisun * ksun * dfx * (xv[xyl] - xv[sun]) + laisun * ksun * fx;
  A[2][2] = -laisha * ksha * fx - qflx_sha * dfsto2;
  A[2][3] = laisha * ksha * dfx * (xv[xyl] - xv[sha]) + laisha * ksha * fx;
  A[3][1] = laisun * ksun * fx;
  A[3][2] = laisha * ksha * fx;
  A[3][3] = -laisun * ksun * dfx * (xv[xyl] - xv[sun]) - laisun * ksun * fx
            - laisha * ksha * dfx * (xv[xyl] - xv[sha]) - laisha * ksha * fx
            - tsai * kxyl / htop * fr;
  A[3][4] = tsai * kxyl / htop * dfr * (xv[root] - xv[xyl] - grav1) + tsai * kxyl / htop * fr;
  A[4][3] = tsai * kxyl / htop * fr;
  A[4][4] = -tsai * kxyl / htop * fr - tsai * kxyl / htop * dfr * (xv[root] - xv[xyl] - grav1) - sum_ksr(x, p);
  const double invfactor = 1.0;
  for (int i = 1; i <= 4; ++i) for (int j = 1; j <= 4; ++j) A[i][j] = invfactor * A[i][j];
  if (laisun > tol_lai && laisha > tol_lai) {
    const double determ = A[4][4] * A[2][2] * A[3][3] * A[1][1] - A[4][4] * A[2][2] * A[3][1] * A[1][3]
                        - A[4][4] * A[3][2] * A[2][3] * A[1][1] - A[4][3] * A[1][1] * A[2][2] * A[3][4];
    if (fabs(determ) <= 1.e-50) { *flag = 1; return; }
    *flag = 0;
    const double leading = 1.0 / determ;
    invA[1][1] = leading * A[4][4] * A[2][2] * A[3][3] - leading * A[4][4] * A[3][2] * A[2][3] - leading * A[4][3] * A[2][2] * A[3][4];
    invA[2][1] = leading * A[2][3] * A[4][4] * A[3][1];
    invA[3][1] = -leading * A[4][4] * A[2][2] * A[3][1];
    invA[4][1] = leading * A[4][3] * A[2][2] * A[3][1];
    invA[1][2] = leading * A[1][3] * A[4][4] * A[3][2];
    invA[2][2] = leading * A[4][4] * A[3][3] * A[1][1] - leading * A[4][4] * A[3][1] * A[1][3] - leading * A[4][3] * A[1][1] * A[3][4];
    invA[3][2] = -leading * A[1][1] * A[4][4] * A[3][2];
    invA[4][2] = leading * A[4][3] * A[1][1] * A[3][2];
    invA[1][3] = -leading * A[1][3] * A[2][2] * A[4][4];
    invA[2][3] = -leading * A[2][3] * A[1][1] * A[4][4];
    invA[3][3] = leading * A[2][2] * A[1][1] * A[4][4];
    invA[4][3] = -leading * A[4][3] * A[1][1] * A[2][2];
    invA[1][4] = leading * A[1][3] * A[3][4] * A[2][2];
    invA[2][4] = leading * A[2][3] * A[3][4] * A[1][1];
    invA[3][4] = -leading * A[3][4] * A[1][1] * A[2][2];
    invA[4][4] = leading * A[2][2] * A[3][3] * A[1][1] - leading * A[2][2] * A[3][1] * A[1][3] - leading * A[3][2] * A[2][3] * A[1][1];
    for (int i = 1; i <= 4; ++i) for (int j = 1; j <= 4; ++j) invA[i][j] = invfactor * invA[i][j];
  } else {
    if (laisha <= tol_lai) {
      A[2][2] = A[1][1];
      A[3][2] = A[3][1];
      A[2][3] = A[1][3];
    }
    const double determ = A[2][2] * A[3][3] * A[4][4] - A[3][4] * A[2][2] * A[4][3] - A[2][3] * A[3][2] * A[4][4];
    if (fabs(determ) <= 1.e-50) { *flag = 1; return; }
    *flag = 0;
    invA[2][2] = A[3][3] * A[4][4] - A[3][4] * A[4][3];
    invA[2][3] = -A[2][3] * A[4][4];
    invA[2][4] = A[3][4] * A[2][3];
    invA[3][2] = -A[3][2] * A[4][4];
    invA[3][3] = A[2][2] * A[4][4];
    invA[3][4] = -A[3][4] * A[2][2];
    invA[4][2] = A[3][2] * A[4][3];
    invA[4][3] = -A[2][2] * A[4][3];
    invA[4][4] = A[2][2] * A[3][3] - A[2][3] * A[3][2];
    const double rd = 1.0 / determ;
    for (int i = 1; i <= 4; ++i) for (int j = 1; j <= 4; ++j) invA[i][j] = rd * invA[i][j];
  }
}

/* calcstress :4490-4710 */
static void calcstress(cf_ctx* x, int p, int c, double* xv, double* bsun, double* bsha, double gb_mol,
                       double gs_mol_sun, double gs_mol_sha, double qsatl, double qaf) {
  const int itmax = 50;
  const double tolf = 1.e-6, toldx = 1.e-9;
  double A[5][5], f[5], dx[5], qflx_sun = 0.0, qflx_sha = 0.0, soilflux;
  int night, flag = 0, iter;
  const double laisun = P1(laisun, p), laisha = P1(laisha, p);
  if (xv[sun] > 0.0) { night = 1; xv[sun] = xv[sha]; }
  else night = 0;
  double gs0sun = gs_mol_sun, gs0sha = gs_mol_sha;
  getqflx(x, p, c, gb_mol, &gs0sun, &gs0sha, &qflx_sun, &qflx_sha, qsatl, qaf, 1);
  if ((laisun > tol_lai || laisha > tol_lai) && (qflx_sun > 0.0 || qflx_sha > 0.0)) {
    iter = 0;
    oracle_phs_counters[0]++;
    for (;;) {
      iter = iter + 1;
      oracle_phs_counters[1]++;
      if (per_patch_newton) per_patch_newton[p - x->begp0]++;
      if (iter > itmax) oracle_phs_counters[2]++;
      spacF(x, p, c, xv, f, qflx_sun, qflx_sha);
      if (sqrt(f[1] * f[1] + f[2] * f[2] + f[3] * f[3] + f[4] * f[4]) < tolf * (qflx_sun + qflx_sha)) { flag = 0; break; }
      if (iter > itmax) { flag = 0; break; }
      spacA(x, p, c, xv, A, qflx_sun, qflx_sha, &flag);
      if (flag) break;
      if (laisun > tol_lai && laisha > tol_lai) {
        for (int i = 1; i <= 4; ++i) {      /* matmul(A,f): gfortran inlines as sum over k in order */
          double s = 0.0;
          for (int k = 1; k <= 4; ++k) s += A[i][k] * f[k];
          dx[i] = s;
        }
      } else {
        dx[sun] = 0.0;
        for (int i = sha; i <= root; ++i) {
          double s = 0.0;
          for (int k = sha; k <= root; ++k) s += A[i][k] * f[k];
          dx[i] = s;
        }
      }
      double mx = fmax(fmax(fabs(dx[1]), fabs(dx[2])), fmax(fabs(dx[3]), fabs(dx[4])));
      if (mx > 50000.0) for (int i = 1; i <= 4; ++i) dx[i] = 50000.0 * dx[i] / mx;
      if (laisun > tol_lai && laisha > tol_lai) {
        for (int i = 1; i <= 4; ++i) xv[i] = xv[i] + dx[i];
      } else if (laisha > tol_lai) {
        for (int i = 1; i <= 4; ++i) xv[i] = xv[i] + dx[i];
        xv[sun] = xv[xyl];
      } else {
        xv[xyl] = xv[xyl] + dx[xyl];
        xv[root] = xv[root] + dx[root];
        xv[sun] = xv[sun] + dx[sha];
        xv[sha] = xv[xyl];
      }
      if (sqrt(dx[1] * dx[1] + dx[2] * dx[2] + dx[3] * dx[3] + dx[4] * dx[4]) < toldx) break;
      if (xv[xyl] > xv[root]) xv[xyl] = xv[root];
      if (xv[sun] > xv[xyl]) xv[sun] = xv[xyl];
      if (xv[sha] > xv[xyl]) xv[sha] = xv[xyl];
    }
    oracle_phs_newton_hist[night][iter < 63 ? iter : 63]++;
  } else {
    flag = 1;
    oracle_phs_newton_hist[night][0]++;
  }
  if (flag) {
    getvegwp(x, p, c, xv, gb_mol, &gs0sun, &gs0sha, qsatl, qaf, &soilflux);
    *bsun = plc(x, xv[sun], p, sun);
    *bsha = plc(x, xv[sha], p, sha);
  } else {
    double qsun = qflx_sun * plc(x, xv[sun], p, sun);
    double qsha = qflx_sha * plc(x, xv[sha], p, sha);
    getqflx(x, p, c, gb_mol, &gs0sun, &gs0sha, &qsun, &qsha, qsatl, qaf, 0);
    if (qflx_sun > 0.0) *bsun = gs0sun / gs_mol_sun;
    else *bsun = plc(x, xv[sun], p, sun);
    if (qflx_sha > 0.0) *bsha = gs0sha / gs_mol_sha;
    else *bsha = plc(x, xv[sha], p, sha);
  }
  if (*bsun < 0.01) *bsun = 0.0;
  if (*bsha < 0.01) *bsha = 0.0;
  if (night) {
    gs0sun = *bsun * gs_mol_sun;
    gs0sha = *bsha * gs_mol_sha;
    getvegwp(x, p, c, xv, gb_mol, &gs0sun, &gs0sha, qsatl, qaf, &soilflux);
    if (soilflux < 0.0) soilflux = 0.0;
    P1(qflx_tran_veg, p) = soilflux;
  }
  const int g = P1(gridcell, p);
  if (night && G1(local_time_lt_noon, g)) {
    for (int i = 1; i <= 4; ++i) P2(vegwp_pd, p, i, 1) = xv[i];
  } else {
    for (int i = 1; i <= 4; ++i) P2(vegwp_pd, p, i, 1) = spval;
  }
}

typedef struct {   /* the per-call invariants hybrid/brent/ci_func pass around */
  int p, iv, c, g;
  double gb_mol, jesun, jesha, cair, oair, lmr_z_sun, lmr_z_sha, par_z_sun, par_z_sha, rh_can, qsatl, qaf;
} ci_args;

/* ci_func_PHS :4227-4486.  iv is always 1 (nlevcan = 1). */
static void ci_func_PHS(cf_ctx* x, const ci_args* a, double* xv, double cisun, double cisha, double* fvalsun,
                        double* fvalsha, double* bsun, double* bsha, int bflag, double gs0sun, double gs0sha,
                        double* gs_mol_sun, double* gs_mol_sha) {
  const int p = a->p, c = a->c, ivt = P1(itype, p);
  const double gb_mol = a->gb_mol, cair = a->cair, oair = a->oair, rh_can = a->rh_can;
  const double forc_pbot = C1(forc_pbot, c);
  const double cp = P1(cp, p), kc = P1(kc, p), ko = P1(ko, p);
  double ai, cs_sun = 0.0, cs_sha, aquad, bquad, cquad, r1, r2, term;
  double *ac_sun = &P2(ac_phs, p, sun, 1), *ac_sha = &P2(ac_phs, p, sha, 1);
  double *aj_sun = &P2(aj_phs, p, sun, 1), *aj_sha = &P2(aj_phs, p, sha, 1);
  double *ap_sun = &P2(ap_phs, p, sun, 1), *ap_sha = &P2(ap_phs, p, sha, 1);
  double *ag_sun = &P2(ag_phs, p, sun, 1), *ag_sha = &P2(ag_phs, p, sha, 1);
  double *an_sun = &P2(an_sun, p, 1, 1), *an_sha = &P2(an_sha, p, 1, 1);
  const double medint = PFT(pft_medlynintercept, ivt), medslope = PFT(pft_medlynslope, ivt);
  const double bbb = x->bbb[p - x->begp0], mbb = x->mbb[p - x->begp0];

  oracle_phs_counters[3]++;
  if (bflag) calcstress(x, p, c, xv, bsun, bsha, gb_mol, gs0sun, gs0sha, a->qsatl, a->qaf);

  if (P1(c3flag, p)) {
    *ac_sun = *bsun * P2(vcmax_z_phs, p, sun, 1) * fmax(cisun - cp, 0.0) / (cisun + kc * (1.0 + oair / ko));
    *ac_sha = *bsha * P2(vcmax_z_phs, p, sha, 1) * fmax(cisha - cp, 0.0) / (cisha + kc * (1.0 + oair / ko));
    *aj_sun = a->jesun * fmax(cisun - cp, 0.0) / (4.0 * cisun + 8.0 * cp);
    *aj_sha = a->jesha * fmax(cisha - cp, 0.0) / (4.0 * cisha + 8.0 * cp);
    *ap_sun = 3.0 * P2(tpu_z_phs, p, sun, 1);
    *ap_sha = 3.0 * P2(tpu_z_phs, p, sha, 1);
  } else {
    *ac_sun = *bsun * P2(vcmax_z_phs, p, sun, 1);
    *ac_sha = *bsha * P2(vcmax_z_phs, p, sha, 1);
    *aj_sun = P1(qe, p) * a->par_z_sun * 4.6;
    *aj_sha = P1(qe, p) * a->par_z_sha * 4.6;
    *ap_sun = P2(kp_z_phs, p, sun, 1) * fmax(cisun, 0.0) / forc_pbot;
    *ap_sha = P2(kp_z_phs, p, sha, 1) * fmax(cisha, 0.0) / forc_pbot;
  }
  aquad = PFT(pft_theta_cj, ivt);
  bquad = -(*ac_sun + *aj_sun);
  cquad = *ac_sun * *aj_sun;
  QUAD(aquad, bquad, cquad, &r1, &r2);
  ai = fmin(r1, r2);
  aquad = x->prm->theta_ip;
  bquad = -(ai + *ap_sun);
  cquad = ai * *ap_sun;
  QUAD(aquad, bquad, cquad, &r1, &r2);
  *ag_sun = fmax(0.0, fmin(r1, r2));
  aquad = PFT(pft_theta_cj, ivt);
  bquad = -(*ac_sha + *aj_sha);
  cquad = *ac_sha * *aj_sha;
  QUAD(aquad, bquad, cquad, &r1, &r2);
  ai = fmin(r1, r2);
  aquad = x->prm->theta_ip;
  bquad = -(ai + *ap_sha);
  cquad = ai * *ap_sha;
  QUAD(aquad, bquad, cquad, &r1, &r2);
  *ag_sha = fmax(0.0, fmin(r1, r2));

  *an_sun = *ag_sun - *bsun * a->lmr_z_sun;
  *an_sha = *ag_sha - *bsha * a->lmr_z_sha;

  const int medlyn = (x->prm->stomatalcond_mtd == 2);
  if (*an_sun < 0.0) {
    *gs_mol_sun = medlyn ? medint : bbb;
    *gs_mol_sun = fmax(*bsun * *gs_mol_sun, 1.0);
    *fvalsun = 0.0;
  }
  if (*an_sha < 0.0) {
    *gs_mol_sha = medlyn ? medint : bbb;
    *gs_mol_sha = fmax(*bsha * *gs_mol_sha, 1.0);
    *fvalsha = 0.0;
  }
  if ((*an_sun < 0.0) && (*an_sha < 0.0)) return;

  if (*an_sun >= 0.0) {
    cs_sun = cair - 1.4 / gb_mol * *an_sun * forc_pbot;
    cs_sun = fmax(cs_sun, max_cs);
  }
  if (medlyn) {
    if (*an_sun >= 0.0) {
      term = 1.6 * *an_sun / (cs_sun / forc_pbot * 1.e06);
      aquad = 1.0;
      bquad = -(2.0 * (medint * 1.e-06 + term) + ((medslope * term) * (medslope * term)) / (gb_mol * 1.e-06 * rh_can));
      cquad = medint * medint * 1.e-12 + (2.0 * medint * 1.e-06 + term * (1.0 - medslope * medslope / rh_can)) * term;
      QUAD(aquad, bquad, cquad, &r1, &r2);
      *gs_mol_sun = fmax(r1, r2) * 1.e06;
    }
    if (*an_sha >= 0.0) {
      cs_sha = cair - 1.4 / gb_mol * *an_sha * forc_pbot;
      cs_sha = fmax(cs_sha, max_cs);
      term = 1.6 * *an_sha / (cs_sha / forc_pbot * 1.e06);
      aquad = 1.0;
      bquad = -(2.0 * (medint * 1.e-06 + term) + ((medslope * term) * (medslope * term)) / (gb_mol * 1.e-06 * rh_can));
      cquad = medint * medint * 1.e-12 + (2.0 * medint * 1.e-06 + term * (1.0 - medslope * medslope / rh_can)) * term;
      QUAD(aquad, bquad, cquad, &r1, &r2);
      *gs_mol_sha = fmax(r1, r2) * 1.e06;
    }
  } else {
    if (*an_sun >= 0.0) {
      aquad = cs_sun;
      bquad = cs_sun * (gb_mol - fmax(*bsun * bbb, 1.0)) - mbb * *an_sun * forc_pbot;
      cquad = -gb_mol * (cs_sun * fmax(*bsun * bbb, 1.0) + mbb * *an_sun * forc_pbot * rh_can);
      QUAD(aquad, bquad, cquad, &r1, &r2);
      *gs_mol_sun = fmax(r1, r2);
    }
    if (*an_sha >= 0.0) {
      cs_sha = cair - 1.4 / gb_mol * *an_sha * forc_pbot;
      cs_sha = fmax(cs_sha, max_cs);
      aquad = cs_sha;
      bquad = cs_sha * (gb_mol - fmax(*bsha * bbb, 1.0)) - mbb * *an_sha * forc_pbot;
      cquad = -gb_mol * (cs_sha * fmax(*bsha * bbb, 1.0) + mbb * *an_sha * forc_pbot * rh_can);
      QUAD(aquad, bquad, cquad, &r1, &r2);
      *gs_mol_sha = fmax(r1, r2);
    }
  }
  if (*an_sun >= 0.0) {
    if (*gs_mol_sun > 0.0)
      *fvalsun = cisun - cair + *an_sun * forc_pbot * (1.4 * *gs_mol_sun + 1.6 * gb_mol) / (gb_mol * *gs_mol_sun);
    else
      *fvalsun = cisun - cair;
  }
  if (*an_sha >= 0.0) {
    if (*gs_mol_sha > 0.0)
      *fvalsha = cisha - cair + *an_sha * forc_pbot * (1.4 * *gs_mol_sha + 1.6 * gb_mol) / (gb_mol * *gs_mol_sha);
    else
      *fvalsha = cisha - cair;
  }
}

/* brent_PHS :4068-4223 */
static void brent_PHS(cf_ctx* x, const ci_args* ca, double* xsun, double x1sun, double x2sun, double f1sun, double f2sun,
                      double* xsha, double x1sha, double x2sha, double f1sha, double f2sha, double tol,
                      double* gs_mol_sun, double* gs_mol_sha, double* bsun, double* bsha) {
  const int itmax = 20;
  const double eps = 1.e-4;
  double a[3], b[3], c[3], d[3] = {0, 0, 0}, e[3] = {0, 0, 0}, fa[3], fb[3], fc[3], pp[3], q[3], r[3], s[3], tol1[3], xm[3];
  double xdummy[5] = {0, 0, 0, 0, 0};
  int iter;
  a[1] = x1sun; a[2] = x1sha;
  b[1] = x2sun; b[2] = x2sha;
  fa[1] = f1sun; fa[2] = f1sha;
  fb[1] = f2sun; fb[2] = f2sha;
  for (int ph = 1; ph <= 2; ++ph)
    if ((fa[ph] > 0.0 && fb[ph] > 0.0) || (fa[ph] < 0.0 && fb[ph] < 0.0)) fail(x, CTSM_ERR_BRENT, ca->p);
  for (int ph = 1; ph <= 2; ++ph) { c[ph] = b[ph]; fc[ph] = fb[ph]; }
  oracle_phs_counters[4]++;
  iter = 0;
  for (;;) {
    if (iter == itmax) break;
    iter = iter + 1;
    for (int ph = 1; ph <= 2; ++ph) {
      if ((fb[ph] > 0.0 && fc[ph] > 0.0) || (fb[ph] < 0.0 && fc[ph] < 0.0)) {
        c[ph] = a[ph]; fc[ph] = fa[ph]; d[ph] = b[ph] - a[ph]; e[ph] = d[ph];
      }
      if (fabs(fc[ph]) < fabs(fb[ph])) {
        a[ph] = b[ph]; b[ph] = c[ph]; c[ph] = a[ph];
        fa[ph] = fb[ph]; fb[ph] = fc[ph]; fc[ph] = fa[ph];
      }
    }
    for (int ph = 1; ph <= 2; ++ph) {
      tol1[ph] = 2.0 * eps * fabs(b[ph]) + 0.5 * tol;
      xm[ph] = 0.5 * (c[ph] - b[ph]);
    }
    if (fabs(xm[sun]) <= tol1[sun] || fb[sun] == 0.0) {
      if (fabs(xm[sha]) <= tol1[sha] || fb[sha] == 0.0) {
        *xsun = b[sun];
        *xsha = b[sha];
        return;
      }
    }
    for (int ph = 1; ph <= 2; ++ph) {
      if (fabs(e[ph]) >= tol1[ph] && fabs(fa[ph]) > fabs(fb[ph])) {
        s[ph] = fb[ph] / fa[ph];
        if (a[ph] == c[ph]) {
          pp[ph] = 2.0 * xm[ph] * s[ph];
          q[ph] = 1.0 - s[ph];
        } else {
          q[ph] = fa[ph] / fc[ph];
          r[ph] = fb[ph] / fc[ph];
          pp[ph] = s[ph] * (2.0 * xm[ph] * q[ph] * (q[ph] - r[ph]) - (b[ph] - a[ph]) * (r[ph] - 1.0));
          q[ph] = (q[ph] - 1.0) * (r[ph] - 1.0) * (s[ph] - 1.0);
        }
        if (pp[ph] > 0.0) q[ph] = -q[ph];
        pp[ph] = fabs(pp[ph]);
        if (2.0 * pp[ph] < fmin(3.0 * xm[ph] * q[ph] - fabs(tol1[ph] * q[ph]), fabs(e[ph] * q[ph]))) {
          e[ph] = d[ph];
          d[ph] = pp[ph] / q[ph];
        } else {
          d[ph] = xm[ph];
          e[ph] = d[ph];
        }
      } else {
        d[ph] = xm[ph];
        e[ph] = d[ph];
      }
      a[ph] = b[ph];
      fa[ph] = fb[ph];
      if (fabs(d[ph]) > tol1[ph]) b[ph] = b[ph] + d[ph];
      else b[ph] = b[ph] + copysign(tol1[ph], xm[ph]);
    }
    const double gs0sun = *gs_mol_sun, gs0sha = *gs_mol_sha;
    ci_func_PHS(x, ca, xdummy, b[sun], b[sha], &fb[sun], &fb[sha], bsun, bsha, 0, gs0sun, gs0sha, gs_mol_sun, gs_mol_sha);
    if ((fb[sun] == 0.0) && (fb[sha] == 0.0)) break;
  }
  *xsun = b[sun];
  *xsha = b[sha];
}

/* hybrid_PHS :3815-4064 */
static void hybrid_PHS(cf_ctx* x, const ci_args* ca, double* x0sun, double* x0sha, double* bsun, double* bsha,
                       double* gs_mol_sun, double* gs_mol_sha, int* iter1, int* iter2) {
  const int p = ca->p, c = ca->c, g = ca->g;
  const double toldb = 1.e-2, eps = 1.e-2, eps1 = 1.e-4;
  const int itmax = 3;
  double xv[5] = {0, 0, 0, 0, 0}, gs0sun, gs0sha, soilflux, x1sun, f0sun = 0.0, f1sun = 0.0, xsun, dxsun, x1sha,
         f0sha = 0.0, f1sha = 0.0, xsha, dxsha, b0sun, b0sha, dbsun, dbsha, tolsun, tolsha, minf = 0.0, minxsun = 0.0,
         minxsha = 0.0;
  int bflag;
  x1sun = *x0sun;
  x1sha = *x0sha;
  bflag = 0;
  b0sun = -1.0;
  b0sha = -1.0;
  gs0sun = 0.0;
  gs0sha = 0.0;
  *bsun = 1.0;
  *bsha = 1.0;
  *iter1 = 0;
  for (;;) {
    for (int i = 1; i <= 4; ++i) xv[i] = P2(vegwp, p, i, 1);
    *iter1 = *iter1 + 1;
    oracle_phs_counters[5]++;
    *iter2 = 0;
    *x0sun = fmax(0.1, x1sun);
    x1sun = 0.99 * x1sun;
    *x0sha = fmax(0.1, x1sha);
    x1sha = 0.99 * x1sha;
    tolsun = fabs(x1sun) * eps;
    tolsha = fabs(x1sha) * eps;
    ci_func_PHS(x, ca, xv, *x0sun, *x0sha, &f0sun, &f0sha, bsun, bsha, bflag, gs0sun, gs0sha, gs_mol_sun, gs_mol_sha);
    dbsun = b0sun - *bsun;
    dbsha = b0sha - *bsha;
    b0sun = *bsun;
    b0sha = *bsha;
    bflag = 0;
    ci_func_PHS(x, ca, xv, x1sun, x1sha, &f1sun, &f1sha, bsun, bsha, bflag, gs0sun, gs0sha, gs_mol_sun, gs_mol_sha);
    for (;;) {
      if ((fabs(f0sun) < eps1) && (fabs(f0sha) < eps1)) {
        x1sun = *x0sun;
        x1sha = *x0sha;
        break;
      }
      if ((fabs(f1sun) < eps1) && (fabs(f1sha) < eps1)) break;
      *iter2 = *iter2 + 1;
      if ((f1sun - f0sun) == 0.0) dxsun = 0.5 * (x1sun + *x0sun) - x1sun;
      else dxsun = -f1sun * (x1sun - *x0sun) / (f1sun - f0sun);
      if ((f1sha - f0sha) == 0.0) dxsha = 0.5 * (x1sha + *x0sha) - x1sha;
      else dxsha = -f1sha * (x1sha - *x0sha) / (f1sha - f0sha);
      *x0sun = x1sun;
      x1sun = x1sun + dxsun;
      *x0sha = x1sha;
      x1sha = x1sha + dxsha;
      ci_func_PHS(x, ca, xv, x1sun, x1sha, &f1sun, &f1sha, bsun, bsha, bflag, gs0sun, gs0sha, gs_mol_sun, gs_mol_sha);
      if ((fabs(dxsun) < tolsun) && (fabs(dxsha) < tolsha)) {
        *x0sun = x1sun;
        *x0sha = x1sha;
        break;
      }
      if (*iter2 == 1) {
        minf = fabs(f1sun + f1sha);
        minxsun = x1sun;
        minxsha = x1sha;
      } else {
        if (fabs(f1sun + f1sha) < minf) {
          minf = fabs(f1sun + f1sha);
          minxsun = x1sun;
          minxsha = x1sha;
        }
      }
      if ((fabs(f1sun) < eps1) && (fabs(f1sha) < eps1)) break;
      if ((f1sun * f0sun < 0.0) && (f1sha * f0sha < 0.0)) {
        brent_PHS(x, ca, &xsun, *x0sun, x1sun, f0sun, f1sun, &xsha, *x0sha, x1sha, f0sha, f1sha, tolsun, gs_mol_sun,
                  gs_mol_sha, bsun, bsha);
        *x0sun = xsun;
        *x0sha = xsha;
        break;
      }
      if (*iter2 > itmax) {
        x1sun = minxsun;
        x1sha = minxsha;
        ci_func_PHS(x, ca, xv, x1sun, x1sha, &f1sun, &f1sha, bsun, bsha, bflag, gs0sun, gs0sha, gs_mol_sun, gs_mol_sha);
        break;
      }
    }
    if (*bsun > 0.01) gs0sun = *gs_mol_sun / *bsun;
    if (*bsha > 0.01) gs0sha = *gs_mol_sha / *bsha;
    bflag = 1;
    if ((fabs(dbsun) < toldb) && (fabs(dbsha) < toldb)) break;
    if (*iter1 > itmax) break;
  }
  *x0sun = x1sun;
  *x0sha = x1sha;
  getvegwp(x, p, c, xv, ca->gb_mol, gs_mol_sun, gs_mol_sha, ca->qsatl, ca->qaf, &soilflux);
  for (int i = 1; i <= 4; ++i) P2(vegwp, p, i, 1) = xv[i];
  if (G1(near_local_noon, g)) {
    for (int i = 1; i <= 4; ++i) P2(vegwp_ln, p, i, 1) = P2(vegwp, p, i, 1);
  } else {
    for (int i = 1; i <= 4; ++i) P2(vegwp_ln, p, i, 1) = spval;
  }
  if (soilflux < 0.0) soilflux = 0.0;
  P1(qflx_tran_veg, p) = soilflux;
}

/* statement functions :2918-2920 */
static double ft(double tl, double ha) { return exp(ha / (rgas * 1.e-3 * (tfrz + 25.0)) * (1.0 - (tfrz + 25.0) / tl)); }
static double fth(double tl, double hd, double se, double scaleFactor) {
  return scaleFactor / (1.0 + exp((-hd + se * tl) / (rgas * 1.e-3 * tl)));
}
static double fth25(double hd, double se) { return 1.0 + exp((-hd + se * (tfrz + 25.0)) / (rgas * 1.e-3 * (tfrz + 25.0))); }

/* TimeStepInit :1143-1201 */
void oracle_photosyns_timestepinit(cf_ctx* x, const ctsm_bounds_t* bounds) {
  for (int p = bounds->begp; p <= bounds->endp; ++p) {
    if (!P1(patch_lakpoi, p)) {
      P1(psnsun, p) = 0.0; P1(psnsun_wc, p) = 0.0; P1(psnsun_wj, p) = 0.0; P1(psnsun_wp, p) = 0.0;
      P1(psnsha, p) = 0.0; P1(psnsha_wc, p) = 0.0; P1(psnsha_wj, p) = 0.0; P1(psnsha_wp, p) = 0.0;
      P1(fpsn, p) = 0.0; P1(fpsn_wc, p) = 0.0; P1(fpsn_wj, p) = 0.0; P1(fpsn_wp, p) = 0.0;
    }
  }
}

/* PhotosynthesisTotal :2065-2151 */
void oracle_photosynthesis_total(cf_ctx* x, int fn, const int32_t* filterp) {
  for (int f = 0; f < fn; ++f) {
    const int p = filterp[f];
    P1(fpsn, p) = P1(psnsun, p) * P1(laisun, p) + P1(psnsha, p) * P1(laisha, p);
    P1(fpsn_wc, p) = P1(psnsun_wc, p) * P1(laisun, p) + P1(psnsha_wc, p) * P1(laisha, p);
    P1(fpsn_wj, p) = P1(psnsun_wj, p) * P1(laisun, p) + P1(psnsha_wj, p) * P1(laisha, p);
    P1(fpsn_wp, p) = P1(psnsun_wp, p) * P1(laisun, p) + P1(psnsha_wp, p) * P1(laisha, p);
  }
}

/* PhotosynthesisHydraulicStress :2704-3811.  Patch-indexed dummies are (begp:endp) arrays offset by begp0. */
void oracle_photosynthesis_hydraulic_stress(cf_ctx* x, int fn, const int32_t* filterp, const double* esat_tv,
                                            const double* eair, const double* oair, const double* cair,
                                            const double* rb, double* bsun, double* bsha, double* btran,
                                            const double* dayl_factor, const double* qsatl, const double* qaf) {
  const ctsm_params_t* pr = x->prm;
  const int medlyn = (pr->stomatalcond_mtd == 2);
  const double croot_lateral_length = 0.25, c_to_b = 2.0;
  const double rsmax0 = 2.e4;
  const int b0 = x->begp0;
  const int np = x->np;
  double* jmax_z = (double*)calloc((size_t)2 * np, sizeof(double));        /* (p, sun/sha) */
  double* psn_wc_z_sun = (double*)calloc(np, sizeof(double));
  double* psn_wj_z_sun = (double*)calloc(np, sizeof(double));
  double* psn_wp_z_sun = (double*)calloc(np, sizeof(double));
  double* psn_wc_z_sha = (double*)calloc(np, sizeof(double));
  double* psn_wj_z_sha = (double*)calloc(np, sizeof(double));
  double* psn_wp_z_sha = (double*)calloc(np, sizeof(double));
#define JMAX(p, s) jmax_z[(size_t)((s) - 1) * np + ((p) - b0)]
#define L(arr, p) arr[(p) - b0]

  const double lmrc = fth25(pr->lmrhd, pr->lmrse);

  for (int f = 0; f < fn; ++f) {        /* :3063-3114 root-soil interface conductance */
    const int p = filterp[f], c = P1(column, p), ivt = P1(itype, p);
    double rai[NLEVSOI + 1], fs[NLEVSOI + 1];
    for (int j = 1; j <= NLEVSOI; ++j) {
      double root_biomass_density = c_to_b * P1(froot_carbon, p) * P2(rootfr, p, j, 1) / C2(dz, c, j, SNOSOI_LO);
      root_biomass_density = fmax(c_to_b * 1.0, root_biomass_density);
      const double root_cross_sec_area = rpi * (PFT(pft_root_radius, ivt) * PFT(pft_root_radius, ivt));
      const double root_length_density = root_biomass_density / (PFT(pft_root_density, ivt) * root_cross_sec_area);
      rai[j] = (P1(tsai, p) + P1(tlai, p)) * PFT(pft_froot_leaf, ivt) * P2(rootfr, p, j, 1);
      const double croot_average_length = croot_lateral_length;
      const double r_soil = sqrt(1. / (rpi * root_length_density));
      double soil_conductance = fmin(C2(hksat, c, j, 1), C2(hk_l, c, j, 1)) / (1.e3 * r_soil);
      fs[j] = plc(x, C2(smp_l, c, j, 1), p, root);
      double root_conductance = (fs[j] * rai[j] * PFT(pft_krmax, ivt)) / (croot_average_length + C2(z, c, j, SNOSOI_LO));
      soil_conductance = fmax(soil_conductance, 1.e-16);
      root_conductance = fmax(root_conductance, 1.e-16);
      P2(root_conductance, p, j, 1) = root_conductance;
      P2(soil_conductance, p, j, 1) = soil_conductance;
      const double rs_resis = 1.0 / soil_conductance + 1.0 / root_conductance;
      if (rai[j] * P2(rootfr, p, j, 1) > 0.0 && j > 1) P2(k_soil_root, p, j, 1) = 1.0 / rs_resis;
      else P2(k_soil_root, p, j, 1) = 0.0;
    }
  }

  for (int f = 0; f < fn; ++f) {        /* :3118-3164 */
    const int p = filterp[f], c = P1(column, p), ivt = P1(itype, p);
    if ((int)nearbyint(PFT(pft_c3psn, ivt)) == 1) P1(c3flag, p) = 1;
    else if ((int)nearbyint(PFT(pft_c3psn, ivt)) == 0) P1(c3flag, p) = 0;
    double bbbopt = 0.0;
    if (P1(c3flag, p)) { P1(qe, p) = 0.0; if (!medlyn) bbbopt = bbbopt_c3; }
    else { P1(qe, p) = 0.05; if (!medlyn) bbbopt = bbbopt_c4; }
    if (!medlyn) { L(x->bbb, p) = bbbopt; L(x->mbb, p) = PFT(pft_mbbopt, ivt); }
    const double kc25 = pr->kc25_coef * C1(forc_pbot, c);
    const double ko25 = pr->ko25_coef * C1(forc_pbot, c);
    const double sco = 0.5 * 0.209 / pr->cp25_yr2000;
    const double cp25 = 0.5 * L(oair, p) / sco;
    P1(kc, p) = kc25 * ft(P1(t_veg, p), pr->kcha);
    P1(ko, p) = ko25 * ft(P1(t_veg, p), pr->koha);
    P1(cp, p) = cp25 * ft(P1(t_veg, p), pr->cpha);
  }

  for (int f = 0; f < fn; ++f) {        /* :3170-3469 */
    const int p = filterp[f], ivt = P1(itype, p);
    const double t_veg = P1(t_veg, p), t10 = P1(t_a10, p);
    const double leafcn_local = PFT(pft_leafcn, ivt);
    P1(lnca, p) = 1.0 / (PFT(pft_slatop, ivt) * leafcn_local);
    P1(lnca, p) = fmin(P1(lnca, p), 10.0);
    double vcmax25top = P1(lnca, p) * PFT(pft_flnr, ivt) * pr->fnr * pr->act25 * L(dayl_factor, p);
    vcmax25top = vcmax25top * PFT(pft_fnitr, ivt);
    const double jmax25top = ((2.59 - 0.035 * fmin(fmax((t10 - tfrz), 11.0), 35.0)) * vcmax25top) * pr->jmax25top_sf;
    const double tpu25top = pr->tpu25ratio * vcmax25top;
    const double kp25top = pr->kp25ratio * vcmax25top;
    P1(luvcmax25top, p) = vcmax25top;
    P1(lujmax25top, p) = jmax25top;
    P1(lutpu25top, p) = tpu25top;
    /* kn(p) (:3276-3280) is computed by the reference but unused when nlevcan == 1 */
    double lmr25top;
    if (P1(c3flag, p)) lmr25top = vcmax25top * pr->leaf_mr_vcm;
    else lmr25top = vcmax25top * 0.025;

    for (int iv = 1; iv <= P1(nrad, p); ++iv) {
      const double nscaler_sun = P1(vcmaxcintsun, p), nscaler_sha = P1(vcmaxcintsha, p);
      double lmr25_sun = lmr25top * nscaler_sun;
      double lmr25_sha = lmr25top * nscaler_sha;
      const int luna_patch = pr->use_luna && P1(c3flag, p) && PFT(pft_crop, ivt) == 0.0;
      if (luna_patch) {
        lmr25_sun = pr->leaf_mr_vcm * P2(vcmx25_z, p, iv, 1);
        lmr25_sha = pr->leaf_mr_vcm * P2(vcmx25_z, p, iv, 1);
      }
      double lmr_sun, lmr_sha;
      if (P1(c3flag, p)) {
        lmr_sun = lmr25_sun * ft(t_veg, pr->lmrha) * fth(t_veg, pr->lmrhd, pr->lmrse, lmrc);
        lmr_sha = lmr25_sha * ft(t_veg, pr->lmrha) * fth(t_veg, pr->lmrhd, pr->lmrse, lmrc);
      } else {
        lmr_sun = lmr25_sun * pow(2.0, (t_veg - (tfrz + 25.0)) / 10.0);
        lmr_sun = lmr_sun / (1.0 + exp(1.3 * (t_veg - (tfrz + 55.0))));
        lmr_sha = lmr25_sha * pow(2.0, (t_veg - (tfrz + 25.0)) / 10.0);
        lmr_sha = lmr_sha / (1.0 + exp(1.3 * (t_veg - (tfrz + 55.0))));
      }
      lmr_sun = lmr_sun * fmin((0.2 * exp(3.218 * P2(tlai_z, p, iv, 1))), 1.0);
      lmr_sha = lmr_sha * fmin((0.2 * exp(3.218 * P2(tlai_z, p, iv, 1))), 1.0);

      if (P2(parsun_z, p, iv, 1) <= 0.0) {
        P2(vcmax_z_phs, p, sun, 1) = 0.0; JMAX(p, sun) = 0.0; P2(tpu_z_phs, p, sun, 1) = 0.0; P2(kp_z_phs, p, sun, 1) = 0.0;
        P2(vcmax_z_phs, p, sha, 1) = 0.0; JMAX(p, sha) = 0.0; P2(tpu_z_phs, p, sha, 1) = 0.0; P2(kp_z_phs, p, sha, 1) = 0.0;
      } else {
        double vcmax25_sun, vcmax25_sha, jmax25_sun, jmax25_sha, tpu25_sun, tpu25_sha;
        if (luna_patch) {
          vcmax25_sun = P2(vcmx25_z, p, iv, 1);
          vcmax25_sha = P2(vcmx25_z, p, iv, 1);
          jmax25_sun = P2(jmx25_z, p, iv, 1);
          jmax25_sha = P2(jmx25_z, p, iv, 1);
          tpu25_sun = pr->tpu25ratio * vcmax25_sun;
          tpu25_sha = pr->tpu25ratio * vcmax25_sha;
          if (P1(vcmaxcintsun, p) > 0.0) {
            vcmax25_sha = vcmax25_sun * P1(vcmaxcintsha, p) / P1(vcmaxcintsun, p);
            jmax25_sha = jmax25_sun * P1(vcmaxcintsha, p) / P1(vcmaxcintsun, p);
            tpu25_sha = tpu25_sun * P1(vcmaxcintsha, p) / P1(vcmaxcintsun, p);
          }
        } else {
          vcmax25_sun = vcmax25top * nscaler_sun;
          jmax25_sun = jmax25top * nscaler_sun;
          tpu25_sun = tpu25top * nscaler_sun;
          vcmax25_sha = vcmax25top * nscaler_sha;
          jmax25_sha = jmax25top * nscaler_sha;
          tpu25_sha = tpu25top * nscaler_sha;
        }
        const double kp25_sun = kp25top * nscaler_sun;
        const double kp25_sha = kp25top * nscaler_sha;
        const double vcmaxse = (668.39 - 1.07 * fmin(fmax((t10 - tfrz), 11.0), 35.0)) * pr->vcmaxse_sf;
        const double jmaxse = (659.70 - 0.75 * fmin(fmax((t10 - tfrz), 11.0), 35.0)) * pr->jmaxse_sf;
        const double tpuse = (668.39 - 1.07 * fmin(fmax((t10 - tfrz), 11.0), 35.0)) * pr->tpuse_sf;
        const double vcmaxc = fth25(pr->vcmaxhd, vcmaxse);
        const double jmaxc = fth25(pr->jmaxhd, jmaxse);
        const double tpuc = fth25(pr->tpuhd, tpuse);
        P2(vcmax_z_phs, p, sun, 1) = vcmax25_sun * ft(t_veg, pr->vcmaxha) * fth(t_veg, pr->vcmaxhd, vcmaxse, vcmaxc);
        JMAX(p, sun) = jmax25_sun * ft(t_veg, pr->jmaxha) * fth(t_veg, pr->jmaxhd, jmaxse, jmaxc);
        P2(tpu_z_phs, p, sun, 1) = tpu25_sun * ft(t_veg, pr->tpuha) * fth(t_veg, pr->tpuhd, tpuse, tpuc);
        P2(vcmax_z_phs, p, sha, 1) = vcmax25_sha * ft(t_veg, pr->vcmaxha) * fth(t_veg, pr->vcmaxhd, vcmaxse, vcmaxc);
        JMAX(p, sha) = jmax25_sha * ft(t_veg, pr->jmaxha) * fth(t_veg, pr->jmaxhd, jmaxse, jmaxc);
        P2(tpu_z_phs, p, sha, 1) = tpu25_sha * ft(t_veg, pr->tpuha) * fth(t_veg, pr->tpuhd, tpuse, tpuc);
        if (!P1(c3flag, p)) {
          double v = vcmax25_sun * pow(2.0, (t_veg - (tfrz + 25.0)) / 10.0);
          v = v / (1.0 + exp(0.2 * ((tfrz + 15.0) - t_veg)));
          v = v / (1.0 + exp(0.3 * (t_veg - (tfrz + 40.0))));
          P2(vcmax_z_phs, p, sun, 1) = v;
          v = vcmax25_sha * pow(2.0, (t_veg - (tfrz + 25.0)) / 10.0);
          v = v / (1.0 + exp(0.2 * ((tfrz + 15.0) - t_veg)));
          v = v / (1.0 + exp(0.3 * (t_veg - (tfrz + 40.0))));
          P2(vcmax_z_phs, p, sha, 1) = v;
        }
        P2(kp_z_phs, p, sun, 1) = kp25_sun * pow(2.0, (t_veg - (tfrz + 25.0)) / 10.0);
        P2(kp_z_phs, p, sha, 1) = kp25_sha * pow(2.0, (t_veg - (tfrz + 25.0)) / 10.0);
      }
      if (pr->light_inhibit && P2(parsun_z, p, 1, 1) > 0.0) lmr_sun = lmr_sun * 0.67;
      if (pr->light_inhibit && P2(parsha_z, p, 1, 1) > 0.0) lmr_sha = lmr_sha * 0.67;
      P2(lmrsun_z, p, iv, 1) = lmr_sun;
      P2(lmrsha_z, p, iv, 1) = lmr_sha;
    }
  }

  for (int f = 0; f < fn; ++f) {        /* :3477-3714 leaf-level photosynthesis and stomatal conductance */
    const int p = filterp[f], c = P1(column, p), g = P1(gridcell, p), ivt = P1(itype, p);
    const double forc_pbot = C1(forc_pbot, c);
    const double cf = forc_pbot / (rgas * 1.e-3 * P1(thm, p)) * 1.e06;
    const double gb = 1.0 / L(rb, p);
    P1(gb_mol, p) = gb * cf;
    const double gb_mol = P1(gb_mol, p);
    const int notcrop_or_nomod = (PFT(pft_crop, ivt) == 0.0 || !pr->modifyphoto_and_lmr_forcrop);
    for (int iv = 1; iv <= P1(nrad, p); ++iv) {
      double gsminsun, gsminsha, gs_slope_sun, gs_slope_sha;
      if (P2(parsun_z, p, iv, 1) <= 0.0) {        /* night */
        P2(vegwp, p, sun, 1) = 1.0;
        if (!medlyn) { gsminsun = L(x->bbb, p); gsminsha = L(x->bbb, p); }
        else { gsminsun = PFT(pft_medlynintercept, ivt); gsminsha = PFT(pft_medlynintercept, ivt); }
        double xv[5] = {0, P2(vegwp, p, 1, 1), P2(vegwp, p, 2, 1), P2(vegwp, p, 3, 1), P2(vegwp, p, 4, 1)};
        calcstress(x, p, c, xv, &L(bsun, p), &L(bsha, p), gb_mol, gsminsun, gsminsha, L(qsatl, p), L(qaf, p));
        for (int i = 1; i <= 4; ++i) P2(vegwp, p, i, 1) = xv[i];
        P2(ac_phs, p, sun, 1) = 0.0; P2(aj_phs, p, sun, 1) = 0.0; P2(ap_phs, p, sun, 1) = 0.0; P2(ag_phs, p, sun, 1) = 0.0;
        if (notcrop_or_nomod) P2(an_sun, p, iv, 1) = P2(ag_phs, p, sun, 1) - L(bsun, p) * P2(lmrsun_z, p, iv, 1);
        else P2(an_sun, p, iv, 1) = P2(ag_phs, p, sun, 1) - P2(lmrsun_z, p, iv, 1);
        P2(psnsun_z, p, iv, 1) = 0.0;
        L(psn_wc_z_sun, p) = 0.0; L(psn_wj_z_sun, p) = 0.0; L(psn_wp_z_sun, p) = 0.0;
        P2(rssun_z, p, iv, 1) = fmin(rsmax0, 1.0 / (fmax(L(bsun, p) * gsminsun, 1.0)) * cf);
        P2(cisun_z, p, iv, 1) = 0.0;
        P2(ac_phs, p, sha, 1) = 0.0; P2(aj_phs, p, sha, 1) = 0.0; P2(ap_phs, p, sha, 1) = 0.0; P2(ag_phs, p, sha, 1) = 0.0;
        if (notcrop_or_nomod) P2(an_sha, p, iv, 1) = P2(ag_phs, p, sha, 1) - L(bsha, p) * P2(lmrsha_z, p, iv, 1);
        else P2(an_sha, p, iv, 1) = P2(ag_phs, p, sha, 1) - P2(lmrsha_z, p, iv, 1);
        P2(psnsha_z, p, iv, 1) = 0.0;
        L(psn_wc_z_sha, p) = 0.0; L(psn_wj_z_sha, p) = 0.0; L(psn_wp_z_sha, p) = 0.0;
        P2(rssha_z, p, iv, 1) = fmin(rsmax0, 1.0 / (fmax(L(bsha, p) * gsminsha, 1.0)) * cf);
        P2(cisha_z, p, iv, 1) = 0.0;
        P2(gs_mol_sun, p, iv, 1) = cf / P2(rssun_z, p, iv, 1);
        P2(gs_mol_sha, p, iv, 1) = cf / P2(rssha_z, p, iv, 1);
      } else {                                    /* day */
        const double ceair = fmin(L(eair, p), L(esat_tv, p));
        double rh_can;
        if (!medlyn) rh_can = ceair / L(esat_tv, p);
        else {
          rh_can = fmax((L(esat_tv, p) - ceair), medlyn_rh_can_max) * medlyn_rh_can_fact;
          P1(vpd_can, p) = rh_can;
        }
        double qabs, aquad, bquad, cquad, r1, r2;
        qabs = 0.5 * (1.0 - pr->fnps) * P2(parsun_z, p, iv, 1) * 4.6;
        aquad = pr->theta_psii;
        bquad = -(qabs + JMAX(p, sun));
        cquad = qabs * JMAX(p, sun);
        QUAD(aquad, bquad, cquad, &r1, &r2);
        const double je_sun = fmin(r1, r2);
        qabs = 0.5 * (1.0 - pr->fnps) * P2(parsha_z, p, iv, 1) * 4.6;
        aquad = pr->theta_psii;
        bquad = -(qabs + JMAX(p, sha));
        cquad = qabs * JMAX(p, sha);
        QUAD(aquad, bquad, cquad, &r1, &r2);
        const double je_sha = fmin(r1, r2);
        if (P1(c3flag, p)) { P2(cisun_z, p, iv, 1) = 0.7 * L(cair, p); P2(cisha_z, p, iv, 1) = 0.7 * L(cair, p); }
        else { P2(cisun_z, p, iv, 1) = 0.4 * L(cair, p); P2(cisha_z, p, iv, 1) = 0.4 * L(cair, p); }
        ci_args ca = {p, iv, c, g, gb_mol, je_sun, je_sha, L(cair, p), L(oair, p), P2(lmrsun_z, p, iv, 1),
                      P2(lmrsha_z, p, iv, 1), P2(parsun_z, p, iv, 1), P2(parsha_z, p, iv, 1), rh_can, L(qsatl, p), L(qaf, p)};
        int iter1, iter2;
        hybrid_PHS(x, &ca, &P2(cisun_z, p, iv, 1), &P2(cisha_z, p, iv, 1), &L(bsun, p), &L(bsha, p),
                   &P2(gs_mol_sun, p, iv, 1), &P2(gs_mol_sha, p, iv, 1), &iter1, &iter2);
        if (medlyn) {
          gsminsun = PFT(pft_medlynintercept, ivt); gsminsha = gsminsun;
          gs_slope_sun = PFT(pft_medlynslope, ivt); gs_slope_sha = gs_slope_sun;
        } else {
          gsminsun = L(x->bbb, p); gsminsha = gsminsun;
          gs_slope_sun = L(x->mbb, p); gs_slope_sha = gs_slope_sun;
        }
        (void)gs_slope_sun; (void)gs_slope_sha;
        if (P2(an_sun, p, iv, 1) < 0.0) P2(gs_mol_sun, p, iv, 1) = fmax(L(bsun, p) * gsminsun, 1.0);
        if (P2(an_sha, p, iv, 1) < 0.0) P2(gs_mol_sha, p, iv, 1) = fmax(L(bsha, p) * gsminsha, 1.0);
        if (G1(near_local_noon, g)) {
          P2(gs_mol_sun_ln, p, iv, 1) = P2(gs_mol_sun, p, iv, 1);
          P2(gs_mol_sha_ln, p, iv, 1) = P2(gs_mol_sha, p, iv, 1);
        } else {
          P2(gs_mol_sun_ln, p, iv, 1) = spval;
          P2(gs_mol_sha_ln, p, iv, 1) = spval;
        }
        P2(cisun_z, p, iv, 1) = L(cair, p) - P2(an_sun, p, iv, 1) * forc_pbot *
            (1.4 * P2(gs_mol_sun, p, iv, 1) + 1.6 * gb_mol) / (gb_mol * P2(gs_mol_sun, p, iv, 1));
        P2(cisun_z, p, iv, 1) = fmax(P2(cisun_z, p, iv, 1), 1.e-06);
        P2(cisha_z, p, iv, 1) = L(cair, p) - P2(an_sha, p, iv, 1) * forc_pbot *
            (1.4 * P2(gs_mol_sha, p, iv, 1) + 1.6 * gb_mol) / (gb_mol * P2(gs_mol_sha, p, iv, 1));
        P2(cisha_z, p, iv, 1) = fmax(P2(cisha_z, p, iv, 1), 1.e-06);
        double gs = P2(gs_mol_sun, p, iv, 1) / cf;
        P2(rssun_z, p, iv, 1) = fmin(1.0 / gs, rsmax0);
        P2(rssun_z, p, iv, 1) = P2(rssun_z, p, iv, 1) / P1(o3coefgsun, p);
        gs = P2(gs_mol_sha, p, iv, 1) / cf;
        P2(rssha_z, p, iv, 1) = fmin(1.0 / gs, rsmax0);
        P2(rssha_z, p, iv, 1) = P2(rssha_z, p, iv, 1) / P1(o3coefgsha, p);
        P2(psnsun_z, p, iv, 1) = P2(ag_phs, p, sun, 1);
        P2(psnsun_z, p, iv, 1) = P2(psnsun_z, p, iv, 1) * P1(o3coefvsun, p);
        L(psn_wc_z_sun, p) = 0.0; L(psn_wj_z_sun, p) = 0.0; L(psn_wp_z_sun, p) = 0.0;
        {
          const double ac = P2(ac_phs, p, sun, 1), aj = P2(aj_phs, p, sun, 1), ap = P2(ap_phs, p, sun, 1);
          if (ac <= aj && ac <= ap) L(psn_wc_z_sun, p) = P2(psnsun_z, p, iv, 1);
          else if (aj < ac && aj <= ap) L(psn_wj_z_sun, p) = P2(psnsun_z, p, iv, 1);
          else if (ap < ac && ap < aj) L(psn_wp_z_sun, p) = P2(psnsun_z, p, iv, 1);
        }
        P2(psnsha_z, p, iv, 1) = P2(ag_phs, p, sha, 1);
        P2(psnsha_z, p, iv, 1) = P2(psnsha_z, p, iv, 1) * P1(o3coefvsha, p);
        L(psn_wc_z_sha, p) = 0.0; L(psn_wj_z_sha, p) = 0.0; L(psn_wp_z_sha, p) = 0.0;
        {
          const double ac = P2(ac_phs, p, sha, 1), aj = P2(aj_phs, p, sha, 1), ap = P2(ap_phs, p, sha, 1);
          if (ac <= aj && ac <= ap) L(psn_wc_z_sha, p) = P2(psnsha_z, p, iv, 1);
          else if (aj < ac && aj <= ap) L(psn_wj_z_sha, p) = P2(psnsha_z, p, iv, 1);
          else if (ap < ac && ap < aj) L(psn_wp_z_sha, p) = P2(psnsha_z, p, iv, 1);
        }
        if (P2(gs_mol_sun, p, iv, 1) < 0.0 || P2(gs_mol_sha, p, iv, 1) < 0.0) fail(x, CTSM_ERR_GS_NEG, p);
        /* the Ball-Berry consistency check (:3692-3710) only writes to the log */
      }
    }
  }

  for (int f = 0; f < fn; ++f) {        /* :3724-3807 canopy sums */
    const int p = filterp[f], ivt = P1(itype, p);
    const int scale_lmr = (PFT(pft_crop, ivt) == 0.0 && pr->modifyphoto_and_lmr_forcrop);
    double psncan = 0.0, psncan_wc = 0.0, psncan_wj = 0.0, psncan_wp = 0.0, lmrcan = 0.0, gscan = 0.0, laican_sun = 0.0;
    for (int iv = 1; iv <= P1(nrad, p); ++iv) {
      const double lz = P2(laisun_z, p, iv, 1);
      psncan = psncan + P2(psnsun_z, p, iv, 1) * lz;
      psncan_wc = psncan_wc + L(psn_wc_z_sun, p) * lz;
      psncan_wj = psncan_wj + L(psn_wj_z_sun, p) * lz;
      psncan_wp = psncan_wp + L(psn_wp_z_sun, p) * lz;
      if (scale_lmr) lmrcan = lmrcan + P2(lmrsun_z, p, iv, 1) * lz * L(bsun, p);
      else lmrcan = lmrcan + P2(lmrsun_z, p, iv, 1) * lz;
      gscan = gscan + lz / (L(rb, p) + P2(rssun_z, p, iv, 1));
      laican_sun = laican_sun + lz;
    }
    if (laican_sun > 0.0) {
      P1(psnsun, p) = psncan / laican_sun;
      P1(psnsun_wc, p) = psncan_wc / laican_sun;
      P1(psnsun_wj, p) = psncan_wj / laican_sun;
      P1(psnsun_wp, p) = psncan_wp / laican_sun;
      P1(lmrsun, p) = lmrcan / laican_sun;
      P1(rssun, p) = laican_sun / gscan - L(rb, p);
    } else {
      P1(psnsun, p) = 0.0; P1(psnsun_wc, p) = 0.0; P1(psnsun_wj, p) = 0.0; P1(psnsun_wp, p) = 0.0;
      P1(lmrsun, p) = 0.0; P1(rssun, p) = 0.0;
    }
    double laican_sha = 0.0;
    psncan = 0.0; psncan_wc = 0.0; psncan_wj = 0.0; psncan_wp = 0.0; lmrcan = 0.0; gscan = 0.0;
    for (int iv = 1; iv <= P1(nrad, p); ++iv) {
      const double lz = P2(laisha_z, p, iv, 1);
      psncan = psncan + P2(psnsha_z, p, iv, 1) * lz;
      psncan_wc = psncan_wc + L(psn_wc_z_sha, p) * lz;
      psncan_wj = psncan_wj + L(psn_wj_z_sha, p) * lz;
      psncan_wp = psncan_wp + L(psn_wp_z_sha, p) * lz;
      if (scale_lmr) lmrcan = lmrcan + P2(lmrsha_z, p, iv, 1) * lz * L(bsha, p);
      else lmrcan = lmrcan + P2(lmrsha_z, p, iv, 1) * lz;
      gscan = gscan + lz / (L(rb, p) + P2(rssha_z, p, iv, 1));
      laican_sha = laican_sha + lz;
    }
    if (laican_sha > 0.0) {
      P1(psnsha, p) = psncan / laican_sha;
      P1(psnsha_wc, p) = psncan_wc / laican_sha;
      P1(psnsha_wj, p) = psncan_wj / laican_sha;
      P1(psnsha_wp, p) = psncan_wp / laican_sha;
      P1(lmrsha, p) = lmrcan / laican_sha;
      P1(rssha, p) = laican_sha / gscan - L(rb, p);
    } else {
      P1(psnsha, p) = 0.0; P1(psnsha_wc, p) = 0.0; P1(psnsha_wj, p) = 0.0; P1(psnsha_wp, p) = 0.0;
      P1(lmrsha, p) = 0.0; P1(rssha, p) = 0.0;
    }
    if (laican_sha + laican_sun > 0.0)
      L(btran, p) = L(bsun, p) * (laican_sun / (laican_sun + laican_sha)) + L(bsha, p) * (laican_sha / (laican_sun + laican_sha));
    else
      L(btran, p) = L(bsun, p);
  }
  free(jmax_z); free(psn_wc_z_sun); free(psn_wj_z_sun); free(psn_wp_z_sun);
  free(psn_wc_z_sha); free(psn_wj_z_sha); free(psn_wp_z_sha);
#undef JMAX
#undef L
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Stand-alone entry points for the pinning tests (tests/test_oracle_phs.py compares them with an independent Python restatement
 * written from the Fortran): one calcstress solve of one patch, and PhotosynthesisHydraulicStress over a filter. */
static void phs_ctx(cf_ctx* x, const ctsm_params_t* prm, const ctsm_canopyfluxes_fields_t* fld) {
  memset(x, 0, sizeof *x);
  x->f = fld; x->prm = prm;
  x->begp0 = fld->alloc.begp; x->begc0 = fld->alloc.begc; x->begg0 = fld->alloc.begg;
  x->ldp = (size_t)(fld->alloc.endp - fld->alloc.begp + 1);
  x->ldc = (size_t)(fld->alloc.endc - fld->alloc.begc + 1);
  x->np = (int)x->ldp;
}

/* xv: vegwp(sun, sha, xyl, root) in / out (4 values, 0-based) */
int oracle_phs_calcstress(const ctsm_params_t* prm, const ctsm_canopyfluxes_fields_t* fld, int p, double* xv4, double* bsun,
                          double* bsha, double gb_mol, double gs_mol_sun, double gs_mol_sha, double qsatl, double qaf) {
  cf_ctx ctx, *x = &ctx;
  phs_ctx(x, prm, fld);
  double xv[5] = {0.0, xv4[0], xv4[1], xv4[2], xv4[3]};
  calcstress(x, p, P1(column, p), xv, bsun, bsha, gb_mol, gs_mol_sun, gs_mol_sha, qsatl, qaf);
  for (int i = 0; i < 4; ++i) xv4[i] = xv[i + 1];
  return x->err_code;
}

/* the (begp0:endp0) work arrays are the caller's: esat_tv, eair, oair, cair, rb, dayl_factor, qsatl, qaf in; bsun, bsha, btran out */
int oracle_phs_standalone(const ctsm_params_t* prm, const ctsm_canopyfluxes_fields_t* fld, int fn, const int32_t* filterp,
                          const double* esat_tv, const double* eair, const double* oair, const double* cair, const double* rb,
                          double* bsun, double* bsha, double* btran, const double* dayl_factor, const double* qsatl,
                          const double* qaf) {
  cf_ctx ctx, *x = &ctx;
  phs_ctx(x, prm, fld);
  x->bbb = (double*)calloc((size_t)x->np, sizeof(double));
  x->mbb = (double*)calloc((size_t)x->np, sizeof(double));
  oracle_photosynthesis_hydraulic_stress(x, fn, filterp, esat_tv, eair, oair, cair, rb, bsun, bsha, btran, dayl_factor, qsatl, qaf);
  free(x->bbb); free(x->mbb);
  return x->err_code;
}
