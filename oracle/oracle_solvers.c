/* oracle_solvers.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * CPU restatement of TridiagonalMod::Tridiagonal, BandDiagonalMod::BandDiagonal
 * and the dgtsv call site of soilwater_moisture_form.  PARITY UNPINNED by the
 * reference's own tests; checked against dense solves in
 * tests/test_oracle_solvers.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static void set_status(ctsm_status_t* st, int code, int level, int index, int info, const char* msg) {
  if (!st) return;
  st->code = code; st->subgrid_level = level; st->subgrid_index = index; st->info = info;
  snprintf(st->msg, sizeof st->msg, "%s", msg);
}

/* TridiagonalMod.F90:23-91.  Arrays (begc:endc, lbj:ubj). */
void oracle_tridiagonal(const ctsm_bounds_t* bounds, int lbj, int ubj, const int32_t* jtop, int numf,
                        const int32_t* filter, const double* a, const double* b, const double* c,
                        const double* r, double* u) {
  const int begc = bounds->begc, endc = bounds->endc;
  const size_t ld = (size_t)(endc - begc + 1);
  const int nl = ubj - lbj + 1;
#define A2(x, ci, j) x[(size_t)((j) - lbj) * ld + ((ci) - begc)]
  double* gam = (double*)malloc(sizeof(double) * ld * nl);   /* :47 */
  double* bet = (double*)malloc(sizeof(double) * ld);        /* :48 */
  for (int fc = 0; fc < numf; ++fc) {                        /* :62-65 */
    const int ci = filter[fc];
    bet[ci - begc] = A2(b, ci, jtop[ci - begc]);
  }
  for (int j = lbj; j <= ubj; ++j) {                         /* :67-80 */
    for (int fc = 0; fc < numf; ++fc) {
      const int ci = filter[fc];
      const int jt = jtop[ci - begc];
      if (j >= jt) {
        if (j == jt) {
          A2(u, ci, j) = A2(r, ci, j) / bet[ci - begc];
        } else {
          A2(gam, ci, j) = A2(c, ci, j - 1) / bet[ci - begc];
          bet[ci - begc] = A2(b, ci, j) - A2(a, ci, j) * A2(gam, ci, j);
          A2(u, ci, j) = (A2(r, ci, j) - A2(a, ci, j) * A2(u, ci, j - 1)) / bet[ci - begc];
        }
      }
    }
  }
  for (int j = ubj - 1; j >= lbj; --j) {                     /* :82-89 */
    for (int fc = 0; fc < numf; ++fc) {
      const int ci = filter[fc];
      if (j >= jtop[ci - begc]) A2(u, ci, j) = A2(u, ci, j) - A2(gam, ci, j + 1) * A2(u, ci, j + 1);
    }
  }
  free(gam); free(bet);
#undef A2
}

/* BandDiagonalMod.F90:29-221.  b (begc:endc, nband, lbj:ubj); r,u (begc:endc, lbj:ubj). */
int oracle_banddiagonal(const ctsm_bounds_t* bounds, int lbj, int ubj, const int32_t* jtop,
                        const int32_t* jbot, int numf, const int32_t* filter, int nband,
                        const double* b, const double* r, double* u, ctsm_status_t* st) {
  const int begc = bounds->begc, endc = bounds->endc;
  const size_t ld = (size_t)(endc - begc + 1);
  (void)ubj;
#define B3(ci, k, j) b[((size_t)((j) - lbj) * nband + ((k) - 1)) * ld + ((ci) - begc)]
#define R2(x, ci, j) x[(size_t)((j) - lbj) * ld + ((ci) - begc)]
  int rc = 0;
  if (st) memset(st, 0, sizeof *st);
  for (int fc = 0; fc < numf; ++fc) {                        /* :167 */
    const int ci = filter[fc];
    const int kl = (nband - 1) / 2, ku = kl;                 /* :170-171 */
    const int m = 2 * kl + ku + 1;                           /* :173 */
    const int jt = jtop[ci - begc], jb = jbot[ci - begc];
    const int n = jb - jt + 1;                               /* :176 */
    double* ab = (double*)calloc((size_t)m * n, sizeof(double));   /* :178-179 */
    int* ipiv = (int*)malloc(sizeof(int) * n);
    double* result = (double*)malloc(sizeof(double) * n);
#define ABm(i, jc) ab[((i) - 1) + (size_t)((jc) - 1) * m]
    /* :181-185 (written for nband=5: rows kl+ku-1 .. kl+ku+3) */
    for (int jc = 3; jc <= n; ++jc) ABm(kl + ku - 1, jc) = B3(ci, 1, jt + jc - 3);
    for (int jc = 2; jc <= n; ++jc) ABm(kl + ku + 0, jc) = B3(ci, 2, jt + jc - 2);
    for (int jc = 1; jc <= n; ++jc) ABm(kl + ku + 1, jc) = B3(ci, 3, jt + jc - 1);
    for (int jc = 1; jc <= n - 1; ++jc) ABm(kl + ku + 2, jc) = B3(ci, 4, jt + jc);
    for (int jc = 1; jc <= n - 2; ++jc) ABm(kl + ku + 3, jc) = B3(ci, 5, jt + jc + 1);
    for (int i = 0; i < n; ++i) result[i] = R2(r, ci, jt + i);    /* :194 */
    int info = 0;
    oracle_dgbsv(n, kl, ku, 1, ab, m, ipiv, result, n, &info);     /* :197 */
    for (int i = 0; i < n; ++i) R2(u, ci, jt + i) = result[i];     /* :198 */
    free(ab); free(ipiv); free(result);
#undef ABm
    if (info != 0 && rc == 0) {                                    /* :200-213 endrun */
      rc = CTSM_ERR_DGBSV;
      set_status(st, rc, CTSM_SUBGRID_COLUMN, ci, info, "BandDiagonal ERROR: dgbsv returned error code");
      break;   /* endrun aborts the model at the first failing column */
    }
  }
  return rc;
#undef B3
#undef R2
}

/* SoilWaterMovementMod.F90:1279-1299: gather, dgtsv, scatter, endrun on err/=0. */
int oracle_dgtsv_batch(const ctsm_bounds_t* bounds, int nlev, const int32_t* nlayers, int numf,
                       const int32_t* filter, const double* amx, const double* bmx, const double* cmx,
                       const double* rmx, double* x, ctsm_status_t* st) {
  const int begc = bounds->begc, endc = bounds->endc;
  const size_t ld = (size_t)(endc - begc + 1);
#define M2(a, ci, j) a[(size_t)((j) - 1) * ld + ((ci) - begc)]
  double* dLow = (double*)malloc(sizeof(double) * nlev);
  double* dUpp = (double*)malloc(sizeof(double) * nlev);
  double* diag = (double*)malloc(sizeof(double) * nlev);
  double* rhs = (double*)malloc(sizeof(double) * nlev);
  int rc = 0;
  if (st) memset(st, 0, sizeof *st);
  for (int fc = 0; fc < numf; ++fc) {
    const int ci = filter[fc];
    const int n = nlayers[ci - begc];
    for (int j = 1; j <= n - 1; ++j) dLow[j - 1] = M2(amx, ci, j + 1);   /* :1281 */
    for (int j = 1; j <= n; ++j) diag[j - 1] = M2(bmx, ci, j);           /* :1282 */
    for (int j = 1; j <= n - 1; ++j) dUpp[j - 1] = M2(cmx, ci, j);       /* :1283 */
    for (int j = 1; j <= n; ++j) rhs[j - 1] = M2(rmx, ci, j);            /* :1286 */
    int err = 0;
    oracle_dgtsv(n, 1, dLow, diag, dUpp, rhs, n, &err);                  /* :1287 */
    if (err != 0) {                                                      /* :1295 */
      rc = CTSM_ERR_DGTSV;
      set_status(st, rc, CTSM_SUBGRID_COLUMN, ci, err, "soilwater_moisture_form:: problem with the lapack solver");
      break;
    }
    for (int j = 1; j <= n; ++j) M2(x, ci, j) = rhs[j - 1];              /* :1299 */
  }
  free(dLow); free(dUpp); free(diag); free(rhs);
  return rc;
#undef M2
}

#ifdef _OPENMP
#include <omp.h>
int oracle_num_threads(void) { return omp_get_max_threads(); }
/* launchers such as torch.distributed.run export OMP_NUM_THREADS=1; the CPU arm must use the host cores it is given */
void oracle_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
#else
int oracle_num_threads(void) { return 1; }
void oracle_set_num_threads(int n) { (void)n; }
#endif
