/* oracle_lapack.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the two LAPACK drivers the reference's hot path calls:
 *   dgbsv  <- src/biogeophys/BandDiagonalMod.F90:197      (n<=38, kl=ku=2, nrhs=1)
 *   dgtsv  <- src/biogeophys/SoilWaterMovementMod.F90:1287 (n<=20, nrhs=1)
 *
 * LAPACK itself is a third-party dependency that is NOT vendored under
 * /root/reference (linked through ESMF / MKL, version unpinned).
 * What is restated here is the published reference-LAPACK (netlib 3.x)
 * algorithm: dgbsv -> dgbtrf -> dgbtf2 (unblocked, because ILAENV returns
 * NB=1 for KU<=64 so NB<=1) + dgbtrs('N') -> dger/dtbsv; dgtsv as in dgtsv.f.
 * Loop order and operation order follow those sources so that results agree
 * with any non-FMA build of reference LAPACK bit for bit.
 *
 * Pinning: tests/test_oracle_lapack.py compares these routines against
 * scipy.linalg.lapack.{dgbsv,dgtsv} (OpenBLAS build of the same netlib code;
 * its BLAS-1/2 kernels may use FMA, so agreement is asserted to a few ulp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use it.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define AB(i, j) ab[((i) - 1) + (size_t)((j) - 1) * ldab]

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* dgbtf2.f: unblocked LU of a general band matrix with partial pivoting. */
static void oracle_dgbtf2(int m, int n, int kl, int ku, double* ab, int ldab, int* ipiv, int* info) {
  const int kv = ku + kl;
  *info = 0;
  /* set fill-in elements in columns ku+2 .. kv to zero */
  for (int j = ku + 2; j <= imin(kv, n); ++j)
    for (int i = kv - j + 2; i <= kl; ++i) AB(i, j) = 0.0;
  int ju = 1;
  for (int j = 1; j <= imin(m, n); ++j) {
    if (j + kv <= n)
      for (int i = 1; i <= kl; ++i) AB(i, j + kv) = 0.0;
    const int km = imin(kl, m - j);
    /* idamax(km+1, ab(kv+1,j), 1): first index of the largest |x| */
    int jp = 1;
    double dmax = fabs(AB(kv + 1, j));
    for (int i = 2; i <= km + 1; ++i) {
      const double v = fabs(AB(kv + i, j));
      if (v > dmax) { dmax = v; jp = i; }
    }
    ipiv[j - 1] = jp + j - 1;
    if (AB(kv + jp, j) != 0.0) {
      ju = imax(ju, imin(j + ku + jp - 1, n));
      /* dswap(ju-j+1, ab(kv+jp,j), ldab-1, ab(kv+1,j), ldab-1) */
      if (jp != 1) {
        for (int k = 0; k < ju - j + 1; ++k) {
          double* x = &AB(kv + jp, j) + (size_t)k * (ldab - 1);
          double* y = &AB(kv + 1, j) + (size_t)k * (ldab - 1);
          const double t = *x; *x = *y; *y = t;
        }
      }
      if (km > 0) {
        /* dscal(km, one/ab(kv+1,j), ab(kv+2,j), 1) */
        const double rp = 1.0 / AB(kv + 1, j);
        for (int i = 0; i < km; ++i) AB(kv + 2 + i, j) = rp * AB(kv + 2 + i, j);
        /* dger(km, ju-j, -one, ab(kv+2,j),1, ab(kv,j+1),ldab-1, ab(kv+1,j+1),ldab-1) */
        if (ju > j) {
          for (int jj = 0; jj < ju - j; ++jj) {
            const double y = *(&AB(kv, j + 1) + (size_t)jj * (ldab - 1));
            if (y != 0.0) {
              const double temp = -1.0 * y;
              double* acol = &AB(kv + 1, j + 1) + (size_t)jj * (ldab - 1);
              for (int i = 0; i < km; ++i) acol[i] = acol[i] + AB(kv + 2 + i, j) * temp;
            }
          }
        }
      }
    } else if (*info == 0) {
      *info = j;
    }
  }
}

/* dgbtrs.f, TRANS='N', followed by dtbsv('Upper','No transpose','Non-unit'). */
static void oracle_dgbtrs_n(int n, int kl, int ku, const double* ab, int ldab, const int* ipiv, double* b) {
  const int kd = ku + kl + 1;
  if (kl > 0) {
    for (int j = 1; j <= n - 1; ++j) {
      const int lm = imin(kl, n - j);
      const int l = ipiv[j - 1];
      if (l != j) { const double t = b[l - 1]; b[l - 1] = b[j - 1]; b[j - 1] = t; }
      /* dger(lm, 1, -one, ab(kd+1,j),1, b(j),ldb, b(j+1),ldb) */
      if (b[j - 1] != 0.0) {
        const double temp = -1.0 * b[j - 1];
        for (int i = 0; i < lm; ++i) b[j + i] = b[j + i] + AB(kd + 1 + i, j) * temp;
      }
    }
  }
  /* dtbsv upper / no-trans / non-unit with k = kl+ku super-diagonals */
  const int k = kl + ku;
  const int kplus1 = k + 1;
  for (int j = n; j >= 1; --j) {
    if (b[j - 1] != 0.0) {
      const int l = kplus1 - j;
      b[j - 1] = b[j - 1] / AB(kplus1, j);
      const double temp = b[j - 1];
      for (int i = j - 1; i >= imax(1, j - k); --i) b[i - 1] = b[i - 1] - temp * AB(l + i, j);
    }
  }
}

void oracle_dgbsv(int n, int kl, int ku, int nrhs, double* ab, int ldab, int* ipiv, double* b, int ldb, int* info) {
  (void)ldb;
  *info = 0;
  if (n < 0) { *info = -1; return; }
  if (kl < 0) { *info = -2; return; }
  if (ku < 0) { *info = -3; return; }
  if (nrhs != 1) { *info = -4; return; }   /* only the reference's nrhs=1 call shape is restated */
  if (ldab < 2 * kl + ku + 1) { *info = -6; return; }
  oracle_dgbtf2(n, n, kl, ku, ab, ldab, ipiv, info);
  if (*info == 0) oracle_dgbtrs_n(n, kl, ku, ab, ldab, ipiv, b);
}

/* dgtsv.f (nrhs = 1 branch of the NRHS<=2 loop). dl, d, du, b are overwritten. */
void oracle_dgtsv(int n, int nrhs, double* dl, double* d, double* du, double* b, int ldb, int* info) {
  (void)ldb;
  *info = 0;
  if (n < 0) { *info = -1; return; }
  if (nrhs != 1) { *info = -2; return; }
  if (n == 0) return;
  /* 0-based: dl[i] couples row i+1 to row i */
  for (int i = 0; i < n - 2; ++i) {
    if (fabs(d[i]) >= fabs(dl[i])) {
      /* no row interchange required */
      if (d[i] != 0.0) {
        const double fact = dl[i] / d[i];
        d[i + 1] = d[i + 1] - fact * du[i];
        b[i + 1] = b[i + 1] - fact * b[i];
      } else {
        *info = i + 1;
        return;
      }
      dl[i] = 0.0;
    } else {
      /* interchange rows i and i+1 */
      const double fact = d[i] / dl[i];
      d[i] = dl[i];
      double temp = d[i + 1];
      d[i + 1] = du[i] - fact * temp;
      dl[i] = du[i + 1];
      du[i + 1] = -fact * dl[i];
      du[i] = temp;
      temp = b[i];
      b[i] = b[i + 1];
      b[i + 1] = temp - fact * b[i + 1];
    }
  }
  if (n > 1) {
    const int i = n - 2;
    if (fabs(d[i]) >= fabs(dl[i])) {
      if (d[i] != 0.0) {
        const double fact = dl[i] / d[i];
        d[i + 1] = d[i + 1] - fact * du[i];
        b[i + 1] = b[i + 1] - fact * b[i];
      } else {
        *info = i + 1;
        return;
      }
    } else {
      const double fact = d[i] / dl[i];
      d[i] = dl[i];
      double temp = d[i + 1];
      d[i + 1] = du[i] - fact * temp;
      du[i] = temp;
      temp = b[i];
      b[i] = b[i + 1];
      b[i + 1] = temp - fact * b[i + 1];
    }
  }
  if (d[n - 1] == 0.0) { *info = n; return; }
  /* back solve with the matrix U from the factorization */
  b[n - 1] = b[n - 1] / d[n - 1];
  if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / d[n - 2];
  for (int i = n - 3; i >= 0; --i) b[i] = (b[i] - du[i] * b[i + 1] - dl[i] * b[i + 2]) / d[i];
}
