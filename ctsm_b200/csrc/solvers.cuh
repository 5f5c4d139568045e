// solvers.cuh — per-thread (one column per thread) linear solvers.
//
// Each routine reproduces the operation order of the algorithm the reference
// calls, so results agree with a non-FMA CPU build to round-off identity:
//   thomas_solve     TridiagonalMod.F90:62-89 (no pivoting)
//   dgtsv_solve      LAPACK dgtsv, nrhs = 1 (SoilWaterMovementMod.F90:1287)
//   band5_solve      LAPACK dgbsv(kl=ku=2, nrhs=1) = dgbtf2 + dgbtrs('N')
//                    (BandDiagonalMod.F90:197)
// The code is compiled with -fmad=false (the reference builds with
// -ffp-contract=off).
#pragma once
#include "common.cuh"

// LAPACK dgtsv on per-thread arrays (0-based, length n).  dl[i] couples row i+1
// to row i.  On exit b holds the solution.  Returns LAPACK info (0 = ok).
__device__ __forceinline__ int dgtsv_solve(int n, double* dl, double* d, double* du, double* b) {
  if (n == 0) return 0;
  for (int i = 0; i < n - 2; ++i) {
    if (fabs(d[i]) >= fabs(dl[i])) {
      if (d[i] != 0.0) {
        const double fact = dl[i] / d[i];
        d[i + 1] = d[i + 1] - fact * du[i];
        b[i + 1] = b[i + 1] - fact * b[i];
      } else {
        return i + 1;
      }
      dl[i] = 0.0;
    } else {
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
        return i + 1;
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
  if (d[n - 1] == 0.0) return n;
  b[n - 1] = b[n - 1] / d[n - 1];
  if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / d[n - 2];
  for (int i = n - 3; i >= 0; --i) b[i] = (b[i] - du[i] * b[i + 1] - dl[i] * b[i + 2]) / d[i];
  return 0;
}

// Streaming LU with partial pivoting for a 5-band matrix (kl = ku = 2), one
// right-hand side, with the arithmetic of dgbtf2 + dgbtrs('N') + dtbsv.
//
// The matrix is consumed row by row through `row(i, e)` which fills
// e[0..4] = A(i, i-2 .. i+2) (entries outside the matrix are ignored) and
// returns the right-hand side b(i).  Only a 3x5 window of partially eliminated
// rows is live; the finished rows of U (diagonal + kl+ku = 4 super-diagonals,
// which is where dgbtf2 stores its fill-in) and the permuted/eliminated
// right-hand side go to the per-thread arrays U and y for the back substitution.
//
// Equivalence with dgbtf2's column sweep: at elimination step j the candidates
// are rows j..j+2 (idamax: first largest |a|), rows are interchanged over the
// whole remaining band (entries beyond LAPACK's `ju` are still zero in both
// rows), multipliers are a*(1/pivot) (dscal) and every trailing element
// receives a + l*(-u) (dger with alpha = -1) from step j-2 first, then j-1 —
// the same sequence of operations per element.  dgbtrs applies the same
// interchanges/multipliers to b after the factorisation; doing it alongside
// changes no operand.  dtbsv's column-oriented back substitution subtracts
// x(j)*U(i,j) from y(i) for j = i+4 down to i+1, reproduced below.
template <int MAXN, typename RowFn>
__device__ __forceinline__ int band5_solve(int n, RowFn row, double (*U)[5], double* y) {
  double w0[5], w1[5], w2[5];
  double r0, r1, r2;
  double e[5];
  int info = 0;
  // prime the window with rows 0, 1, 2 aligned so that index 0 is column j = 0
  r0 = row(0, e);
  w0[0] = e[2]; w0[1] = (n > 1) ? e[3] : 0.0; w0[2] = (n > 2) ? e[4] : 0.0; w0[3] = 0.0; w0[4] = 0.0;
  if (n > 1) {
    r1 = row(1, e);
    w1[0] = e[1]; w1[1] = e[2]; w1[2] = (n > 2) ? e[3] : 0.0; w1[3] = (n > 3) ? e[4] : 0.0; w1[4] = 0.0;
  } else { r1 = 0.0; w1[0] = w1[1] = w1[2] = w1[3] = w1[4] = 0.0; }
  if (n > 2) {
    r2 = row(2, e);
    w2[0] = e[0]; w2[1] = e[1]; w2[2] = e[2]; w2[3] = (n > 3) ? e[3] : 0.0; w2[4] = (n > 4) ? e[4] : 0.0;
  } else { r2 = 0.0; w2[0] = w2[1] = w2[2] = w2[3] = w2[4] = 0.0; }

  for (int j = 0; j < n; ++j) {
    const int km = min(2, n - 1 - j);
    // idamax over rows j..j+km of column j
    int jp = 0;
    double dmax = fabs(w0[0]);
    if (km >= 1 && fabs(w1[0]) > dmax) { dmax = fabs(w1[0]); jp = 1; }
    if (km >= 2 && fabs(w2[0]) > dmax) { dmax = fabs(w2[0]); jp = 2; }
    if (jp == 1) {
#pragma unroll
      for (int k = 0; k < 5; ++k) { const double t = w0[k]; w0[k] = w1[k]; w1[k] = t; }
      const double t = r0; r0 = r1; r1 = t;
    } else if (jp == 2) {
#pragma unroll
      for (int k = 0; k < 5; ++k) { const double t = w0[k]; w0[k] = w2[k]; w2[k] = t; }
      const double t = r0; r0 = r2; r2 = t;
    }
    if (w0[0] != 0.0) {
      if (km > 0) {
        const double rp = 1.0 / w0[0];
        if (km >= 1) {
          const double l = rp * w1[0];
#pragma unroll
          for (int k = 1; k < 5; ++k) w1[k] = w1[k] + l * (-1.0 * w0[k]);
          r1 = r1 + l * (-1.0 * r0);
        }
        if (km >= 2) {
          const double l = rp * w2[0];
#pragma unroll
          for (int k = 1; k < 5; ++k) w2[k] = w2[k] + l * (-1.0 * w0[k]);
          r2 = r2 + l * (-1.0 * r0);
        }
      }
    } else if (info == 0) {
      info = j + 1;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) U[j][k] = w0[k];
    y[j] = r0;
    // slide the window: row j+1 -> slot 0, row j+2 -> slot 1, fetch row j+3
#pragma unroll
    for (int k = 0; k < 4; ++k) { w0[k] = w1[k + 1]; w1[k] = w2[k + 1]; }
    w0[4] = 0.0; w1[4] = 0.0;
    r0 = r1; r1 = r2;
    if (j + 3 < n) {
      r2 = row(j + 3, e);
      // columns j+1 .. j+5 of row j+3
      w2[0] = e[0]; w2[1] = e[1]; w2[2] = e[2];
      w2[3] = (j + 4 < n) ? e[3] : 0.0;
      w2[4] = (j + 5 < n) ? e[4] : 0.0;
    } else {
      r2 = 0.0; w2[0] = w2[1] = w2[2] = w2[3] = w2[4] = 0.0;
    }
  }
  if (info != 0) return info;   // dgbsv skips the solve; the reference then aborts
  // dtbsv('U','N','N'), k = 4
  double x1 = 0.0, x2 = 0.0, x3 = 0.0, x4 = 0.0;   // x(i+1) .. x(i+4)
  for (int i = n - 1; i >= 0; --i) {
    double acc = y[i];
    if (i + 4 < n && x4 != 0.0) acc = acc - x4 * U[i][4];
    if (i + 3 < n && x3 != 0.0) acc = acc - x3 * U[i][3];
    if (i + 2 < n && x2 != 0.0) acc = acc - x2 * U[i][2];
    if (i + 1 < n && x1 != 0.0) acc = acc - x1 * U[i][1];
    if (acc != 0.0) acc = acc / U[i][0];
    y[i] = acc;
    x4 = x3; x3 = x2; x2 = x1; x1 = acc;
  }
  return 0;
}
