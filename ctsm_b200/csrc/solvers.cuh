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
// Generic form: the finished rows go to `st` (st.put_u(j, k, v), st.put_y(j, v); read back with st.u(i, k), st.y(i)) and
// every solution component is handed to `putx(i, x_i, x_{i+1})` as the back substitution produces it (i descending;
// x_{i+1} = 0 for i = n-1).  Storage can therefore live anywhere (per-thread arrays, a coalesced global scratch).
template <typename RowFn, typename Store, typename PutX>
__device__ __forceinline__ int band5_solve_stream(int n, RowFn row, Store& st, PutX putx, int lead = 0) {
  // One loop, one call site of `row`: iteration g fetches row i = g - lead into the bottom slot of the 3x5 window and then
  // runs elimination step j = i - 2 (rows j..j+2 are in the window by then); the first two iterations only slide.
  // `lead` idle iterations in front let the threads of a warp run DIFFERENT system sizes aligned at their LAST row
  // (n + lead is then the same for every lane: SoilTemperature aligns soil level 1 across columns with 0..12 snow layers).
  double w0[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, w1[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, w2[5];
  double r0 = 0.0, r1 = 0.0, r2;
  double e[5];
  int info = 0;
  const int total = n + 2 + lead;
  for (int g = 0; g < total; ++g) {
    const int j = g - lead - 2;
    if (j < -2) continue;
    if (j + 2 < n) {
      r2 = row(j + 2, e);
      // columns j .. j+4 of row j+2 (entries left of column 0 or right of column n-1 do not exist)
      w2[0] = (j >= 0) ? e[0] : 0.0;
      w2[1] = (j + 1 >= 0) ? e[1] : 0.0;
      w2[2] = e[2];
      w2[3] = (j + 3 < n) ? e[3] : 0.0;
      w2[4] = (j + 4 < n) ? e[4] : 0.0;
    } else {
      r2 = 0.0; w2[0] = w2[1] = w2[2] = w2[3] = w2[4] = 0.0;
    }
    if (j >= 0) {
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
      for (int k = 0; k < 5; ++k) st.put_u(j, k, w0[k]);
      st.put_y(j, r0);
    }
    // slide the window one column: row j+1 -> slot 0, row j+2 -> slot 1
#pragma unroll
    for (int k = 0; k < 4; ++k) { w0[k] = w1[k + 1]; w1[k] = w2[k + 1]; }
    w0[4] = 0.0; w1[4] = 0.0;
    r0 = r1; r1 = r2;
  }
  if (info != 0) return info;   // dgbsv skips the solve; the reference then aborts
  // dtbsv('U','N','N'), k = 4
  double x1 = 0.0, x2 = 0.0, x3 = 0.0, x4 = 0.0;   // x(i+1) .. x(i+4)
  for (int i = n - 1; i >= 0; --i) {
    double acc = st.y(i);
    if (i + 4 < n && x4 != 0.0) acc = acc - x4 * st.u(i, 4);
    if (i + 3 < n && x3 != 0.0) acc = acc - x3 * st.u(i, 3);
    if (i + 2 < n && x2 != 0.0) acc = acc - x2 * st.u(i, 2);
    if (i + 1 < n && x1 != 0.0) acc = acc - x1 * st.u(i, 1);
    if (acc != 0.0) acc = acc / st.u(i, 0);
    putx(i, acc, x1);
    x4 = x3; x3 = x2; x2 = x1; x1 = acc;
  }
  return 0;
}

// per-thread array storage (the original interface): U[j][0..4], y[j]; on exit y holds the solution
struct Band5ArrayStore {
  double (*U)[5];
  double* yv;
  __device__ __forceinline__ void put_u(int j, int k, double v) { U[j][k] = v; }
  __device__ __forceinline__ void put_y(int j, double v) { yv[j] = v; }
  __device__ __forceinline__ double u(int i, int k) const { return U[i][k]; }
  __device__ __forceinline__ double y(int i) const { return yv[i]; }
};
template <int MAXN, typename RowFn>
__device__ __forceinline__ int band5_solve(int n, RowFn row, double (*U)[5], double* y) {
  Band5ArrayStore st{U, y};
  return band5_solve_stream(n, row, st, [&](int i, double x, double) { y[i] = x; });
}
