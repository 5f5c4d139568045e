// phs_direct.cuh — TEST INFRASTRUCTURE.  The direct (nested-loop) formulation of brent_PHS / hybrid_PHS that the
// first CUDA path used and that was parity-checked against the oracle on B200.  tests/host/phs_tasks_check.cu runs
// it on the CPU against the resumable task formulation of ctsm_b200/csrc/phs.cuh; the two must agree bit for bit.
// Reference: src/biogeophys/PhotosynthesisMod.F90 hybrid_PHS :3815-4064, brent_PHS :4068-4223.
#pragma once
#include "../../ctsm_b200/csrc/phs.cuh"
namespace phs {
// brent_PHS :4068-4223
PHS_FN void brent(const PhsPatch& P, const Leaf& L, double& xsun, double x1sun, double x2sun, double f1sun,
                                   double f2sun, double& xsha, double x1sha, double x2sha, double f1sha, double f2sha, double tol,
                                   double& gs_sun, double& gs_sha, double bsun, double bsha, CiOut& o, bool* bad, bool* notbracketed) {
  double a[2] = {x1sun, x1sha}, b[2] = {x2sun, x2sha}, c[2], d[2] = {0.0, 0.0}, e[2] = {0.0, 0.0};
  double fa[2] = {f1sun, f1sha}, fb[2] = {f2sun, f2sha}, fc[2], tol1[2], xm[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if ((fa[s] > 0.0 && fb[s] > 0.0) || (fa[s] < 0.0 && fb[s] < 0.0)) *notbracketed = true;
    c[s] = b[s]; fc[s] = fb[s];
  }
  for (int iter = 0; iter < 20;) {
    ++iter;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if ((fb[s] > 0.0 && fc[s] > 0.0) || (fb[s] < 0.0 && fc[s] < 0.0)) { c[s] = a[s]; fc[s] = fa[s]; d[s] = b[s] - a[s]; e[s] = d[s]; }
      if (fabs(fc[s]) < fabs(fb[s])) { a[s] = b[s]; b[s] = c[s]; c[s] = a[s]; fa[s] = fb[s]; fb[s] = fc[s]; fc[s] = fa[s]; }
      tol1[s] = 2.0 * 1.e-4 * fabs(b[s]) + 0.5 * tol;
      xm[s] = 0.5 * (c[s] - b[s]);
    }
    if ((fabs(xm[0]) <= tol1[0] || fb[0] == 0.0) && (fabs(xm[1]) <= tol1[1] || fb[1] == 0.0)) { xsun = b[0]; xsha = b[1]; return; }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (fabs(e[s]) >= tol1[s] && fabs(fa[s]) > fabs(fb[s])) {
        const double sv = fb[s] / fa[s];
        double pv, qv;
        if (a[s] == c[s]) {
          pv = 2.0 * xm[s] * sv;
          qv = 1.0 - sv;
        } else {
          qv = fa[s] / fc[s];
          const double rv = fb[s] / fc[s];
          pv = sv * (2.0 * xm[s] * qv * (qv - rv) - (b[s] - a[s]) * (rv - 1.0));
          qv = (qv - 1.0) * (rv - 1.0) * (sv - 1.0);
        }
        if (pv > 0.0) qv = -qv;
        pv = fabs(pv);
        if (2.0 * pv < fmin(3.0 * xm[s] * qv - fabs(tol1[s] * qv), fabs(e[s] * qv))) { e[s] = d[s]; d[s] = pv / qv; }
        else { d[s] = xm[s]; e[s] = d[s]; }
      } else {
        d[s] = xm[s]; e[s] = d[s];
      }
      a[s] = b[s]; fa[s] = fb[s];
      if (fabs(d[s]) > tol1[s]) b[s] = b[s] + d[s];
      else b[s] = b[s] + copysign(tol1[s], xm[s]);
    }
    ci_func(P, L, b[0], b[1], bsun, bsha, fb[0], fb[1], gs_sun, gs_sha, o, bad);
    if ((fb[0] == 0.0) && (fb[1] == 0.0)) break;
  }
  xsun = b[0]; xsha = b[1];
}

struct HybridOut { double bsun, bsha, gs_sun, gs_sha, tran; double x[4]; };

// hybrid_PHS :3815-4064.  vegwp_in = canopystate_inst%vegwp_patch(p,:) at entry.
PHS_FN HybridOut hybrid(const PhsPatch& P, const Leaf& L, const double* vegwp_in, double ci0, CiOut& o,
                                            bool* bad, bool* notbracketed) {
  HybridOut h;
  double x[4];
  double x0sun, x0sha, x1sun = ci0, x1sha = ci0, f0sun = 0.0, f0sha = 0.0, f1sun = 0.0, f1sha = 0.0;
  double gs0sun = 0.0, gs0sha = 0.0, gs_sun = 0.0, gs_sha = 0.0, bsun = 1.0, bsha = 1.0, b0sun = -1.0, b0sha = -1.0;
  double minf = 0.0, minxsun = 0.0, minxsha = 0.0, unused_tran = 0.0;
  bool bflag = false;
  for (int iter1 = 1;; ++iter1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = vegwp_in[i];
    int iter2 = 0;
    x0sun = fmax(0.1, x1sun); x1sun = 0.99 * x1sun;
    x0sha = fmax(0.1, x1sha); x1sha = 0.99 * x1sha;
    const double tolsun = fabs(x1sun) * 1.e-2, tolsha = fabs(x1sha) * 1.e-2;
    if (bflag) {                                      // ci_func_PHS prologue :4306-4311
      const Stress s = calcstress(P, x, gs0sun, gs0sha, &unused_tran);
      bsun = s.bsun; bsha = s.bsha;
    }
    ci_func(P, L, x0sun, x0sha, bsun, bsha, f0sun, f0sha, gs_sun, gs_sha, o, bad);
    const double dbsun = b0sun - bsun, dbsha = b0sha - bsha;
    b0sun = bsun; b0sha = bsha;
    bflag = false;
    ci_func(P, L, x1sun, x1sha, bsun, bsha, f1sun, f1sha, gs_sun, gs_sha, o, bad);
    for (;;) {
      if ((fabs(f0sun) < 1.e-4) && (fabs(f0sha) < 1.e-4)) { x1sun = x0sun; x1sha = x0sha; break; }
      if ((fabs(f1sun) < 1.e-4) && (fabs(f1sha) < 1.e-4)) break;
      iter2 = iter2 + 1;
      const double dxsun = ((f1sun - f0sun) == 0.0) ? 0.5 * (x1sun + x0sun) - x1sun : -f1sun * (x1sun - x0sun) / (f1sun - f0sun);
      const double dxsha = ((f1sha - f0sha) == 0.0) ? 0.5 * (x1sha + x0sha) - x1sha : -f1sha * (x1sha - x0sha) / (f1sha - f0sha);
      x0sun = x1sun; x1sun = x1sun + dxsun;
      x0sha = x1sha; x1sha = x1sha + dxsha;
      ci_func(P, L, x1sun, x1sha, bsun, bsha, f1sun, f1sha, gs_sun, gs_sha, o, bad);
      if ((fabs(dxsun) < tolsun) && (fabs(dxsha) < tolsha)) { x0sun = x1sun; x0sha = x1sha; break; }
      if (iter2 == 1 || fabs(f1sun + f1sha) < minf) { minf = fabs(f1sun + f1sha); minxsun = x1sun; minxsha = x1sha; }
      if ((fabs(f1sun) < 1.e-4) && (fabs(f1sha) < 1.e-4)) break;
      if ((f1sun * f0sun < 0.0) && (f1sha * f0sha < 0.0)) {
        double xs, xh;
        brent(P, L, xs, x0sun, x1sun, f0sun, f1sun, xh, x0sha, x1sha, f0sha, f1sha, tolsun, gs_sun, gs_sha, bsun, bsha, o, bad,
              notbracketed);
        x0sun = xs; x0sha = xh;
        break;
      }
      if (iter2 > 3) {
        x1sun = minxsun; x1sha = minxsha;
        ci_func(P, L, x1sun, x1sha, bsun, bsha, f1sun, f1sha, gs_sun, gs_sha, o, bad);
        break;
      }
    }
    if (bsun > 0.01) gs0sun = gs_sun / bsun;
    if (bsha > 0.01) gs0sha = gs_sha / bsha;
    bflag = true;
    if ((fabs(dbsun) < 1.e-2) && (fabs(dbsha) < 1.e-2)) break;
    if (iter1 > 3) break;
  }
  double sf = getvegwp(P, x, gs_sun, gs_sha);           // :4048-4050
  if (sf < 0.0) sf = 0.0;
  h.bsun = bsun; h.bsha = bsha; h.gs_sun = gs_sun; h.gs_sha = gs_sha; h.tran = sf;
#pragma unroll
  for (int i = 0; i < 4; ++i) h.x[i] = x[i];
  return h;
}

}  // namespace phs
