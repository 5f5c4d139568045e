// phs.cuh — per-thread (one patch per thread) plant-hydraulic-stress photosynthesis.
//
// Reference: src/biogeophys/PhotosynthesisMod.F90
//   hybrid_PHS :3815-4064   brent_PHS :4068-4223   ci_func_PHS :4227-4486
//   calcstress :4490-4710   spacA :4715-4893       spacF :4898-4976
//   getvegwp :4979-5077     getqflx :5080-5164     plc :5167   d1plc :5199
//   quadratic  src/utils/quadraticMod.F90:17-74
//
// B200 mapping (not the reference's call structure):
//  * everything a patch needs is gathered ONCE into a register-resident PhsPatch
//    (pft segment parameters, leaf areas, conversion factors) instead of being
//    re-read through p/c/ivt indirection inside every callee;
//  * the 20-level root-zone vectors k_soil_root(p,:), smp_l(c,:), 1000*z(c,:) are
//    staged in shared memory ([level][thread], conflict-free) for the two sums that
//    depend on the root potential; the level sums that do not depend on it
//    (sum k, sum k*smp, sum k*(smp-grav2)) are formed once per call;
//  * spacF and spacA are evaluated at the same potentials, so the vulnerability
//    curve and its derivative are evaluated once per Newton step (Weibull: one
//    pow + one exp2 per segment) and shared by both;
//  * fmad is off (the reference is built with -ffp-contract=off).
#pragma once
#include "common.cuh"

namespace phs {

constexpr double rgas = 6.02214e26 * 1.38065e-23;   // SHR_CONST_AVOGAD*SHR_CONST_BOLTZ
constexpr double tfrz = 273.15;
constexpr double spval = 1.e36;
constexpr double max_cs = 1.e-06;                    // :89
constexpr double tol_lai = .001;                     // :4543

enum Seg { SUN = 0, SHA = 1, XYL = 2, ROOT = 3 };

// Powers.  Every pow on this path has a positive base (ratios of potentials, -zeta, saturations), so
// a**b = exp2(b*log2(a)) holds; libdevice's exp2/log2 (<= 1-2 ulp each) cost about a third of its
// correctly-rounded-to-1-ulp pow, and the Newton solve of calcstress spends most of its instructions in the
// eight powers of the Weibull curve.  The result differs from glibc's pow by a few ulp (|b*log2 a| <= ~50),
// far inside the 1e-10 parity tolerance (tests/test_gpu_canopy.py); -DPHS_LIBM_POW restores pow().
__device__ __forceinline__ double pw(double a, double b) {
#ifdef PHS_LIBM_POW
  return pow(a, b);
#else
  return exp2(b * log2(a));
#endif
}
__device__ __forceinline__ double pw2(double b) {   // 2**b
#ifdef PHS_LIBM_POW
  return pow(2.0, b);
#else
  return exp2(b);
#endif
}
__device__ __forceinline__ double dexp(double a) { return exp(a); }
__device__ __forceinline__ double dlog(double a) { return log(a); }

struct Quad { double r1, r2; };
// quadraticMod.F90:17-74; *bad is set where the reference calls endrun
__device__ __forceinline__ Quad quadratic(double a, double b, double c, bool* bad) {
  Quad o;
  if (a == 0.0) { *bad = true; o.r1 = 0.0; o.r2 = 0.0; return o; }
  double root = b * b - 4.0 * a * c;
  if (root < 0.0) {
    if (-root < 3.0 * 2.220446049250313e-16) root = 0.0;
    else { *bad = true; o.r1 = 0.0; o.r2 = 0.0; return o; }
  }
  const double sq = sqrt(root);
  const double q = (b >= 0.0) ? -0.5 * (b + sq) : -0.5 * (b - sq);
  o.r1 = q / a;
  o.r2 = (q != 0.0) ? c / q : 1.e36;
  return o;
}

struct Weibull { double v, d; };   // plc and d1plc at one potential
// plc :5186-5187, d1plc :5218-5220 (same pow / exp2 feeds both)
__device__ __noinline__ Weibull weibull(double x, double psi50, double ck, bool want_d) {
  Weibull w;
  const double r = x / psi50;
  const double t = (r > 0.0) ? pw(r, ck) : pow(r, ck);
  const double e = pw2(-t);
  w.d = want_d ? (-ck * 0.6931471805599453 * e * t / x) : 0.0;   // log(2._r8)
  w.v = (e < 0.005) ? 0.0 : e;
  return w;
}
__device__ __forceinline__ double plc(double x, double psi50, double ck) { return weibull(x, psi50, ck, false).v; }

// everything about one patch that the hydraulics needs, held in registers
struct PhsPatch {
  double psi50[4], ck[4], kmax[4];
  double laisun, laisha, elai, esai, tsai, htop, fdry;
  double forc_rho, forc_pbot, cf;        // cf = forc_pbot/(rgas*1e-3*thm)*1e6
  double qsatl, qaf, gb_mol;
  double ksum;                           // sum_j k_soil_root(p,j)
  double ksmp;                           // sum_j k*smp
  double ksmpg;                          // sum_j k*(smp-grav2)
  double smpg_mean;                      // sum_j (smp-grav2) / nlevsoi  (k == 0 everywhere)
  const double* sk;                      // shared: k_soil_root [j*stride]
  const double* sg;                      // shared: grav2
  const double* ss;                      // shared: smp
  int stride;
  int* work;                             // per-pass work estimate of this patch (drives the cost bins of the next pass)
  __device__ __forceinline__ double K(int j) const { return sk[j * stride]; }
  __device__ __forceinline__ double G(int j) const { return sg[j * stride]; }
  __device__ __forceinline__ double S(int j) const { return ss[j * stride]; }
};

// getqflx :5128-5146 (havegs = .true.)
__device__ __forceinline__ void qflx_from_gs(const PhsPatch& P, double gs_sun, double gs_sha, double& qsun, double& qsha) {
  const double wtl = (P.elai + P.esai) * P.gb_mol;
  const double efpot = P.forc_rho * wtl * (P.qsatl - P.qaf);
  qsun = 0.0; qsha = 0.0;
  if ((efpot > 0.0) && (P.elai > 0.0)) {
    if (gs_sun > 0.0) {
      const double rpp = P.fdry / P.gb_mol * (P.laisun / (1.0 / P.gb_mol + 1.0 / gs_sun)) / P.elai;
      qsun = efpot * rpp / P.cf;
    }
    if (gs_sha > 0.0) {
      const double rpp = P.fdry / P.gb_mol * (P.laisha / (1.0 / P.gb_mol + 1.0 / gs_sha)) / P.elai;
      qsha = efpot * rpp / P.cf;
    }
  }
}
// getqflx :5148-5158 (havegs = .false.)
__device__ __forceinline__ void gs_from_qflx(const PhsPatch& P, double qsun, double qsha, double& gs_sun, double& gs_sha) {
  const double wtl = (P.elai + P.esai) * P.gb_mol;
  const double efpot = P.forc_rho * wtl * (P.qsatl - P.qaf);
  gs_sun = (qsun > 0.0) ? P.gb_mol * qsun * P.cf * P.elai / (efpot * P.fdry * P.laisun - qsun * P.cf * P.elai) : 0.0;
  gs_sha = (qsha > 0.0) ? P.gb_mol * qsha * P.cf * P.elai / (efpot * P.fdry * P.laisha - qsha * P.cf * P.elai) : 0.0;
}

// getvegwp :4979-5077.  x = {sun, sha, xyl, root}; returns soilflux.
__device__ __noinline__ double getvegwp(const PhsPatch& P, double* x, double gs_sun, double gs_sha) {
  double qsun, qsha;
  qflx_from_gs(P, gs_sun, gs_sha, qsun, qsha);
  const double grav1 = 1000.0 * P.htop;
  if (fabs(P.ksum) == 0.0) x[ROOT] = P.smpg_mean;
  else x[ROOT] = (P.ksmpg - qsun - qsha) / P.ksum;
  const double fr = plc(x[ROOT], P.psi50[ROOT], P.ck[ROOT]);
  if ((P.tsai > 0.0) && (fr > 0.0)) x[XYL] = x[ROOT] - grav1 - (qsun + qsha) / (fr * P.kmax[ROOT] / P.htop * P.tsai);
  else x[XYL] = x[ROOT] - grav1;
  const double fx = plc(x[XYL], P.psi50[XYL], P.ck[XYL]);
  if ((P.laisha > 0.0) && (fx > 0.0)) x[SHA] = x[XYL] - (qsha / (fx * P.kmax[XYL] * P.laisha));
  else x[SHA] = x[XYL];
  if ((P.laisun > 0.0) && (fx > 0.0)) x[SUN] = x[XYL] - (qsun / (fx * P.kmax[XYL] * P.laisun));
  else x[SUN] = x[XYL];
  double soilflux = 0.0;
  const double xr = x[ROOT];
  for (int j = 0; j < NLEVSOI; ++j) soilflux = soilflux + P.K(j) * (P.S(j) - xr - P.G(j));
  return soilflux;
}

struct Stress { double bsun, bsha; bool night; };

// calcstress :4490-4710.  On return x holds the potentials the reference leaves in x;
// *tran receives qflx_tran_veg when night (else untouched).
__device__ __noinline__ Stress calcstress(const PhsPatch& P, double* x, double gs_sun_in, double gs_sha_in, double* tran) {
  Stress out;
  out.night = (x[SUN] > 0.0);                         // night sentinel :4563-4568
  if (out.night) x[SUN] = x[SHA];
  double qsun, qsha;
  qflx_from_gs(P, gs_sun_in, gs_sha_in, qsun, qsha);
  const bool both = (P.laisun > tol_lai && P.laisha > tol_lai);
  const bool sha_only = (!both && P.laisha > tol_lai);
  bool flag;
  if ((P.laisun > tol_lai || P.laisha > tol_lai) && (qsun > 0.0 || qsha > 0.0)) {
    const double grav1 = P.htop * 1000.0;
    const double ls = P.laisun * P.kmax[SUN], lh = P.laisha * P.kmax[SHA];
    const double tk = P.tsai * P.kmax[XYL] / P.htop;
    flag = false;
    for (int iter = 1;; ++iter) {
      *P.work += 2;
      // segment conductance attenuation at x (shared by spacF :4951-4954 and spacA :4790-4799)
      const Weibull w1 = weibull(x[SUN], P.psi50[SUN], P.ck[SUN], true);
      const Weibull w2 = weibull(x[SHA], P.psi50[SHA], P.ck[SHA], true);
      const Weibull wx = weibull(x[XYL], P.psi50[XYL], P.ck[XYL], true);
      const Weibull wr = weibull(x[ROOT], P.psi50[ROOT], P.ck[ROOT], true);
      // spacF :4957-4972
      double f0 = qsun * w1.v - ls * wx.v * (x[XYL] - x[SUN]);
      double f1 = qsha * w2.v - lh * wx.v * (x[XYL] - x[SHA]);
      const double f2 = ls * wx.v * (x[XYL] - x[SUN]) + lh * wx.v * (x[XYL] - x[SHA]) - tk * wr.v * (x[ROOT] - x[XYL] - grav1);
      double s1 = 0.0;
      const double xr = x[ROOT];
      for (int j = 0; j < NLEVSOI; ++j) s1 += P.K(j) * (xr + P.G(j));
      const double f3 = tk * wr.v * (x[ROOT] - x[XYL] - grav1) + s1 - P.ksmp;
      if (P.laisha < tol_lai) { const double t = f0; f0 = f1; f1 = t; }
      if (sqrt(f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3) < 1.e-6 * (qsun + qsha)) { flag = false; break; }
      if (iter > 50) { flag = false; break; }
      // spacA :4802-4822 (only the structurally non-zero entries)
      double a11 = -ls * wx.v - qsun * w1.d;
      double a13 = ls * wx.d * (x[XYL] - x[SUN]) + ls * wx.v;
      double a22 = -lh * wx.v - qsha * w2.d;
      double a23 = lh * wx.d * (x[XYL] - x[SHA]) + lh * wx.v;
      double a31 = ls * wx.v;
      double a32 = lh * wx.v;
      const double a33 = -ls * wx.d * (x[XYL] - x[SUN]) - ls * wx.v - lh * wx.d * (x[XYL] - x[SHA]) - lh * wx.v - tk * wr.v;
      const double a34 = tk * wr.d * (x[ROOT] - x[XYL] - grav1) + tk * wr.v;
      const double a43 = tk * wr.v;
      const double a44 = -tk * wr.v - tk * wr.d * (x[ROOT] - x[XYL] - grav1) - P.ksum;
      double d0, d1, d2, d3;
      if (both) {                                      // :4831-4858
        const double determ = a44 * a22 * a33 * a11 - a44 * a22 * a31 * a13 - a44 * a32 * a23 * a11 - a43 * a11 * a22 * a34;
        if (fabs(determ) <= 1.e-50) { flag = true; break; }
        const double L = 1.0 / determ;
        const double i11 = L * a44 * a22 * a33 - L * a44 * a32 * a23 - L * a43 * a22 * a34;
        const double i21 = L * a23 * a44 * a31;
        const double i31 = -L * a44 * a22 * a31;
        const double i41 = L * a43 * a22 * a31;
        const double i12 = L * a13 * a44 * a32;
        const double i22 = L * a44 * a33 * a11 - L * a44 * a31 * a13 - L * a43 * a11 * a34;
        const double i32 = -L * a11 * a44 * a32;
        const double i42 = L * a43 * a11 * a32;
        const double i13 = -L * a13 * a22 * a44;
        const double i23 = -L * a23 * a11 * a44;
        const double i33 = L * a22 * a11 * a44;
        const double i43 = -L * a43 * a11 * a22;
        const double i14 = L * a13 * a34 * a22;
        const double i24 = L * a23 * a34 * a11;
        const double i34 = -L * a34 * a11 * a22;
        const double i44 = L * a22 * a33 * a11 - L * a22 * a31 * a13 - L * a32 * a23 * a11;
        d0 = ((0.0 + i11 * f0) + i12 * f1 + i13 * f2) + i14 * f3;    // matmul(A,f), k ascending
        d1 = ((0.0 + i21 * f0) + i22 * f1 + i23 * f2) + i24 * f3;
        d2 = ((0.0 + i31 * f0) + i32 * f1 + i33 * f2) + i34 * f3;
        d3 = ((0.0 + i41 * f0) + i42 * f1 + i43 * f2) + i44 * f3;
      } else {                                         // :4863-4887, 3x3 in rows/cols 2..4
        if (P.laisha <= tol_lai) { a22 = a11; a32 = a31; a23 = a13; }
        const double determ = a22 * a33 * a44 - a34 * a22 * a43 - a23 * a32 * a44;
        if (fabs(determ) <= 1.e-50) { flag = true; break; }
        const double rd = 1.0 / determ;
        const double i22 = rd * (a33 * a44 - a34 * a43), i23 = rd * (-a23 * a44), i24 = rd * (a34 * a23);
        const double i32 = rd * (-a32 * a44), i33 = rd * (a22 * a44), i34 = rd * (-a34 * a22);
        const double i42 = rd * (a32 * a43), i43 = rd * (-a22 * a43), i44 = rd * (a22 * a33 - a23 * a32);
        d0 = 0.0;
        d1 = (0.0 + i22 * f1) + i23 * f2 + i24 * f3;
        d2 = (0.0 + i32 * f1) + i33 * f2 + i34 * f3;
        d3 = (0.0 + i42 * f1) + i43 * f2 + i44 * f3;
      }
      const double mx = fmax(fmax(fabs(d0), fabs(d1)), fmax(fabs(d2), fabs(d3)));
      if (mx > 50000.0) { d0 = 50000.0 * d0 / mx; d1 = 50000.0 * d1 / mx; d2 = 50000.0 * d2 / mx; d3 = 50000.0 * d3 / mx; }
      if (both) {
        x[SUN] += d0; x[SHA] += d1; x[XYL] += d2; x[ROOT] += d3;
      } else if (sha_only) {
        x[SUN] += d0; x[SHA] += d1; x[XYL] += d2; x[ROOT] += d3;
        x[SUN] = x[XYL];
      } else {
        x[XYL] += d2; x[ROOT] += d3;
        x[SUN] += d1;
        x[SHA] = x[XYL];
      }
      if (sqrt(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) < 1.e-9) break;
      if (x[XYL] > x[ROOT]) x[XYL] = x[ROOT];
      if (x[SUN] > x[XYL]) x[SUN] = x[XYL];
      if (x[SHA] > x[XYL]) x[SHA] = x[XYL];
    }
  } else {
    flag = true;
  }
  if (flag) {
    (void)getvegwp(P, x, gs_sun_in, gs_sha_in);
    out.bsun = plc(x[SUN], P.psi50[SUN], P.ck[SUN]);
    out.bsha = plc(x[SHA], P.psi50[SHA], P.ck[SHA]);
  } else {
    const double p1 = plc(x[SUN], P.psi50[SUN], P.ck[SUN]);
    const double p2 = plc(x[SHA], P.psi50[SHA], P.ck[SHA]);
    double g1, g2;
    gs_from_qflx(P, qsun * p1, qsha * p2, g1, g2);
    out.bsun = (qsun > 0.0) ? g1 / gs_sun_in : p1;
    out.bsha = (qsha > 0.0) ? g2 / gs_sha_in : p2;
  }
  if (out.bsun < 0.01) out.bsun = 0.0;
  if (out.bsha < 0.01) out.bsha = 0.0;
  if (out.night) {
    double sf = getvegwp(P, x, out.bsun * gs_sun_in, out.bsha * gs_sha_in);
    if (sf < 0.0) sf = 0.0;
    *tran = sf;
  }
  return out;
}

// leaf biochemistry of one patch for the current t_veg (PhotosynthesisHydraulicStress :3118-3469)
struct Leaf {
  bool c3, medlyn;
  double vcmax[2], tpu[2], kp[2], lmr[2], je[2], par[2];
  double cp, kc, ko, qe;
  double theta_cj, theta_ip;
  double medint, medslope, bbb, mbb;
  double cair, oair, rh_can;
};
struct CiOut {   // what ci_func_PHS leaves in the photosyns arrays
  double ac[2], aj[2], ap[2], ag[2], an[2];
};

// ci_func_PHS :4227-4486 minus the bflag/calcstress prologue (done by the caller).
__device__ __noinline__ void ci_func(const PhsPatch& P, const Leaf& L, double cisun, double cisha, double bsun, double bsha,
                                     double& fsun, double& fsha, double& gs_sun, double& gs_sha, CiOut& o, bool* bad) {
  const double ci[2] = {cisun, cisha};
  const double b[2] = {bsun, bsha};
  *P.work += 1;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (L.c3) {
      o.ac[s] = b[s] * L.vcmax[s] * fmax(ci[s] - L.cp, 0.0) / (ci[s] + L.kc * (1.0 + L.oair / L.ko));
      o.aj[s] = L.je[s] * fmax(ci[s] - L.cp, 0.0) / (4.0 * ci[s] + 8.0 * L.cp);
      o.ap[s] = 3.0 * L.tpu[s];
    } else {
      o.ac[s] = b[s] * L.vcmax[s];
      o.aj[s] = L.qe * L.par[s] * 4.6;
      o.ap[s] = L.kp[s] * fmax(ci[s], 0.0) / P.forc_pbot;
    }
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {                      // :4343-4371
    Quad q = quadratic(L.theta_cj, -(o.ac[s] + o.aj[s]), o.ac[s] * o.aj[s], bad);
    const double ai = fmin(q.r1, q.r2);
    q = quadratic(L.theta_ip, -(ai + o.ap[s]), ai * o.ap[s], bad);
    o.ag[s] = fmax(0.0, fmin(q.r1, q.r2));
    o.an[s] = o.ag[s] - b[s] * L.lmr[s];
  }
  double* gs[2] = {&gs_sun, &gs_sha};
  double* fv[2] = {&fsun, &fsha};
#pragma unroll
  for (int s = 0; s < 2; ++s) {                      // :4373-4396
    if (o.an[s] < 0.0) {
      *gs[s] = fmax(b[s] * (L.medlyn ? L.medint : L.bbb), 1.0);
      *fv[s] = 0.0;
    }
  }
  if ((o.an[0] < 0.0) && (o.an[1] < 0.0)) return;
#pragma unroll
  for (int s = 0; s < 2; ++s) {                      // :4405-4484
    if (o.an[s] >= 0.0) {
      double cs = L.cair - 1.4 / P.gb_mol * o.an[s] * P.forc_pbot;
      cs = fmax(cs, max_cs);
      if (L.medlyn) {
        const double term = 1.6 * o.an[s] / (cs / P.forc_pbot * 1.e06);
        const double bq = -(2.0 * (L.medint * 1.e-06 + term) + ((L.medslope * term) * (L.medslope * term)) / (P.gb_mol * 1.e-06 * L.rh_can));
        const double cq = L.medint * L.medint * 1.e-12 + (2.0 * L.medint * 1.e-06 + term * (1.0 - L.medslope * L.medslope / L.rh_can)) * term;
        const Quad q = quadratic(1.0, bq, cq, bad);
        *gs[s] = fmax(q.r1, q.r2) * 1.e06;
      } else {
        const double gmin = fmax(b[s] * L.bbb, 1.0);
        const double bq = cs * (P.gb_mol - gmin) - L.mbb * o.an[s] * P.forc_pbot;
        const double cq = -P.gb_mol * (cs * gmin + L.mbb * o.an[s] * P.forc_pbot * L.rh_can);
        const Quad q = quadratic(cs, bq, cq, bad);
        *gs[s] = fmax(q.r1, q.r2);
      }
      if (*gs[s] > 0.0) *fv[s] = ci[s] - L.cair + o.an[s] * P.forc_pbot * (1.4 * *gs[s] + 1.6 * P.gb_mol) / (P.gb_mol * *gs[s]);
      else *fv[s] = ci[s] - L.cair;
    }
  }
}

// brent_PHS :4068-4223
__device__ __noinline__ void brent(const PhsPatch& P, const Leaf& L, double& xsun, double x1sun, double x2sun, double f1sun,
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
__device__ __noinline__ HybridOut hybrid(const PhsPatch& P, const Leaf& L, const double* vegwp_in, double ci0, CiOut& o,
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
