// phs.cuh — per-thread (one patch per thread) plant-hydraulic-stress photosynthesis.
//
// Reference: src/biogeophys/PhotosynthesisMod.F90
//   hybrid_PHS :3815-4064   brent_PHS :4068-4223   ci_func_PHS :4227-4486
//   calcstress :4490-4710   spacA :4715-4893       spacF :4898-4976
//   getvegwp :4979-5077     getqflx :5080-5164     plc :5167   d1plc :5199
//   quadratic  src/utils/quadraticMod.F90:17-74
//
// B200 mapping (not the reference's call structure):
//  * the nested, data-dependent loops of the reference (hybrid -> calcstress Newton -> ci_func -> secant -> brent)
//    are restated as two kinds of RESUMABLE TASKS - a calcstress solve (newton_begin / newton_step / newton_finish)
//    and one outer pass of the ci solve (ci_task_begin / ci_step / ci_task_end) - so that the kernels in canopy.cu can
//    advance 32 independent tasks per warp one step at a time and refill lanes whose task has ended;
//  * the functions are templates over the "patch view" they read (registers + shared memory inside the Newton loop,
//    the patch's global-memory record in prologues / epilogues): no copies through local memory;
//  * spacF and spacA are evaluated at the same potentials, so the vulnerability curve and its derivative are
//    evaluated once per Newton step (Weibull: one pow + one exp2 per segment) and shared by both; the level sums that
//    do not depend on the root potential (sum k, sum k*smp, sum k*(smp-grav2)) are formed once per call;
//  * fmad is off (the reference is built with -ffp-contract=off); the arithmetic and its order are the reference's,
//    tests/host/phs_tasks_check.cu checks the task formulation bit for bit against the nested-loop one on the CPU.
#pragma once
#include "common.cuh"

// Every function here is host+device: the device build is the product; the host build exists only so that
// tests/host/phs_tasks_check.cu can run the task state machines against the direct formulation on the CPU.
#define PHS_FN static __host__ __device__ __noinline__
#define PHS_INL __host__ __device__ __forceinline__

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
PHS_INL double pw(double a, double b) {
#ifdef PHS_LIBM_POW
  return pow(a, b);
#else
  return exp2(b * log2(a));
#endif
}
PHS_INL double pw2(double b) {   // 2**b
#ifdef PHS_LIBM_POW
  return pow(2.0, b);
#else
  return exp2(b);
#endif
}
PHS_INL double dexp(double a) { return exp(a); }
PHS_INL double dlog(double a) { return log(a); }

struct Quad { double r1, r2; };
// quadraticMod.F90:17-74; *bad is set where the reference calls endrun
PHS_INL Quad quadratic(double a, double b, double c, bool* bad) {
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
// x**ck for x <= 0 (never on a physical state: potentials and psi50 are negative); kept out of line
PHS_FN double pow_nonpos(double r, double ck) { return pow(r, ck); }
// plc :5186-5187, d1plc :5218-5220 (same pow / exp2 feeds both)
PHS_INL Weibull weibull_inl(double x, double psi50, double ck, bool want_d) {
  Weibull w;
  const double r = x / psi50;
  const double t = (r > 0.0) ? pw(r, ck) : pow_nonpos(r, ck);
  const double e = pw2(-t);
  w.d = want_d ? (-ck * 0.6931471805599453 * e * t / x) : 0.0;   // log(2._r8)
  w.v = (e < 0.005) ? 0.0 : e;
  return w;
}
PHS_FN Weibull weibull(double x, double psi50, double ck, bool want_d) { return weibull_inl(x, psi50, ck, want_d); }
PHS_INL double plc(double x, double psi50, double ck) { return weibull(x, psi50, ck, false).v; }

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
  PHS_INL double K(int j) const { return sk[j * stride]; }
  PHS_INL double G(int j) const { return sg[j * stride]; }
  PHS_INL double S(int j) const { return ss[j * stride]; }
};

// what one Newton iteration reads of a patch (kept in registers / shared memory by the calcstress task kernel)
struct NewtonCtx {
  double psi50[4], ck[4];
  double laisha, ksum, ksmp;
  const double* sk;
  const double* sg;
  int stride;
  PHS_INL double K(int j) const { return sk[j * stride]; }
  PHS_INL double G(int j) const { return sg[j * stride]; }
};

// getqflx :5128-5146 (havegs = .true.)
template <class PT>
PHS_INL void qflx_from_gs(const PT& P, double gs_sun, double gs_sha, double& qsun, double& qsha) {
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
template <class PT>
PHS_INL void gs_from_qflx(const PT& P, double qsun, double qsha, double& gs_sun, double& gs_sha) {
  const double wtl = (P.elai + P.esai) * P.gb_mol;
  const double efpot = P.forc_rho * wtl * (P.qsatl - P.qaf);
  gs_sun = (qsun > 0.0) ? P.gb_mol * qsun * P.cf * P.elai / (efpot * P.fdry * P.laisun - qsun * P.cf * P.elai) : 0.0;
  gs_sha = (qsha > 0.0) ? P.gb_mol * qsha * P.cf * P.elai / (efpot * P.fdry * P.laisha - qsha * P.cf * P.elai) : 0.0;
}

// getvegwp :4979-5077.  x = {sun, sha, xyl, root}; returns soilflux.
template <class PT>
PHS_FN double getvegwp(const PT& P, double* x, double gs_sun, double gs_sha) {
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

// calcstress :4490-4710 as a resumable solve: newton_begin / newton_step (one Newton iteration of the 4x4
// plant-water-potential system) / newton_finish.  One lane carries one solve; the warp scheduler in canopy.cu
// refills lanes whose solve has ended, so that a solve that needs 50 iterations does not hold 31 others hostage.
struct Newton {
  double x[4];            // {sun, sha, xyl, root} potentials
  double qsun, qsha;      // getqflx at the input conductances
  double ls, lh, tk, grav1;
  int iter;
  bool both, sha_only, night, flag;
};

// returns true when Newton iterations are needed (:4573-4575)
template <class PT>
PHS_INL bool newton_begin(Newton& N, const PT& P, const double* xin, double gs_sun_in, double gs_sha_in) {
#pragma unroll
  for (int i = 0; i < 4; ++i) N.x[i] = xin[i];
  N.night = (N.x[SUN] > 0.0);                         // night sentinel :4563-4568
  if (N.night) N.x[SUN] = N.x[SHA];
  qflx_from_gs(P, gs_sun_in, gs_sha_in, N.qsun, N.qsha);
  N.both = (P.laisun > tol_lai && P.laisha > tol_lai);
  N.sha_only = (!N.both && P.laisha > tol_lai);
  N.grav1 = P.htop * 1000.0;
  N.ls = P.laisun * P.kmax[SUN];
  N.lh = P.laisha * P.kmax[SHA];
  N.tk = P.tsai * P.kmax[XYL] / P.htop;
  N.iter = 0;
  if ((P.laisun > tol_lai || P.laisha > tol_lai) && (N.qsun > 0.0 || N.qsha > 0.0)) { N.flag = false; return true; }
  N.flag = true;
  return false;
}

// one iteration of the loop :4579-4640; returns true while the iteration continues
// The Weibull curves of the four segments at the current potentials: one after the other in one lane (default), or
// one segment per lane of a quad (canopy.cu, small queues, where the latency of a single solve is what matters).
struct WeibullSerial {
  PHS_INL void operator()(const double* x, const double* psi50, const double* ck, Weibull* w) const {
    w[SUN] = weibull(x[SUN], psi50[SUN], ck[SUN], true);
    w[SHA] = weibull(x[SHA], psi50[SHA], ck[SHA], true);
    w[XYL] = weibull(x[XYL], psi50[XYL], ck[XYL], true);
    w[ROOT] = weibull(x[ROOT], psi50[ROOT], ck[ROOT], true);
  }
};

template <class PT, class WB = WeibullSerial>
PHS_INL bool newton_step(Newton& N, const PT& P, const WB wb = WB()) {
  double* x = N.x;
  const double qsun = N.qsun, qsha = N.qsha, ls = N.ls, lh = N.lh, tk = N.tk, grav1 = N.grav1;
  const int iter = ++N.iter;
  // segment conductance attenuation at x (shared by spacF :4951-4954 and spacA :4790-4799)
  Weibull w4[4];
  wb(x, P.psi50, P.ck, w4);
  const Weibull w1 = w4[SUN], w2 = w4[SHA], wx = w4[XYL], wr = w4[ROOT];
  // spacF :4957-4972
  double f0 = qsun * w1.v - ls * wx.v * (x[XYL] - x[SUN]);
  double f1 = qsha * w2.v - lh * wx.v * (x[XYL] - x[SHA]);
  const double f2 = ls * wx.v * (x[XYL] - x[SUN]) + lh * wx.v * (x[XYL] - x[SHA]) - tk * wr.v * (x[ROOT] - x[XYL] - grav1);
  double s1 = 0.0;
  const double xr = x[ROOT];
  for (int j = 0; j < NLEVSOI; ++j) s1 += P.K(j) * (xr + P.G(j));
  const double f3 = tk * wr.v * (x[ROOT] - x[XYL] - grav1) + s1 - P.ksmp;
  if (P.laisha < tol_lai) { const double t = f0; f0 = f1; f1 = t; }
  if (sqrt(f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3) < 1.e-6 * (qsun + qsha)) { N.flag = false; return false; }
  if (iter > 50) { N.flag = false; return false; }
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
  if (N.both) {                                      // :4831-4858
    const double determ = a44 * a22 * a33 * a11 - a44 * a22 * a31 * a13 - a44 * a32 * a23 * a11 - a43 * a11 * a22 * a34;
    if (fabs(determ) <= 1.e-50) { N.flag = true; return false; }
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
    if (fabs(determ) <= 1.e-50) { N.flag = true; return false; }
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
  if (N.both) {
    x[SUN] += d0; x[SHA] += d1; x[XYL] += d2; x[ROOT] += d3;
  } else if (N.sha_only) {
    x[SUN] += d0; x[SHA] += d1; x[XYL] += d2; x[ROOT] += d3;
    x[SUN] = x[XYL];
  } else {
    x[XYL] += d2; x[ROOT] += d3;
    x[SUN] += d1;
    x[SHA] = x[XYL];
  }
  if (sqrt(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) < 1.e-9) return false;
  if (x[XYL] > x[ROOT]) x[XYL] = x[ROOT];
  if (x[SUN] > x[XYL]) x[SUN] = x[XYL];
  if (x[SHA] > x[XYL]) x[SHA] = x[XYL];
  return true;
}

// :4642-4708.  On return N.x holds the potentials the reference leaves in x; *tran receives qflx_tran_veg when
// night (else untouched).
template <class PT>
PHS_INL Stress newton_finish(Newton& N, const PT& P, double gs_sun_in, double gs_sha_in, double* tran) {
  Stress out;
  double* x = N.x;
  out.night = N.night;
  if (N.flag) {
    (void)getvegwp(P, x, gs_sun_in, gs_sha_in);
    out.bsun = plc(x[SUN], P.psi50[SUN], P.ck[SUN]);
    out.bsha = plc(x[SHA], P.psi50[SHA], P.ck[SHA]);
  } else {
    const double p1 = plc(x[SUN], P.psi50[SUN], P.ck[SUN]);
    const double p2 = plc(x[SHA], P.psi50[SHA], P.ck[SHA]);
    double g1, g2;
    gs_from_qflx(P, N.qsun * p1, N.qsha * p2, g1, g2);
    out.bsun = (N.qsun > 0.0) ? g1 / gs_sun_in : p1;
    out.bsha = (N.qsha > 0.0) ? g2 / gs_sha_in : p2;
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

// calcstress in one go (used by the direct formulation the host check compares against)
PHS_FN Stress calcstress(const PhsPatch& P, double* x, double gs_sun_in, double gs_sha_in, double* tran) {
  Newton N;
  if (newton_begin(N, P, x, gs_sun_in, gs_sha_in)) {
    while (newton_step(N, P)) {}
  }
  const Stress out = newton_finish(N, P, gs_sun_in, gs_sha_in, tran);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = N.x[i];
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
template <class PT>
PHS_INL void ci_func(const PT& P, const Leaf& L, double cisun, double cisha, double bsun, double bsha,
                                     double& fsun, double& fsha, double& gs_sun, double& gs_sha, CiOut& o, bool* bad) {
  const double ci[2] = {cisun, cisha};
  const double b[2] = {bsun, bsha};
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

// hybrid_PHS :3815-4064 and brent_PHS :4068-4223 as a resumable solve.  The reference nests
//   hybrid (<= 4 outer passes) { calcstress; ci_func x2; secant (<= 4 ci_func) { brent (<= 20 ci_func) } }
// and every patch takes its own path through it.  Here one outer pass is a "ci task": a lane state (CiLane) that
// is advanced by ci_step = ONE ci_func evaluation followed by the control logic of whichever loop the lane is in.
// All lanes of a warp therefore execute the same expensive code (ci_func) whatever their position in the nest.
// What survives from one outer pass to the next (HybridCarry) lives in the patch record between tasks.
enum CiState { CI_F0 = 0, CI_F1 = 1, CI_SEC = 2, CI_FINAL = 3, CI_BRENT = 4 };

struct HybridCarry {
  double x1sun, x1sha;       // secant iterate carried across outer passes
  double gs0sun, gs0sha;     // unstressed conductances handed to calcstress
  double bsun, bsha;         // stress factors of the current outer pass
  double b0sun, b0sha;       // ... of the previous one
  int iter1;
};
PHS_INL void hybrid_carry_init(HybridCarry& H, double ci0) {       // :3893-3915
  H.x1sun = ci0; H.x1sha = ci0; H.gs0sun = 0.0; H.gs0sha = 0.0; H.bsun = 1.0; H.bsha = 1.0; H.b0sun = -1.0; H.b0sha = -1.0;
  H.iter1 = 1;
}

struct Brent { double a[2], b[2], c[2], d[2], e[2], fa[2], fb[2], fc[2], tol; };

struct CiLane {
  double x0sun, x0sha, x1sun, x1sha, f0sun, f0sha, f1sun, f1sha;
  double gs_sun, gs_sha, bsun, bsha;
  double dxsun, dxsha, tolsun, tolsha, minf, minxsun, minxsha;
  double dbsun, dbsha;
  double cs, ch;             // where ci_func is evaluated next
  int st, iter2, biter;
  CiOut o;
};

// top of one outer pass :3925-3950 (calcstress, when due, has already updated H.bsun / H.bsha)
PHS_INL void ci_task_begin(CiLane& C, HybridCarry& H) {
  C.x1sun = H.x1sun; C.x1sha = H.x1sha;
  C.x0sun = fmax(0.1, C.x1sun); C.x1sun = 0.99 * C.x1sun;
  C.x0sha = fmax(0.1, C.x1sha); C.x1sha = 0.99 * C.x1sha;
  C.tolsun = fabs(C.x1sun) * 1.e-2; C.tolsha = fabs(C.x1sha) * 1.e-2;
  C.bsun = H.bsun; C.bsha = H.bsha;
  C.dbsun = H.b0sun - H.bsun; C.dbsha = H.b0sha - H.bsha;
  H.b0sun = H.bsun; H.b0sha = H.bsha;
  C.iter2 = 0; C.biter = 0;
  C.gs_sun = 0.0; C.gs_sha = 0.0;
  C.st = CI_F0; C.cs = C.x0sun; C.ch = C.x0sha;
}

// one ci_func evaluation + the control logic that follows it; returns true while the outer pass continues
template <class PT>
PHS_INL bool ci_step(CiLane& C, Brent& B, const PT& P, const Leaf& L, bool* bad, bool* notbracketed) {
  double fs, fh;
  ci_func(P, L, C.cs, C.ch, C.bsun, C.bsha, fs, fh, C.gs_sun, C.gs_sha, C.o, bad);
  bool top = false;
  switch (C.st) {
    case CI_F0:
      C.f0sun = fs; C.f0sha = fh;
      C.st = CI_F1; C.cs = C.x1sun; C.ch = C.x1sha;
      return true;
    case CI_F1:
      C.f1sun = fs; C.f1sha = fh;
      top = true;
      break;
    case CI_SEC:                                       // :3975-4030
      C.f1sun = fs; C.f1sha = fh;
      if ((fabs(C.dxsun) < C.tolsun) && (fabs(C.dxsha) < C.tolsha)) { C.x0sun = C.x1sun; C.x0sha = C.x1sha; return false; }
      if (C.iter2 == 1 || fabs(C.f1sun + C.f1sha) < C.minf) { C.minf = fabs(C.f1sun + C.f1sha); C.minxsun = C.x1sun; C.minxsha = C.x1sha; }
      if ((fabs(C.f1sun) < 1.e-4) && (fabs(C.f1sha) < 1.e-4)) return false;
      if ((C.f1sun * C.f0sun < 0.0) && (C.f1sha * C.f0sha < 0.0)) {      // brent_PHS :4125-4141
        B.a[0] = C.x0sun; B.a[1] = C.x0sha; B.b[0] = C.x1sun; B.b[1] = C.x1sha;
        B.fa[0] = C.f0sun; B.fa[1] = C.f0sha; B.fb[0] = C.f1sun; B.fb[1] = C.f1sha;
        B.tol = C.tolsun;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          B.d[s] = 0.0; B.e[s] = 0.0;
          if ((B.fa[s] > 0.0 && B.fb[s] > 0.0) || (B.fa[s] < 0.0 && B.fb[s] < 0.0)) *notbracketed = true;
          B.c[s] = B.b[s]; B.fc[s] = B.fb[s];
        }
        C.biter = 0;
        break;
      }
      if (C.iter2 > 3) {
        C.x1sun = C.minxsun; C.x1sha = C.minxsha;
        C.st = CI_FINAL; C.cs = C.x1sun; C.ch = C.x1sha;
        return true;
      }
      top = true;
      break;
    case CI_FINAL:
      C.f1sun = fs; C.f1sha = fh;
      return false;
    default:                                           // CI_BRENT :4204-4217
      B.fb[0] = fs; B.fb[1] = fh;
      if (((fs == 0.0) && (fh == 0.0)) || C.biter >= 20) { C.x0sun = B.b[0]; C.x0sha = B.b[1]; return false; }
      break;
  }
  if (top) {                                           // :3953-3973
    if ((fabs(C.f0sun) < 1.e-4) && (fabs(C.f0sha) < 1.e-4)) { C.x1sun = C.x0sun; C.x1sha = C.x0sha; return false; }
    if ((fabs(C.f1sun) < 1.e-4) && (fabs(C.f1sha) < 1.e-4)) return false;
    C.iter2 = C.iter2 + 1;
    C.dxsun = ((C.f1sun - C.f0sun) == 0.0) ? 0.5 * (C.x1sun + C.x0sun) - C.x1sun : -C.f1sun * (C.x1sun - C.x0sun) / (C.f1sun - C.f0sun);
    C.dxsha = ((C.f1sha - C.f0sha) == 0.0) ? 0.5 * (C.x1sha + C.x0sha) - C.x1sha : -C.f1sha * (C.x1sha - C.x0sha) / (C.f1sha - C.f0sha);
    C.x0sun = C.x1sun; C.x1sun = C.x1sun + C.dxsun;
    C.x0sha = C.x1sha; C.x1sha = C.x1sha + C.dxsha;
    C.st = CI_SEC; C.cs = C.x1sun; C.ch = C.x1sha;
    return true;
  }
  // one pass of the brent_PHS loop up to its ci_func call :4143-4203
  ++C.biter;
  double tol1[2], xm[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if ((B.fb[s] > 0.0 && B.fc[s] > 0.0) || (B.fb[s] < 0.0 && B.fc[s] < 0.0)) { B.c[s] = B.a[s]; B.fc[s] = B.fa[s]; B.d[s] = B.b[s] - B.a[s]; B.e[s] = B.d[s]; }
    if (fabs(B.fc[s]) < fabs(B.fb[s])) { B.a[s] = B.b[s]; B.b[s] = B.c[s]; B.c[s] = B.a[s]; B.fa[s] = B.fb[s]; B.fb[s] = B.fc[s]; B.fc[s] = B.fa[s]; }
    tol1[s] = 2.0 * 1.e-4 * fabs(B.b[s]) + 0.5 * B.tol;
    xm[s] = 0.5 * (B.c[s] - B.b[s]);
  }
  if ((fabs(xm[0]) <= tol1[0] || B.fb[0] == 0.0) && (fabs(xm[1]) <= tol1[1] || B.fb[1] == 0.0)) { C.x0sun = B.b[0]; C.x0sha = B.b[1]; return false; }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (fabs(B.e[s]) >= tol1[s] && fabs(B.fa[s]) > fabs(B.fb[s])) {
      const double sv = B.fb[s] / B.fa[s];
      double pv, qv;
      if (B.a[s] == B.c[s]) {
        pv = 2.0 * xm[s] * sv;
        qv = 1.0 - sv;
      } else {
        qv = B.fa[s] / B.fc[s];
        const double rv = B.fb[s] / B.fc[s];
        pv = sv * (2.0 * xm[s] * qv * (qv - rv) - (B.b[s] - B.a[s]) * (rv - 1.0));
        qv = (qv - 1.0) * (rv - 1.0) * (sv - 1.0);
      }
      if (pv > 0.0) qv = -qv;
      pv = fabs(pv);
      if (2.0 * pv < fmin(3.0 * xm[s] * qv - fabs(tol1[s] * qv), fabs(B.e[s] * qv))) { B.e[s] = B.d[s]; B.d[s] = pv / qv; }
      else { B.d[s] = xm[s]; B.e[s] = B.d[s]; }
    } else {
      B.d[s] = xm[s]; B.e[s] = B.d[s];
    }
    B.a[s] = B.b[s]; B.fa[s] = B.fb[s];
    if (fabs(B.d[s]) > tol1[s]) B.b[s] = B.b[s] + B.d[s];
    else B.b[s] = B.b[s] + copysign(tol1[s], xm[s]);
  }
  C.st = CI_BRENT; C.cs = B.b[0]; C.ch = B.b[1];
  return true;
}

// bottom of one outer pass :4034-4046; returns true when hybrid_PHS is finished
PHS_INL bool ci_task_end(const CiLane& C, HybridCarry& H) {
  H.x1sun = C.x1sun; H.x1sha = C.x1sha;
  if (C.bsun > 0.01) H.gs0sun = C.gs_sun / C.bsun;
  if (C.bsha > 0.01) H.gs0sha = C.gs_sha / C.bsha;
  const bool last = ((fabs(C.dbsun) < 1.e-2) && (fabs(C.dbsha) < 1.e-2)) || (H.iter1 > 3);
  H.iter1 = H.iter1 + 1;
  return last;
}

}  // namespace phs
