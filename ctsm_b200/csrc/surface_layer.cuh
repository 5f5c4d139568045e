// surface_layer.cuh — the surface-layer device functions shared by CanopyFluxes (canopy.cu) and the routines that run just
// before it (preflux.cu: BiogeophysPreFluxCalcs, CalculateSurfaceHumidity, BareGroundFluxes):
//   QSat                               QSatMod.F90:61-127
//   StabilityFunc1/2, FrictionVelocity FrictionVelocityMod.F90:754-1155 (one patch per call; the four-regime log laws are
//                                      shared between ustar / u10 and temp1 / temp2 / temp12m / temp22m)
//   the math wrappers of the uniform kernels and the physical constants (clm_varcon.F90, shr_const_mod.F90)
// Everything is internal to the including translation unit (anonymous namespace).
#pragma once
#include "phs.cuh"

namespace {

using phs::rgas;
using phs::tfrz;
using phs::spval;
// Transcendentals of the uniform kernels (close / fric / leaf / init / final).  Out of line by default: the inlined
// libdevice bodies were half of the 150 KB step kernel, and what bounds these kernels is instruction fetch, not the
// call overhead (-DSTEP_MATH_INLINE restores inlining).  Same arithmetic as phs::pw / dexp / dlog.
#ifdef STEP_MATH_INLINE
#define SM_FN __device__ __forceinline__
#else
#define SM_FN __device__ __noinline__
#endif
SM_FN double pw(double a, double b) { return phs::pw(a, b); }
SM_FN double pw2(double b) { return phs::pw2(b); }
SM_FN double dexp(double a) { return exp(a); }
SM_FN double dlog(double a) { return log(a); }
constexpr double rpi = 3.14159265358979323846;
constexpr double sb = 5.67e-8, cpair = 1.00464e3, hvap = 2.501e6, vkc = 0.4, grav = 9.80616;
constexpr double denice = 0.917e3, denh2o = 1.000e3, c_to_b = 2.0, tlsai_crit = 2.0, alpha_aero = 1.0;
constexpr double c_water = 4.188e3, c_dry_biomass = 1400.0, nu_param = 1.5e-5, cd1_param = 7.5;

__device__ __forceinline__ double pow4(double t) { const double t2 = t * t; return t2 * t2; }
__device__ __forceinline__ double pow3(double t) { return (t * t) * t; }

// QSatMod.F90:61-127
struct QS { double qs, es, qsdT; };
__device__ __noinline__ QS qsat(double T, double p, bool deriv) {
  QS o;
  const double td = fmin(100.0, fmax(-75.0, T - tfrz));
  double es;
  if (td >= 0.0)
    es = 6.11213476 + td * (0.444007856 + td * (0.143064234e-01 + td * (0.264461437e-03 + td * (0.305903558e-05 + td * (0.196237241e-07 + td * (0.892344772e-10 + td * (-0.373208410e-12 + td * 0.209339997e-15)))))));
  else
    es = 6.11123516 + td * (0.503109514 + td * (0.188369801e-01 + td * (0.420547422e-03 + td * (0.614396778e-05 + td * (0.602780717e-07 + td * (0.387940929e-09 + td * (0.149436277e-11 + td * 0.262655803e-14)))))));
  es = es * 100.0;
  const double vp = 1.0 / (p - 0.378 * es);
  const double vp1 = 0.622 * vp;
  o.qs = es * vp1;
  o.es = es;
  o.qsdT = 0.0;
  if (deriv) {
    double d;
    if (td >= 0.0)
      d = 0.444017302 + td * (0.286064092e-01 + td * (0.794683137e-03 + td * (0.121211669e-04 + td * (0.103354611e-06 + td * (0.404125005e-09 + td * (-0.788037859e-12 + td * (-0.114596802e-13 + td * 0.381294516e-16)))))));
    else
      d = 0.503277922 + td * (0.377289173e-01 + td * (0.126801703e-02 + td * (0.249468427e-04 + td * (0.313703411e-06 + td * (0.257180651e-08 + td * (0.133268878e-10 + td * (0.394116744e-13 + td * 0.498070196e-16)))))));
    d = d * 100.0;
    const double vp2 = vp1 * vp;
    o.qsdT = d * vp2 * p;
  }
  return o;
}

// FrictionVelocityMod.F90:1120-1155
__device__ __noinline__ double stab1(double zeta) {
  const double chik2 = sqrt(1.0 - 16.0 * zeta);
  const double chik = sqrt(chik2);
  return 2.0 * dlog((1.0 + chik) * 0.5) + dlog((1.0 + chik2) * 0.5) - 2.0 * atan(chik) + rpi * 0.5;
}
__device__ __noinline__ double stab2(double zeta) {
  const double chik2 = sqrt(1.0 - 16.0 * zeta);
  return 2.0 * dlog((1.0 + chik2) * 0.5);
}
// the four-regime log-law denominator shared by ustar / u10 (momentum) ...
__device__ __noinline__ double prof_m(double zldis, double zeta, double obu, double z0) {
  const double zetam = 1.574;
  if (zeta < -zetam)
    return dlog(-zetam * obu / z0) - stab1(-zetam) + stab1(z0 / obu) + 1.14 * (pw(-zeta, 0.333) - pw(zetam, 0.333));
  if (zeta < 0.0) return dlog(zldis / z0) - stab1(zeta) + stab1(z0 / obu);
  if (zeta <= 1.0) return dlog(zldis / z0) + 5.0 * zeta - 5.0 * z0 / obu;
  return dlog(obu / z0) + 5.0 - 5.0 * z0 / obu + (5.0 * dlog(zeta) + zeta - 1.0);
}
// ... and by temp1 / temp2 / temp12m / temp22m (scalars)
__device__ __noinline__ double prof_h(double zldis, double zeta, double obu, double z0) {
  const double zetat = 0.465;
  if (zeta < -zetat)
    return dlog(-zetat * obu / z0) - stab2(-zetat) + stab2(z0 / obu) + 0.8 * (pw(zetat, -0.333) - pw(-zeta, -0.333));
  if (zeta < 0.0) return dlog(zldis / z0) - stab2(zeta) + stab2(z0 / obu);
  if (zeta <= 1.0) return dlog(zldis / z0) + 5.0 * zeta - 5.0 * z0 / obu;
  return dlog(obu / z0) + 5.0 - 5.0 * z0 / obu + (5.0 * dlog(zeta) + zeta - 1.0);
}

struct FricOut { double ustar, temp1, temp2, temp12m, temp22m, fm, vds, u10_clm, u10; };
// FrictionVelocity :842-1113 for one patch
__device__ __noinline__ FricOut friction_velocity(double hgt_u, double hgt_t, double hgt_q, double displa, double z0m,
                                                     double z0h, double z0q, double obu, int iter, double ur, double um,
                                                     double fm_prev) {
  FricOut o;
  double zldis = hgt_u - displa;
  double zeta = zldis / obu;
  o.ustar = vkc * um / prof_m(zldis, zeta, obu, z0m);
  if (zeta < 0.0) o.vds = 2.e-3 * o.ustar * (1.0 + pw(300.0 / (-obu), 0.666));
  else o.vds = 2.e-3 * o.ustar;
  if (zldis - z0m <= 10.0) o.u10_clm = um;
  else o.u10_clm = um - (o.ustar / vkc * prof_m(zldis, zeta, obu, 10.0 + z0m));
  zldis = hgt_t - displa;
  zeta = zldis / obu;
  o.temp1 = vkc / prof_h(zldis, zeta, obu, z0h);
  if (hgt_q == hgt_t && z0q == z0h) {
    o.temp2 = o.temp1;
  } else {
    zldis = hgt_q - displa;
    zeta = zldis / obu;
    o.temp2 = vkc / prof_h(zldis, zeta, obu, z0q);
  }
  zldis = 2.0 + z0h;
  zeta = zldis / obu;
  o.temp12m = vkc / prof_h(zldis, zeta, obu, z0h);
  if (z0q == z0h) {
    o.temp22m = o.temp12m;
  } else {
    zldis = 2.0 + z0q;
    zeta = zldis / obu;
    o.temp22m = vkc / prof_h(zldis, zeta, obu, z0q);
  }
  zldis = hgt_u - displa;
  zeta = zldis / obu;
  double fmnew;
  if (fmin(zeta, 1.0) < 0.0) {
    const double t1 = pw(1.0 - 16.0 * fmin(zeta, 1.0), 0.25);
    const double t2 = dlog((1.0 + t1 * t1) / 2.0);
    const double t3 = dlog((1.0 + t1) / 2.0);
    fmnew = 2.0 * t3 + t2 - 2.0 * atan(t1) + 1.5707963;
  } else {
    fmnew = -5.0 * fmin(zeta, 1.0);
  }
  o.fm = (iter == 1) ? fmnew : 0.5 * (fm_prev + fmnew);
  double zeta10 = fmin(10.0 / obu, 1.0);
  if (zeta == 0.0) zeta10 = 0.0;
  double fm10;
  if (zeta10 < 0.0) {
    const double t1 = pw(1.0 - 16.0 * zeta10, 0.25);
    const double t2 = dlog((1.0 + t1 * t1) / 2.0);
    const double t3 = dlog((1.0 + t1) / 2.0);
    fm10 = 2.0 * t3 + t2 - 2.0 * atan(t1) + 1.5707963;
  } else {
    fm10 = -5.0 * zeta10;
  }
  const double t4 = dlog(fmax(1.0, hgt_u / 10.0));
  o.u10 = ur - o.ustar / vkc * (t4 - o.fm + fm10);
  return o;
}


}  // namespace
