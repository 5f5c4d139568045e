/* oracle_pert.h — TEST INFRASTRUCTURE.  Conditioning probe for the CanopyFluxes / PHS oracle.
 *
 * Built with -DORACLE_LIBM_PERTURB (liboracle_pert.so) every pow / exp / log / log10 / atan result is moved by ONE unit in
 * the last place, up or down depending on a bit of the result (deterministic).  That is the size of the disagreement any
 * two correct libm implementations may have (glibc on the CPU, libdevice on the GPU: both claim <= 1-2 ulp), so the
 * difference between the oracle and its perturbed twin MEASURES, patch by patch, how far legitimate last-bit noise is
 * amplified by the ITERATION loop.  Parity tests use it to decide which patches can be held to 1e-10 at all: a patch whose
 * own result moves by more than the tolerance under a 1-ulp libm change has no 1e-10 answer to compare against
 * (tests/test_gpu_baseline_configs.py).  sqrt is correctly rounded everywhere and is left alone. */
#ifndef ORACLE_PERT_H
#define ORACLE_PERT_H
#ifdef ORACLE_LIBM_PERTURB
#include <math.h>
#include <stdint.h>
#include <string.h>
extern int oracle_pert_mode;        /* which result bit decides the direction (1, 2, ...): independent perturbation patterns */
static inline double oracle_pert_(double x) {
  if (x == 0.0 || !isfinite(x)) return x;
  uint64_t u;
  memcpy(&u, &x, sizeof u);
  if ((u >> oracle_pert_mode) & 1u) u += 1; else u -= 1;
  memcpy(&x, &u, sizeof u);
  return x;
}
static inline double oracle_pert_pow(double a, double b) { return oracle_pert_(pow(a, b)); }
static inline double oracle_pert_exp(double a) { return oracle_pert_(exp(a)); }
static inline double oracle_pert_log(double a) { return oracle_pert_(log(a)); }
static inline double oracle_pert_log10(double a) { return oracle_pert_(log10(a)); }
static inline double oracle_pert_atan(double a) { return oracle_pert_(atan(a)); }
#define pow(a, b) oracle_pert_pow(a, b)
#define exp(a) oracle_pert_exp(a)
#define log(a) oracle_pert_log(a)
#define log10(a) oracle_pert_log10(a)
#define atan(a) oracle_pert_atan(a)
#endif
#endif
