// phs_tasks_check.cu — TEST INFRASTRUCTURE (host build, no GPU needed).
//
// Runs the resumable task formulation of the PHS solve (ctsm_b200/csrc/phs.cuh: newton_begin/step/finish, ci_task_begin /
// ci_step / ci_task_end — what the lane scheduler of canopy.cu executes) on the CPU against the direct nested-loop
// formulation of hybrid_PHS / brent_PHS (tests/host/phs_direct.cuh) over randomised patches.  Both are the same
// arithmetic in the same order, so every output must agree BIT FOR BIT; the program prints branch-coverage counters
// and exits non-zero on the first difference.
//
//   nvcc -O2 -std=c++17 -Xcompiler -ffp-contract=off -o phs_tasks_check tests/host/phs_tasks_check.cu && ./phs_tasks_check N SEED
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "phs_direct.cuh"

using namespace phs;

struct Case {
  PhsPatch P;
  Leaf L;
  double K[NLEVSOI], G[NLEVSOI], S[NLEVSOI];
  double vegwp[4];
  double ci0;
};

static double U(std::mt19937_64& r, double a, double b) { return a + (b - a) * std::uniform_real_distribution<double>(0.0, 1.0)(r); }

static void make_case(std::mt19937_64& r, Case& c) {
  PhsPatch& P = c.P;
  Leaf& L = c.L;
  const double psi = -U(r, 1.5e5, 5.5e5);
  for (int s = 0; s < 4; ++s) { P.psi50[s] = psi * U(r, 0.8, 1.2); P.ck[s] = U(r, 2.5, 5.0); P.kmax[s] = 2.0e-8 * U(r, 0.3, 3.0); }
  P.elai = U(r, 0.1, 6.0);
  P.esai = U(r, 0.1, 1.0);
  const double fsun = U(r, 0.1, 0.9);
  const double pick = U(r, 0.0, 1.0);
  P.laisun = pick < 0.04 ? 5.0e-4 : fsun * P.elai;
  P.laisha = P.elai - P.laisun;
  if (pick > 0.97) { P.laisha = 5.0e-4; P.laisun = P.elai - P.laisha; }
  P.tsai = U(r, 0.0, 1.0) < 0.05 ? 0.0 : U(r, 0.1, 1.0);
  P.htop = U(r, 0.2, 35.0);
  P.fdry = U(r, 0.3, 1.0) * P.elai / (P.elai + P.esai);
  P.forc_rho = U(r, 0.9, 1.3);
  P.forc_pbot = U(r, 7.0e4, 1.02e5);
  const double thm = U(r, 255.0, 310.0);
  P.cf = P.forc_pbot / (rgas * 1.e-3 * thm) * 1.e06;
  P.qsatl = U(r, 0.002, 0.03);
  P.qaf = P.qsatl * (U(r, 0.0, 1.0) < 0.05 ? U(r, 1.0, 1.1) : U(r, 0.2, 0.99));
  P.gb_mol = (1.0 / U(r, 5.0, 80.0)) * P.cf;
  const bool dry = U(r, 0.0, 1.0) < 0.25;
  double ksum = 0.0, ksmp = 0.0, ksmpg = 0.0, smpg = 0.0, z = 0.01;
  const bool nok = U(r, 0.0, 1.0) < 0.02;
  for (int j = 0; j < NLEVSOI; ++j) {
    z += U(r, 0.02, 0.6);
    c.G[j] = 1000.0 * z;
    c.S[j] = dry ? -pow(10.0, U(r, 4.5, 8.0)) : -pow(10.0, U(r, 2.0, 5.5));
    c.K[j] = (j == 0 || nok) ? 0.0 : pow(10.0, dry ? U(r, -14.0, -10.0) : U(r, -11.0, -7.5));
    ksum += c.K[j]; ksmp += c.K[j] * c.S[j]; ksmpg += c.K[j] * (c.S[j] - c.G[j]); smpg += c.S[j] - c.G[j];
  }
  P.ksum = ksum; P.ksmp = ksmp; P.ksmpg = ksmpg; P.smpg_mean = smpg / NLEVSOI;
  P.sk = c.K; P.sg = c.G; P.ss = c.S; P.stride = 1;
  const double wroot = -pow(10.0, U(r, 3.5, 6.0));
  const double wxyl = wroot - U(r, 0.0, 4.0e4);
  c.vegwp[0] = wxyl - U(r, 0.0, 2.0e4); c.vegwp[1] = wxyl - U(r, 0.0, 2.0e4); c.vegwp[2] = wxyl; c.vegwp[3] = wroot;

  L.c3 = U(r, 0.0, 1.0) < 0.8;
  L.medlyn = U(r, 0.0, 1.0) < 0.8;
  L.qe = L.c3 ? 0.0 : 0.05;
  L.bbb = L.c3 ? 10000.0 : 40000.0;
  L.mbb = L.c3 ? 9.0 : 4.0;
  L.medint = 100.0; L.medslope = U(r, 1.5, 6.0);
  L.theta_cj = L.c3 ? 0.9393 : 0.80; L.theta_ip = 0.95;
  L.cair = 40.0 * U(r, 0.9, 1.1); L.oair = 0.209e5;
  L.kc = U(r, 10.0, 120.0); L.ko = U(r, 1.5e4, 5.0e4); L.cp = U(r, 1.5, 8.0);
  const double vbase = pow(10.0, U(r, 0.3, 2.2));
  for (int s = 0; s < 2; ++s) {
    L.vcmax[s] = vbase * U(r, 0.3, 2.0);
    L.tpu[s] = 0.167 * L.vcmax[s] * U(r, 0.5, 1.5);
    L.kp[s] = 2.0e4 * L.vcmax[s];
    L.lmr[s] = 0.015 * L.vcmax[s] * U(r, 0.2, 3.0);
    L.par[s] = U(r, 5.0, 400.0) * (s ? 0.3 : 1.0);
    L.je[s] = fmin(0.5 * 0.85 * L.par[s] * 4.6, 1.9 * L.vcmax[s]) * U(r, 0.6, 1.0);
  }
  if (L.medlyn) L.rh_can = fmax(U(r, -200.0, 3000.0), 50.0) * 0.001;
  else L.rh_can = U(r, 0.2, 1.0);
  c.ci0 = (L.c3 ? 0.7 : 0.4) * L.cair;
}

struct TaskOut { HybridOut h; CiOut o; bool bad, nb; };
static long long n_brent = 0, n_final = 0, n_newton_tasks = 0, n_newton_steps = 0, n_ci_steps = 0, n_flag = 0, n_itmax = 0, n_outer = 0;

// the flow the GPU kernels implement: ci task -> (newton task -> ci task)* -> getvegwp
static TaskOut run_tasks(const Case& c) {
  TaskOut t;
  t.bad = false; t.nb = false;
  memset(&t.o, 0, sizeof(t.o));
  HybridCarry H;
  hybrid_carry_init(H, c.ci0);
  CiLane C;
  Brent B;
  memset(&C, 0, sizeof(C));
  memset(&B, 0, sizeof(B));
  for (;;) {
    ++n_outer;
    if (H.iter1 > 1) {                       // newton task
      Newton N;
      ++n_newton_tasks;
      if (newton_begin(N, c.P, c.vegwp, H.gs0sun, H.gs0sha)) {
        while (newton_step(N, c.P)) ++n_newton_steps;
        if (N.iter > 50) ++n_itmax;
      }
      if (N.flag) ++n_flag;
      double unused = 0.0;
      const Stress s = newton_finish(N, c.P, H.gs0sun, H.gs0sha, &unused);
      H.bsun = s.bsun; H.bsha = s.bsha;
    }
    ci_task_begin(C, H);
    for (;;) {
      ++n_ci_steps;
      const bool more = ci_step(C, B, c.P, c.L, &t.bad, &t.nb);
      if (C.st == CI_BRENT && more && C.biter == 1) ++n_brent;
      if (C.st == CI_FINAL && more) ++n_final;
      if (!more) break;
    }
    if (ci_task_end(C, H)) break;
  }
  double x[4];
  double sf = getvegwp(c.P, x, C.gs_sun, C.gs_sha);
  if (sf < 0.0) sf = 0.0;
  t.h.bsun = H.bsun; t.h.bsha = H.bsha; t.h.gs_sun = C.gs_sun; t.h.gs_sha = C.gs_sha; t.h.tran = sf;
  for (int i = 0; i < 4; ++i) t.h.x[i] = x[i];
  t.o = C.o;
  return t;
}

static bool same(double a, double b) { return memcmp(&a, &b, sizeof(double)) == 0; }

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 200000;
  const unsigned long long seed = argc > 2 ? strtoull(argv[2], nullptr, 10) : 12345ull;
  std::mt19937_64 r(seed);
  long nbad = 0, nnb = 0;
  for (long i = 0; i < n; ++i) {
    Case c;
    make_case(r, c);
    CiOut o;
    memset(&o, 0, sizeof(o));
    bool bad = false, nb = false;
    const HybridOut h = hybrid(c.P, c.L, c.vegwp, c.ci0, o, &bad, &nb);
    const TaskOut t = run_tasks(c);
    bool ok = same(h.bsun, t.h.bsun) && same(h.bsha, t.h.bsha) && same(h.gs_sun, t.h.gs_sun) && same(h.gs_sha, t.h.gs_sha) &&
              same(h.tran, t.h.tran) && bad == t.bad && nb == t.nb;
    for (int k = 0; k < 4; ++k) ok = ok && same(h.x[k], t.h.x[k]);
    for (int s = 0; s < 2; ++s)
      ok = ok && same(o.ac[s], t.o.ac[s]) && same(o.aj[s], t.o.aj[s]) && same(o.ap[s], t.o.ap[s]) && same(o.ag[s], t.o.ag[s]) &&
           same(o.an[s], t.o.an[s]);
    if (!ok) {
      printf("MISMATCH at case %ld: direct bsun=%.17g bsha=%.17g gs=(%.17g,%.17g) tran=%.17g | tasks bsun=%.17g bsha=%.17g gs=(%.17g,%.17g) tran=%.17g\n",
             i, h.bsun, h.bsha, h.gs_sun, h.gs_sha, h.tran, t.h.bsun, t.h.bsha, t.h.gs_sun, t.h.gs_sha, t.h.tran);
      return 1;
    }
    nbad += bad; nnb += nb;
  }
  printf("phs_tasks_check: %ld cases identical; outer passes %lld, ci evaluations %lld, brent entries %lld, minx re-evaluations %lld, "
         "newton tasks %lld (steps %lld, itmax hits %lld, flag exits %lld), bad quadratics %ld, not bracketed %ld\n",
         n, n_outer, n_ci_steps, n_brent, n_final, n_newton_tasks, n_newton_steps, n_itmax, n_flag, nbad, nnb);
  return 0;
}
