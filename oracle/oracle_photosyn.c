/* oracle_photosyn.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the leaf photosynthesis / stomatal conductance path WITHOUT plant hydraulic stress
 * (use_hydrstress = .false., clm4_5 physics): SURVEY.md section 8 row a12.
 *
 * Reference: src/biogeophys/PhotosynthesisMod.F90
 *   Photosynthesis :1243-2063 (called twice per ITERATION pass, phase = 'sun' then 'sha', CanopyFluxesMod.F90:1143-1166)
 *   hybrid :2251-2400   brent :2403-2514   ci_func :2551-2701   ft / fth / fth25 :2517-2548
 * Configuration as on the rest of the hot path: use_cn = .false., lnc_opt = .false., vcmax_opt = 0, nlevcan = 1, no C13.
 * Loops and statement order follow the Fortran; every function cites the lines it restates.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_canopy.h"
#include "oracle_pert.h"

static const double bbbopt_c3 = 10000.0, bbbopt_c4 = 40000.0;       /* :85-86 */
static const double medlyn_rh_can_max = 50.0, medlyn_rh_can_fact = 0.001;   /* :87-88 */
static const double max_cs = 1.e-06;                                 /* :89 */
static const double spval = 1.e36;

static double ft(double tl, double ha) { return exp(ha / (rgas * 1.e-3 * (tfrz + 25.0)) * (1.0 - (tfrz + 25.0) / tl)); }
static double fth(double tl, double hd, double se, double scaleFactor) {
  return scaleFactor / (1.0 + exp((-hd + se * tl) / (rgas * 1.e-3 * tl)));
}
static double fth25(double hd, double se) { return 1.0 + exp((-hd + se * (tfrz + 25.0)) / (rgas * 1.e-3 * (tfrz + 25.0))); }

static void fail(cf_ctx* x, int code, int p) {
  if (!x->err_code) { x->err_code = code; x->err_index = p; }
}

typedef struct leafarg {       /* the scalar arguments hybrid / brent / ci_func pass along */
  int p, c;
  double gb_mol, je, cair, oair, lmr_z, par_z, rh_can;
} leafarg;

/* ci_func :2551-2701 */
static void ci_func(cf_ctx* x, double ci, double* fval, const leafarg* a, double* gs_mol) {
  const ctsm_params_t* pr = x->prm;
  const int p = a->p, c = a->c, ivt = P1(itype, p);
  const double forc_pbot = C1(forc_pbot, c);
  double r1, r2;
  if (P1(c3flag, p)) {
    P2(ac, p, 1, 1) = P2(vcmax_z, p, 1, 1) * fmax(ci - P1(cp, p), 0.0) / (ci + P1(kc, p) * (1.0 + a->oair / P1(ko, p)));
    P2(aj, p, 1, 1) = a->je * fmax(ci - P1(cp, p), 0.0) / (4.0 * ci + 8.0 * P1(cp, p));
    P2(ap, p, 1, 1) = 3.0 * P2(tpu_z, p, 1, 1);
  } else {
    P2(ac, p, 1, 1) = P2(vcmax_z, p, 1, 1);
    P2(aj, p, 1, 1) = P1(qe, p) * a->par_z * 4.6;
    P2(ap, p, 1, 1) = P2(kp_z, p, 1, 1) * fmax(ci, 0.0) / forc_pbot;
  }
  double aquad = PFT(pft_theta_cj, ivt);
  double bquad = -(P2(ac, p, 1, 1) + P2(aj, p, 1, 1));
  double cquad = P2(ac, p, 1, 1) * P2(aj, p, 1, 1);
  if (oracle_quadratic(aquad, bquad, cquad, &r1, &r2)) fail(x, CTSM_ERR_QUADRATIC, p);
  const double ai = fmin(r1, r2);
  aquad = pr->theta_ip;
  bquad = -(ai + P2(ap, p, 1, 1));
  cquad = ai * P2(ap, p, 1, 1);
  if (oracle_quadratic(aquad, bquad, cquad, &r1, &r2)) fail(x, CTSM_ERR_QUADRATIC, p);
  P2(ag, p, 1, 1) = fmax(0.0, fmin(r1, r2));
  P2(an, p, 1, 1) = P2(ag, p, 1, 1) - a->lmr_z;
  if (P2(an, p, 1, 1) < 0.0) { *fval = 0.0; return; }
  double cs = a->cair - 1.4 / a->gb_mol * P2(an, p, 1, 1) * forc_pbot;
  cs = fmax(cs, max_cs);
  if (pr->stomatalcond_mtd == 2) {
    const double mi = PFT(pft_medlynintercept, ivt), ms = PFT(pft_medlynslope, ivt);
    const double term = 1.6 * P2(an, p, 1, 1) / (cs / forc_pbot * 1.e06);
    aquad = 1.0;
    bquad = -(2.0 * (mi * 1.e-06 + term) + ((ms * term) * (ms * term)) / (a->gb_mol * 1.e-06 * a->rh_can));
    cquad = mi * mi * 1.e-12 + (2.0 * mi * 1.e-06 + term * (1.0 - ms * ms / a->rh_can)) * term;
    if (oracle_quadratic(aquad, bquad, cquad, &r1, &r2)) fail(x, CTSM_ERR_QUADRATIC, p);
    *gs_mol = fmax(r1, r2) * 1.e06;
  } else {
    const double bbb = x->bbb[p - x->begp0], mbb = x->mbb[p - x->begp0];
    aquad = cs;
    bquad = cs * (a->gb_mol - bbb) - mbb * P2(an, p, 1, 1) * forc_pbot;
    cquad = -a->gb_mol * (cs * bbb + mbb * P2(an, p, 1, 1) * forc_pbot * a->rh_can);
    if (oracle_quadratic(aquad, bquad, cquad, &r1, &r2)) fail(x, CTSM_ERR_QUADRATIC, p);
    *gs_mol = fmax(r1, r2);
  }
  *fval = ci - a->cair + P2(an, p, 1, 1) * forc_pbot * (1.4 * *gs_mol + 1.6 * a->gb_mol) / (a->gb_mol * *gs_mol);
}

/* brent :2403-2514 */
static void brent(cf_ctx* x, double* xr, double x1, double x2, double f1, double f2, double tol, const leafarg* ar, double* gs_mol) {
  const int itmax = 20;
  const double eps = 1.e-2;
  double a = x1, b = x2, fa = f1, fb = f2, c, fc, d = 0.0, e = 0.0, p, q, r, s, tol1, xm;
  if ((fa > 0.0 && fb > 0.0) || (fa < 0.0 && fb < 0.0)) fail(x, CTSM_ERR_BRENT, ar->p);
  c = b;
  fc = fb;
  int iter = 0;
  for (;;) {
    if (iter == itmax) break;
    iter = iter + 1;
    if ((fb > 0.0 && fc > 0.0) || (fb < 0.0 && fc < 0.0)) { c = a; fc = fa; d = b - a; e = d; }
    if (fabs(fc) < fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
    tol1 = 2.0 * eps * fabs(b) + 0.5 * tol;
    xm = 0.5 * (c - b);
    if (fabs(xm) <= tol1 || fb == 0.) { *xr = b; return; }
    if (fabs(e) >= tol1 && fabs(fa) > fabs(fb)) {
      s = fb / fa;
      if (a == c) {
        p = 2.0 * xm * s;
        q = 1.0 - s;
      } else {
        q = fa / fc;
        r = fb / fc;
        p = s * (2.0 * xm * q * (q - r) - (b - a) * (r - 1.0));
        q = (q - 1.0) * (r - 1.0) * (s - 1.0);
      }
      if (p > 0.0) q = -q;
      p = fabs(p);
      if (2.0 * p < fmin(3.0 * xm * q - fabs(tol1 * q), fabs(e * q))) { e = d; d = p / q; }
      else { d = xm; e = d; }
    } else {
      d = xm;
      e = d;
    }
    a = b;
    fa = fb;
    if (fabs(d) > tol1) b = b + d;
    else b = b + copysign(tol1, xm);
    ci_func(x, b, &fb, ar, gs_mol);
    if (fb == 0.0) break;
  }
  *xr = b;
}

/* hybrid :2251-2400 */
static void hybrid(cf_ctx* x, double* x0, const leafarg* ar, double* gs_mol, int* iter) {
  const double eps = 1.e-2, eps1 = 1.e-4;
  const int itmax = 40;
  double f0, f1, x1, xx, dx, tol, minx, minf;
  ci_func(x, *x0, &f0, ar, gs_mol);
  if (f0 == 0.0) return;
  minx = *x0;
  minf = f0;
  x1 = *x0 * 0.99;
  ci_func(x, x1, &f1, ar, gs_mol);
  if (f1 == 0.0) { *x0 = x1; return; }
  if (f1 < minf) { minx = x1; minf = f1; }
  *iter = 0;
  for (;;) {
    *iter = *iter + 1;
    dx = -f1 * (x1 - *x0) / (f1 - f0);
    xx = x1 + dx;
    tol = fabs(xx) * eps;
    if (fabs(dx) < tol) { *x0 = xx; break; }
    *x0 = x1;
    f0 = f1;
    x1 = xx;
    ci_func(x, x1, &f1, ar, gs_mol);
    if (f1 < minf) { minx = x1; minf = f1; }
    if (fabs(f1) <= eps1) { *x0 = x1; break; }
    if (f1 * f0 < 0.0) {
      brent(x, &xx, *x0, x1, f0, f1, tol, ar, gs_mol);
      *x0 = xx;
      break;
    }
    if (*iter > itmax) {
      ci_func(x, minx, &f1, ar, gs_mol);
      break;
    }
  }
}

/* Photosynthesis :1243-2063; phase: 0 = 'sun', 1 = 'sha'.  esat_tv .. dayl_factor are (begp0:endp0) work arrays of CanopyFluxes */
void oracle_photosynthesis(cf_ctx* x, int fn, const int32_t* filterp, const double* esat_tv, const double* eair,
                           const double* oair, const double* cair, const double* rb, const double* btran,
                           const double* dayl_factor, int phase) {
  const ctsm_params_t* pr = x->prm;
  const ctsm_canopyfluxes_fields_t* F = x->f;
  const int b0 = x->begp0, np = x->np;
  const int medlyn = (pr->stomatalcond_mtd == 2);
  /* the pointer associations of :1399-1437 */
  double* par_z = phase == 0 ? F->parsun_z : F->parsha_z;
  double* lai_z = phase == 0 ? F->laisun_z : F->laisha_z;
  double* vcmaxcint = phase == 0 ? F->vcmaxcintsun : F->vcmaxcintsha;
  double* o3coefv = phase == 0 ? F->o3coefvsun : F->o3coefvsha;
  double* o3coefg = phase == 0 ? F->o3coefgsun : F->o3coefgsha;
  double* ci_z = phase == 0 ? F->cisun_z : F->cisha_z;
  double* rs = phase == 0 ? F->rssun : F->rssha;
  double* rs_z = phase == 0 ? F->rssun_z : F->rssha_z;
  double* lmr = phase == 0 ? F->lmrsun : F->lmrsha;
  double* lmr_z = phase == 0 ? F->lmrsun_z : F->lmrsha_z;
  double* psn = phase == 0 ? F->psnsun : F->psnsha;
  double* psn_z = phase == 0 ? F->psnsun_z : F->psnsha_z;
  double* psn_wc = phase == 0 ? F->psnsun_wc : F->psnsha_wc;
  double* psn_wj = phase == 0 ? F->psnsun_wj : F->psnsha_wj;
  double* psn_wp = phase == 0 ? F->psnsun_wp : F->psnsha_wp;
#define L(arr, p) arr[(p) - b0]
  double* jmax_z = (double*)calloc((size_t)np, sizeof(double));
  double* psn_wc_z = (double*)calloc((size_t)np, sizeof(double));
  double* psn_wj_z = (double*)calloc((size_t)np, sizeof(double));
  double* psn_wp_z = (double*)calloc((size_t)np, sizeof(double));
  const double lmrc = fth25(pr->lmrhd, pr->lmrse);                       /* :1445 */

  for (int f = 0; f < fn; ++f) {                                         /* :1447-1490 */
    const int p = filterp[f], c = P1(column, p), ivt = P1(itype, p);
    if ((int)nearbyint(PFT(pft_c3psn, ivt)) == 1) P1(c3flag, p) = 1;
    else if ((int)nearbyint(PFT(pft_c3psn, ivt)) == 0) P1(c3flag, p) = 0;
    double bbbopt = 0.0;
    if (P1(c3flag, p)) { P1(qe, p) = 0.0; if (!medlyn) bbbopt = bbbopt_c3; }
    else { P1(qe, p) = 0.05; if (!medlyn) bbbopt = bbbopt_c4; }
    if (!medlyn) {
      L(x->bbb, p) = fmax(bbbopt * L(btran, p), 1.0);
      L(x->mbb, p) = PFT(pft_mbbopt, ivt);
    }
    const double kc25 = pr->kc25_coef * C1(forc_pbot, c);
    const double ko25 = pr->ko25_coef * C1(forc_pbot, c);
    const double sco = 0.5 * 0.209 / pr->cp25_yr2000;
    const double cp25 = 0.5 * L(oair, p) / sco;
    P1(kc, p) = kc25 * ft(P1(t_veg, p), pr->kcha);
    P1(ko, p) = ko25 * ft(P1(t_veg, p), pr->koha);
    P1(cp, p) = cp25 * ft(P1(t_veg, p), pr->cpha);
  }

  for (int f = 0; f < fn; ++f) {                                         /* :1500-1755 */
    const int p = filterp[f], ivt = P1(itype, p);
    const double t_veg = P1(t_veg, p), t10 = P1(t_a10, p);
    const double leafcn_local = PFT(pft_leafcn, ivt);
    P1(lnca, p) = 1.0 / (PFT(pft_slatop, ivt) * leafcn_local);           /* :1514 (no upper limit on this path) */
    double vcmax25top = P1(lnca, p) * PFT(pft_flnr, ivt) * pr->fnr * pr->act25 * L(dayl_factor, p);
    vcmax25top = vcmax25top * PFT(pft_fnitr, ivt);
    const double jmax25top = ((2.59 - 0.035 * fmin(fmax((t10 - tfrz), 11.0), 35.0)) * vcmax25top) * pr->jmax25top_sf;
    const double tpu25top = pr->tpu25ratio * vcmax25top;
    const double kp25top = pr->kp25ratio * vcmax25top;
    double lmr25top;
    if (P1(c3flag, p)) lmr25top = vcmax25top * pr->leaf_mr_vcm;
    else lmr25top = vcmax25top * 0.025;
    for (int iv = 1; iv <= P1(nrad, p); ++iv) {
      const double nscaler = L(vcmaxcint, p);                            /* nlevcan == 1 */
      double lmr25 = lmr25top * nscaler;
      const int luna_patch = pr->use_luna && P1(c3flag, p) && PFT(pft_crop, ivt) == 0.0;
      if (luna_patch) lmr25 = pr->leaf_mr_vcm * P2(vcmx25_z, p, iv, 1);
      if (P1(c3flag, p)) {
        L(lmr_z, p) = lmr25 * ft(t_veg, pr->lmrha) * fth(t_veg, pr->lmrhd, pr->lmrse, lmrc);
      } else {
        L(lmr_z, p) = lmr25 * pow(2.0, (t_veg - (tfrz + 25.0)) / 10.0);
        L(lmr_z, p) = L(lmr_z, p) / (1.0 + exp(1.3 * (t_veg - (tfrz + 55.0))));
      }
      if (L(par_z, p) <= 0.0) {
        P2(vcmax_z, p, iv, 1) = 0.0;
        L(jmax_z, p) = 0.0;
        P2(tpu_z, p, iv, 1) = 0.0;
        P2(kp_z, p, iv, 1) = 0.0;
      } else {
        double vcmax25, jmax25, tpu25;
        if (luna_patch) {
          vcmax25 = P2(vcmx25_z, p, iv, 1);
          jmax25 = P2(jmx25_z, p, iv, 1);
          tpu25 = pr->tpu25ratio * vcmax25;
          if (phase == 1 && P1(vcmaxcintsun, p) > 0.0) {
            vcmax25 = vcmax25 * P1(vcmaxcintsha, p) / P1(vcmaxcintsun, p);
            jmax25 = jmax25 * P1(vcmaxcintsha, p) / P1(vcmaxcintsun, p);
            tpu25 = tpu25 * P1(vcmaxcintsha, p) / P1(vcmaxcintsun, p);
          }
        } else {
          vcmax25 = vcmax25top * nscaler;
          jmax25 = jmax25top * nscaler;
          tpu25 = tpu25top * nscaler;
        }
        const double kp25 = kp25top * nscaler;
        const double vcmaxse = (668.39 - 1.07 * fmin(fmax((t10 - tfrz), 11.0), 35.0)) * pr->vcmaxse_sf;
        const double jmaxse = (659.70 - 0.75 * fmin(fmax((t10 - tfrz), 11.0), 35.0)) * pr->jmaxse_sf;
        const double tpuse = (668.39 - 1.07 * fmin(fmax((t10 - tfrz), 11.0), 35.0)) * pr->tpuse_sf;
        const double vcmaxc = fth25(pr->vcmaxhd, vcmaxse);
        const double jmaxc = fth25(pr->jmaxhd, jmaxse);
        const double tpuc = fth25(pr->tpuhd, tpuse);
        P2(vcmax_z, p, iv, 1) = vcmax25 * ft(t_veg, pr->vcmaxha) * fth(t_veg, pr->vcmaxhd, vcmaxse, vcmaxc);
        L(jmax_z, p) = jmax25 * ft(t_veg, pr->jmaxha) * fth(t_veg, pr->jmaxhd, jmaxse, jmaxc);
        P2(tpu_z, p, iv, 1) = tpu25 * ft(t_veg, pr->tpuha) * fth(t_veg, pr->tpuhd, tpuse, tpuc);
        if (!P1(c3flag, p)) {
          P2(vcmax_z, p, iv, 1) = vcmax25 * pow(2.0, (t_veg - (tfrz + 25.0)) / 10.0);
          P2(vcmax_z, p, iv, 1) = P2(vcmax_z, p, iv, 1) / (1.0 + exp(0.2 * ((tfrz + 15.0) - t_veg)));
          P2(vcmax_z, p, iv, 1) = P2(vcmax_z, p, iv, 1) / (1.0 + exp(0.3 * (t_veg - (tfrz + 40.0))));
        }
        P2(kp_z, p, iv, 1) = kp25 * pow(2.0, (t_veg - (tfrz + 25.0)) / 10.0);
      }
      P2(vcmax_z, p, iv, 1) = P2(vcmax_z, p, iv, 1) * L(btran, p);       /* :1745-1746 soil water stress */
      L(lmr_z, p) = L(lmr_z, p) * L(btran, p);
      if (pr->light_inhibit && L(par_z, p) > 0.0) L(lmr_z, p) = L(lmr_z, p) * 0.67;   /* :1749-1751 */
    }
  }

  const double rsmax0 = 2.e4;
  for (int f = 0; f < fn; ++f) {                                         /* :1762-2009 */
    const int p = filterp[f], c = P1(column, p), g = P1(gridcell, p), ivt = P1(itype, p);
    const double forc_pbot = C1(forc_pbot, c);
    const double cf = forc_pbot / (rgas * 1.e-3 * P1(thm, p)) * 1.e06;
    const double gb = 1.0 / L(rb, p);
    P1(gb_mol, p) = gb * cf;
    for (int iv = 1; iv <= P1(nrad, p); ++iv) {
      if (L(par_z, p) <= 0.0) {                                          /* night :1781-1815 */
        P2(ac, p, iv, 1) = 0.0;
        P2(aj, p, iv, 1) = 0.0;
        P2(ap, p, iv, 1) = 0.0;
        P2(ag, p, iv, 1) = 0.0;
        P2(an, p, iv, 1) = P2(ag, p, iv, 1) - L(lmr_z, p);
        L(psn_z, p) = 0.0;
        L(psn_wc_z, p) = 0.0;
        L(psn_wj_z, p) = 0.0;
        L(psn_wp_z, p) = 0.0;
        if (!medlyn) L(rs_z, p) = fmin(rsmax0, 1.0 / L(x->bbb, p) * cf);
        else L(rs_z, p) = fmin(rsmax0, 1.0 / PFT(pft_medlynintercept, ivt) * cf);
        L(ci_z, p) = 0.0;
        P1(rh_leaf, p) = 0.0;
        if (phase == 0) P2(gs_mol_sun, p, iv, 1) = cf / L(rs_z, p);
        else P2(gs_mol_sha, p, iv, 1) = cf / L(rs_z, p);
      } else {                                                           /* day :1817-2006 */
        const double ceair = fmin(L(eair, p), L(esat_tv, p));
        double rh_can;
        if (!medlyn) {
          rh_can = ceair / L(esat_tv, p);
        } else {
          rh_can = fmax((L(esat_tv, p) - ceair), medlyn_rh_can_max) * medlyn_rh_can_fact;
          P1(vpd_can, p) = rh_can;
        }
        const double qabs = 0.5 * (1.0 - pr->fnps) * L(par_z, p) * 4.6;
        double r1, r2;
        if (oracle_quadratic(pr->theta_psii, -(qabs + L(jmax_z, p)), qabs * L(jmax_z, p), &r1, &r2)) fail(x, CTSM_ERR_QUADRATIC, p);
        const double je = fmin(r1, r2);
        if (P1(c3flag, p)) L(ci_z, p) = 0.7 * L(cair, p);
        else L(ci_z, p) = 0.4 * L(cair, p);
        int niter = 0;
        niter = niter + 1;
        double ciold = L(ci_z, p);
        leafarg a;
        a.p = p; a.c = c; a.gb_mol = P1(gb_mol, p); a.je = je; a.cair = L(cair, p); a.oair = L(oair, p); a.lmr_z = L(lmr_z, p);
        a.par_z = L(par_z, p); a.rh_can = rh_can;
        double gs_mol = P2(gs_mol, p, iv, 1);
        hybrid(x, &ciold, &a, &gs_mol, &niter);
        P2(gs_mol, p, iv, 1) = gs_mol;
        if (P2(an, p, iv, 1) < 0.0) {                                     /* :1880-1886 */
          if (!medlyn) P2(gs_mol, p, iv, 1) = L(x->bbb, p);
          else P2(gs_mol, p, iv, 1) = PFT(pft_medlynintercept, ivt);
        }
        if (phase == 0) P2(gs_mol_sun, p, iv, 1) = P2(gs_mol, p, iv, 1);
        else P2(gs_mol_sha, p, iv, 1) = P2(gs_mol, p, iv, 1);
        if (G1(near_local_noon, g)) {                                    /* :1895-1908 */
          if (phase == 0) P2(gs_mol_sun_ln, p, iv, 1) = P2(gs_mol, p, iv, 1);
          else P2(gs_mol_sha_ln, p, iv, 1) = P2(gs_mol, p, iv, 1);
        } else {
          if (phase == 0) P2(gs_mol_sun_ln, p, iv, 1) = spval;
          else P2(gs_mol_sha_ln, p, iv, 1) = spval;
        }
        double cs = L(cair, p) - 1.4 / P1(gb_mol, p) * P2(an, p, iv, 1) * forc_pbot;
        cs = fmax(cs, max_cs);
        L(ci_z, p) = L(cair, p) - P2(an, p, iv, 1) * forc_pbot * (1.4 * P2(gs_mol, p, iv, 1) + 1.6 * P1(gb_mol, p))
                                      / (P1(gb_mol, p) * P2(gs_mol, p, iv, 1));
        L(ci_z, p) = fmax(L(ci_z, p), 1.e-06);
        const double gs = P2(gs_mol, p, iv, 1) / cf;
        L(rs_z, p) = fmin(1.0 / gs, rsmax0);
        L(rs_z, p) = L(rs_z, p) / L(o3coefg, p);
        L(psn_z, p) = P2(ag, p, iv, 1);
        L(psn_z, p) = L(psn_z, p) * L(o3coefv, p);
        L(psn_wc_z, p) = 0.0;
        L(psn_wj_z, p) = 0.0;
        L(psn_wp_z, p) = 0.0;
        if (P2(ac, p, iv, 1) <= P2(aj, p, iv, 1) && P2(ac, p, iv, 1) <= P2(ap, p, iv, 1)) L(psn_wc_z, p) = L(psn_z, p);
        else if (P2(aj, p, iv, 1) < P2(ac, p, iv, 1) && P2(aj, p, iv, 1) <= P2(ap, p, iv, 1)) L(psn_wj_z, p) = L(psn_z, p);
        else if (P2(ap, p, iv, 1) < P2(ac, p, iv, 1) && P2(ap, p, iv, 1) < P2(aj, p, iv, 1)) L(psn_wp_z, p) = L(psn_z, p);
        if (P2(gs_mol, p, iv, 1) < 0.0) fail(x, CTSM_ERR_GS_NEG, p);      /* :1984-1988 */
        if (!medlyn) {                                                   /* :1991-2004 (the error check only writes to the log) */
          const double hs = (P1(gb_mol, p) * ceair + P2(gs_mol, p, iv, 1) * L(esat_tv, p))
                            / ((P1(gb_mol, p) + P2(gs_mol, p, iv, 1)) * L(esat_tv, p));
          P1(rh_leaf, p) = hs;
          const double gs_mol_err = L(x->mbb, p) * fmax(P2(an, p, iv, 1), 0.0) * hs / cs * forc_pbot + L(x->bbb, p);
          if (fabs(P2(gs_mol, p, iv, 1) - gs_mol_err) > 1.e-01) x->n_warnings++;
        }
      }
    }
  }

  for (int f = 0; f < fn; ++f) {                                         /* canopy sums :2015-2058 (nlevcan = 1) */
    const int p = filterp[f];
    double psncan = 0.0, psncan_wc = 0.0, psncan_wj = 0.0, psncan_wp = 0.0, lmrcan = 0.0, gscan = 0.0, laican = 0.0;
    for (int iv = 1; iv <= P1(nrad, p); ++iv) {
      psncan = psncan + L(psn_z, p) * L(lai_z, p);
      psncan_wc = psncan_wc + L(psn_wc_z, p) * L(lai_z, p);
      psncan_wj = psncan_wj + L(psn_wj_z, p) * L(lai_z, p);
      psncan_wp = psncan_wp + L(psn_wp_z, p) * L(lai_z, p);
      lmrcan = lmrcan + L(lmr_z, p) * L(lai_z, p);
      gscan = gscan + L(lai_z, p) / (L(rb, p) + L(rs_z, p));
      laican = laican + L(lai_z, p);
    }
    if (laican > 0.0) {
      L(psn, p) = psncan / laican;
      L(psn_wc, p) = psncan_wc / laican;
      L(psn_wj, p) = psncan_wj / laican;
      L(psn_wp, p) = psncan_wp / laican;
      L(lmr, p) = lmrcan / laican;
      L(rs, p) = laican / gscan - L(rb, p);
    } else {
      L(psn, p) = 0.0;
      L(psn_wc, p) = 0.0;
      L(psn_wj, p) = 0.0;
      L(psn_wp, p) = 0.0;
      L(lmr, p) = 0.0;
      L(rs, p) = 0.0;
    }
  }
  free(jmax_z); free(psn_wc_z); free(psn_wj_z); free(psn_wp_z);
#undef L
}
