/* oracle_canopy.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of CanopyFluxes (src/biogeophys/CanopyFluxesMod.F90:191-1765)
 * and the routines it calls outside PhotosynthesisMod:
 *   QSat                       src/biogeophys/QSatMod.F90:61-127
 *   FrictionVelocity           src/biogeophys/FrictionVelocityMod.F90:754-1117
 *   StabilityFunc1/2           :1120-1155       MoninObukIni :1158-1209
 *   calc_effective_soilporosity, calc_volumetric_h2oliq, calc_root_moist_stress
 *                              src/biogeophys/SoilMoistStressMod.F90:70-114,171-217,312-514
 *   soil_suction (CH78)        SoilWaterRetentionCurveClappHornberg1978Mod.F90:115-120
 *   truncate_small_values      src/utils/NumericsMod.F90:50-99
 *   setExposedvegpFilter       src/main/filterMod.F90:595-648
 * Configuration: use_hydrstress=.true., use_fates=use_cn=use_lch4=.false.,
 * perchroot=perchroot_alt=.false., human stress indices / ozone uptake / LUNA
 * accumulators not computed.
 * Loop structure follows the Fortran: filter loops over clump-sized scratch
 * arrays, the ITERATION loop with order-preserving filter compaction.
 * PARITY UNPINNED by the reference's own tests; invariants in
 * tests/test_oracle_canopy.py (canopy energy closure |err| <= 0.1 W/m2, :1750).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_canopy.h"
#include "oracle_pert.h"

static const double sb = 5.67e-8, cpair = 1.00464e3, hvap = 2.501e6, vkc = 0.4, grav = 9.80616;
static const double denice = 0.917e3, denh2o = 1.000e3, c_to_b = 2.0, tlsai_crit = 2.0, alpha_aero = 1.0;
static const double c_water = 4.188e3, c_dry_biomass = 1400.0, nu_param = 1.5e-5, cd1_param = 7.5;
static const double spval = 1.e36;

/* QSatMod.F90:18-56 */
static const double a0 = 6.11213476, a1 = 0.444007856, a2 = 0.143064234e-01, a3 = 0.264461437e-03, a4 = 0.305903558e-05,
                    a5 = 0.196237241e-07, a6 = 0.892344772e-10, a7 = -0.373208410e-12, a8 = 0.209339997e-15;
static const double b0 = 0.444017302, b1 = 0.286064092e-01, b2 = 0.794683137e-03, b3 = 0.121211669e-04, b4 = 0.103354611e-06,
                    b5 = 0.404125005e-09, b6 = -0.788037859e-12, b7 = -0.114596802e-13, b8 = 0.381294516e-16;
static const double c0 = 6.11123516, c1 = 0.503109514, c2 = 0.188369801e-01, c3 = 0.420547422e-03, c4 = 0.614396778e-05,
                    c5 = 0.602780717e-07, c6 = 0.387940929e-09, c7 = 0.149436277e-11, c8 = 0.262655803e-14;
static const double d0 = 0.503277922, d1 = 0.377289173e-01, d2 = 0.126801703e-02, d3 = 0.249468427e-04, d4 = 0.313703411e-06,
                    d5 = 0.257180651e-08, d6 = 0.133268878e-10, d7 = 0.394116744e-13, d8 = 0.498070196e-16;

/* QSat :61-127; es/qsdT optional (NULL = absent) */
void oracle_qsat(double T, double p, double* qs, double* es, double* qsdT) {
  double es_local, esdT_local, td, vp, vp1, vp2;
  td = fmin(100.0, fmax(-75.0, T - tfrz));
  if (td >= 0.0)
    es_local = a0 + td * (a1 + td * (a2 + td * (a3 + td * (a4 + td * (a5 + td * (a6 + td * (a7 + td * a8)))))));
  else
    es_local = c0 + td * (c1 + td * (c2 + td * (c3 + td * (c4 + td * (c5 + td * (c6 + td * (c7 + td * c8)))))));
  es_local = es_local * 100.0;
  vp = 1.0 / (p - 0.378 * es_local);
  vp1 = 0.622 * vp;
  *qs = es_local * vp1;
  if (es) *es = es_local;
  if (qsdT) {
    if (td >= 0.0)
      esdT_local = b0 + td * (b1 + td * (b2 + td * (b3 + td * (b4 + td * (b5 + td * (b6 + td * (b7 + td * b8)))))));
    else
      esdT_local = d0 + td * (d1 + td * (d2 + td * (d3 + td * (d4 + td * (d5 + td * (d6 + td * (d7 + td * d8)))))));
    esdT_local = esdT_local * 100.0;
    vp2 = vp1 * vp;
    *qsdT = esdT_local * vp2 * p;
  }
}

/* FrictionVelocityMod.F90:1120-1155 */
static double StabilityFunc1(double zeta) {
  const double chik2 = sqrt(1.0 - 16.0 * zeta);
  const double chik = sqrt(chik2);
  return 2.0 * log((1.0 + chik) * 0.5) + log((1.0 + chik2) * 0.5) - 2.0 * atan(chik) + rpi * 0.5;
}
static double StabilityFunc2(double zeta) {
  const double chik2 = sqrt(1.0 - 16.0 * zeta);
  return 2.0 * log((1.0 + chik2) * 0.5);
}

/* MoninObukIni :1158-1209 */
void oracle_moninobukini(double zetamaxstable, double ur, double thv, double dthv, double zldis, double z0m, double* um,
                         double* obu) {
  double wc, rib, zeta;
  wc = 0.5;
  if (dthv >= 0.0) *um = fmax(ur, 0.1);
  else *um = sqrt(ur * ur + wc * wc);
  rib = grav * zldis * dthv / (thv * *um * *um);
  if (rib >= 0.0) {
    zeta = rib * log(zldis / z0m) / (1.0 - 5.0 * fmin(rib, 0.19));
    zeta = fmin(zetamaxstable, fmax(zeta, 0.01));
  } else {
    zeta = rib * log(zldis / z0m);
    zeta = fmax(-100.0, fmin(zeta, -0.01));
  }
  *obu = zldis / zeta;
}

/* FrictionVelocity :754-1117 for one point (patch-level form, landunit_index absent); o->fm carries fm(n) in and out */
void oracle_friction_velocity_point(double hgt_u, double hgt_t, double hgt_q, double displa, double z0m, double z0h, double z0q,
                                    double obu, int iter, double ur, double um, oracle_fricvel_t* o) {
  const double zetam = 1.574, zetat = 0.465;
  double zldis, zeta, tmp1, tmp2, tmp3, tmp4, fmnew, fm10, zeta10, vds_tmp;
  /* wind profile */
  zldis = hgt_u - displa;
  zeta = zldis / obu;
  if (zeta < -zetam) {
    o->ustar = vkc * um / (log(-zetam * obu / z0m) - StabilityFunc1(-zetam) + StabilityFunc1(z0m / obu)
                              + 1.14 * (pow(-zeta, 0.333) - pow(zetam, 0.333)));
  } else if (zeta < 0.0) {
    o->ustar = vkc * um / (log(zldis / z0m) - StabilityFunc1(zeta) + StabilityFunc1(z0m / obu));
  } else if (zeta <= 1.0) {
    o->ustar = vkc * um / (log(zldis / z0m) + 5.0 * zeta - 5.0 * z0m / obu);
  } else {
    o->ustar = vkc * um / (log(obu / z0m) + 5.0 - 5.0 * z0m / obu + (5.0 * log(zeta) + zeta - 1.0));
  }
  if (zeta < 0.0) vds_tmp = 2.e-3 * o->ustar * (1.0 + pow(300.0 / (-obu), 0.666));
  else vds_tmp = 2.e-3 * o->ustar;
  o->vds = vds_tmp;
  /* 10-m wind */
  if (zldis - z0m <= 10.0) {
    o->u10_clm = um;
  } else {
    if (zeta < -zetam) {
      o->u10_clm = um - (o->ustar / vkc * (log(-zetam * obu / (10.0 + z0m)) - StabilityFunc1(-zetam)
                                                    + StabilityFunc1((10.0 + z0m) / obu)
                                                    + 1.14 * (pow(-zeta, 0.333) - pow(zetam, 0.333))));
    } else if (zeta < 0.0) {
      o->u10_clm = um - (o->ustar / vkc * (log(zldis / (10.0 + z0m)) - StabilityFunc1(zeta)
                                                    + StabilityFunc1((10.0 + z0m) / obu)));
    } else if (zeta <= 1.0) {
      o->u10_clm = um - (o->ustar / vkc * (log(zldis / (10.0 + z0m)) + 5.0 * zeta - 5.0 * (10.0 + z0m) / obu));
    } else {
      o->u10_clm = um - (o->ustar / vkc * (log(obu / (10.0 + z0m)) + 5.0 - 5.0 * (10.0 + z0m) / obu
                                                    + (5.0 * log(zeta) + zeta - 1.0)));
    }
  }
  o->va = um;
  /* temperature profile */
  zldis = hgt_t - displa;
  zeta = zldis / obu;
  if (zeta < -zetat) {
    o->temp1 = vkc / (log(-zetat * obu / z0h) - StabilityFunc2(-zetat) + StabilityFunc2(z0h / obu)
                      + 0.8 * (pow(zetat, -0.333) - pow(-zeta, -0.333)));
  } else if (zeta < 0.0) {
    o->temp1 = vkc / (log(zldis / z0h) - StabilityFunc2(zeta) + StabilityFunc2(z0h / obu));
  } else if (zeta <= 1.0) {
    o->temp1 = vkc / (log(zldis / z0h) + 5.0 * zeta - 5.0 * z0h / obu);
  } else {
    o->temp1 = vkc / (log(obu / z0h) + 5.0 - 5.0 * z0h / obu + (5.0 * log(zeta) + zeta - 1.0));
  }
  /* humidity profile */
  if (hgt_q == hgt_t && z0q == z0h) {
    o->temp2 = o->temp1;
  } else {
    zldis = hgt_q - displa;
    zeta = zldis / obu;
    if (zeta < -zetat) {
      o->temp2 = vkc / (log(-zetat * obu / z0q) - StabilityFunc2(-zetat) + StabilityFunc2(z0q / obu)
                        + 0.8 * (pow(zetat, -0.333) - pow(-zeta, -0.333)));
    } else if (zeta < 0.0) {
      o->temp2 = vkc / (log(zldis / z0q) - StabilityFunc2(zeta) + StabilityFunc2(z0q / obu));
    } else if (zeta <= 1.0) {
      o->temp2 = vkc / (log(zldis / z0q) + 5.0 * zeta - 5.0 * z0q / obu);
    } else {
      o->temp2 = vkc / (log(obu / z0q) + 5.0 - 5.0 * z0q / obu + (5.0 * log(zeta) + zeta - 1.0));
    }
  }
  /* temperature profile applied at 2-m */
  zldis = 2.0 + z0h;
  zeta = zldis / obu;
  if (zeta < -zetat) {
    o->temp12m = vkc / (log(-zetat * obu / z0h) - StabilityFunc2(-zetat) + StabilityFunc2(z0h / obu)
                        + 0.8 * (pow(zetat, -0.333) - pow(-zeta, -0.333)));
  } else if (zeta < 0.0) {
    o->temp12m = vkc / (log(zldis / z0h) - StabilityFunc2(zeta) + StabilityFunc2(z0h / obu));
  } else if (zeta <= 1.0) {
    o->temp12m = vkc / (log(zldis / z0h) + 5.0 * zeta - 5.0 * z0h / obu);
  } else {
    o->temp12m = vkc / (log(obu / z0h) + 5.0 - 5.0 * z0h / obu + (5.0 * log(zeta) + zeta - 1.0));
  }
  /* humidity profile applied at 2-m */
  if (z0q == z0h) {
    o->temp22m = o->temp12m;
  } else {
    zldis = 2.0 + z0q;
    zeta = zldis / obu;
    if (zeta < -zetat) {
      o->temp22m = vkc / (log(-zetat * obu / z0q) - StabilityFunc2(-zetat) + StabilityFunc2(z0q / obu)
                          + 0.8 * (pow(zetat, -0.333) - pow(-zeta, -0.333)));
    } else if (zeta < 0.0) {
      o->temp22m = vkc / (log(zldis / z0q) - StabilityFunc2(zeta) + StabilityFunc2(z0q / obu));
    } else if (zeta <= 1.0) {
      o->temp22m = vkc / (log(zldis / z0q) + 5.0 * zeta - 5.0 * z0q / obu);
    } else {
      o->temp22m = vkc / (log(obu / z0q) + 5.0 - 5.0 * z0q / obu + (5.0 * log(zeta) + zeta - 1.0));
    }
  }
  /* 10-m wind for the dust model */
  zldis = hgt_u - displa;
  zeta = zldis / obu;
  if (fmin(zeta, 1.0) < 0.0) {
    tmp1 = pow(1.0 - 16.0 * fmin(zeta, 1.0), 0.25);
    tmp2 = log((1.0 + tmp1 * tmp1) / 2.0);
    tmp3 = log((1.0 + tmp1) / 2.0);
    fmnew = 2.0 * tmp3 + tmp2 - 2.0 * atan(tmp1) + 1.5707963;
  } else {
    fmnew = -5.0 * fmin(zeta, 1.0);
  }
  if (iter == 1) o->fm = fmnew;
  else o->fm = 0.5 * (o->fm + fmnew);
  zeta10 = fmin(10.0 / obu, 1.0);
  if (zeta == 0.0) zeta10 = 0.0;
  if (zeta10 < 0.0) {
    tmp1 = pow(1.0 - 16.0 * zeta10, 0.25);
    tmp2 = log((1.0 + tmp1 * tmp1) / 2.0);
    tmp3 = log((1.0 + tmp1) / 2.0);
    fm10 = 2.0 * tmp3 + tmp2 - 2.0 * atan(tmp1) + 1.5707963;
  } else {
    fm10 = -5.0 * zeta10;
  }
  tmp4 = log(fmax(1.0, hgt_u / 10.0));
  o->u10 = ur - o->ustar / vkc * (tmp4 - o->fm + fm10);
  o->fv = o->ustar;
}

/* FrictionVelocity over a patch filter.  Dummy arrays are (begp0:endp0). */
static void FrictionVelocity(cf_ctx* x, int fn, const int32_t* filtern, const double* displa, const double* z0m,
                             const double* z0h, const double* z0q, const double* obu, int iter, const double* ur,
                             const double* um, double* ustar, double* temp1, double* temp2, double* temp12m,
                             double* temp22m, double* fm) {
  const int o = x->begp0;
  for (int f = 0; f < fn; ++f) {
    const int n = filtern[f], i = n - o;
    oracle_fricvel_t r;
    r.fm = fm[i];
    oracle_friction_velocity_point(P1(forc_hgt_u_patch, n), P1(forc_hgt_t_patch, n), P1(forc_hgt_q_patch, n), displa[i], z0m[i],
                                   z0h[i], z0q[i], obu[i], iter, ur[i], um[i], &r);
    ustar[i] = r.ustar; temp1[i] = r.temp1; temp2[i] = r.temp2; temp12m[i] = r.temp12m; temp22m[i] = r.temp22m; fm[i] = r.fm;
    P1(vds, n) = r.vds; P1(u10_clm, n) = r.u10_clm; P1(va, n) = r.va; P1(u10, n) = r.u10; P1(fv, n) = r.fv;
  }
}

/* setExposedvegpFilter, filterMod.F90:595-648 */
void oracle_set_exposedvegp_filter(const ctsm_bounds_t* bounds, int num_nolakeurbanp, const int32_t* nolakeurbanp,
                                   const int32_t* frac_veg_nosno, int32_t* exposedvegp, int32_t* num_exposedvegp,
                                   int32_t* noexposedvegp, int32_t* num_noexposedvegp) {
  int fe = 0, fn = 0;
  for (int fp = 0; fp < num_nolakeurbanp; ++fp) {
    const int p = nolakeurbanp[fp];
    if (frac_veg_nosno[p - bounds->begp] > 0) { exposedvegp[fe] = p; fe = fe + 1; }
    else { noexposedvegp[fn] = p; fn = fn + 1; }
  }
  *num_exposedvegp = fe;
  *num_noexposedvegp = fn;
}

/* Wet_BulbS, HumanIndexMod.F90:987-1030 (Stull 2011); pinned by the reference's own vectors (HumanStress_test/test_humanstress.pf:26-29) */
double oracle_wet_bulbs(double tc, double rh) {
  return tc * atan(0.151977 * sqrt(rh + 8.313659)) + atan(tc + rh) - atan(rh - 1.676331)
         + 0.00391838 * pow(rh, (3.0 / 2.0)) * atan(0.023101 * rh) - 4.686035;
}

int oracle_canopyfluxes(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_exposedvegp,
                        const int32_t* filter_exposedvegp, const ctsm_canopyfluxes_fields_t* fld, ctsm_status_t* st) {
  cf_ctx ctx, *x = &ctx;
  memset(&ctx, 0, sizeof ctx);
  ctx.f = fld; ctx.prm = prm;
  ctx.begp0 = fld->alloc.begp; ctx.begc0 = fld->alloc.begc; ctx.begg0 = fld->alloc.begg;
  ctx.ldp = (size_t)(fld->alloc.endp - fld->alloc.begp + 1);
  ctx.ldc = (size_t)(fld->alloc.endc - fld->alloc.begc + 1);
  ctx.np = (int)ctx.ldp;
  if (st) memset(st, 0, sizeof *st);

  const double btran0 = 0.0, zii = 1000.0, beta = 1.0, delmax = 1.0, dlemin = 0.1, dtmin = 0.01, ria = 0.5;
  const int itmin = 2;
  const double k_vert = 0.1, k_cyl_vol = 1.0, k_cyl_area = 1.0, k_internal = 0.0, min_stem_diameter = 0.05, min_lai = 0.1;
  const double dtime = prm->dtime;
  const int np = ctx.np, o = ctx.begp0;
  const int begp = bounds->begp, endp = bounds->endp;

#define NEWA(name) double* name = (double*)calloc((size_t)np, sizeof(double))
  NEWA(zldis); NEWA(dth); NEWA(dthv); NEWA(dqh); NEWA(ur); NEWA(temp1); NEWA(temp12m); NEWA(temp2); NEWA(temp22m);
  NEWA(rb); NEWA(rah_a); NEWA(rah_b); NEWA(raw_a); NEWA(raw_b); NEWA(wtg); NEWA(wta0); NEWA(wtl0); NEWA(wtstem0);
  NEWA(wtal); NEWA(wtga); NEWA(wtgq); NEWA(wtaq0); NEWA(wtlq0); NEWA(wtalq); NEWA(el); NEWA(qsatl); NEWA(qsatldT);
  NEWA(air); NEWA(bir); NEWA(cir); NEWA(delq); NEWA(del); NEWA(del2); NEWA(dele); NEWA(det); NEWA(efeb); NEWA(efe);
  NEWA(obuold); NEWA(tlbef); NEWA(tl_ini); NEWA(ts_ini); NEWA(err); NEWA(co2); NEWA(o2); NEWA(svpts); NEWA(eah);
  NEWA(fm); NEWA(dayl_factor); NEWA(dt_veg); NEWA(dbh); NEWA(cp_leaf); NEWA(cp_stem); NEWA(rstem); NEWA(dt_stem);
  NEWA(frac_rad_abs_by_stem); NEWA(lw_stem); NEWA(lw_leaf); NEWA(sa_stem); NEWA(sa_leaf); NEWA(sa_internal); NEWA(uuc);
  NEWA(snocan_baseline); NEWA(bbb_); NEWA(mbb_);
  int* nmozsgn = (int*)calloc((size_t)np, sizeof(int));
  int32_t* filterp = (int32_t*)calloc((size_t)np + 1, sizeof(int32_t));
  int32_t* fporig = (int32_t*)calloc((size_t)np + 1, sizeof(int32_t));
  int32_t* filterc_tmp = (int32_t*)calloc((size_t)np + 1, sizeof(int32_t));
  ctx.bbb = bbb_; ctx.mbb = mbb_;
#define A(arr, p) arr[(p) - o]
  /* the associate-d pointers of :638-641 */
  double* bsun = fld->bsun; double* bsha = fld->bsha; double* btran = fld->btran;

  int fn = num_exposedvegp;                                            /* :656-657 */
  for (int f = 0; f < fn; ++f) filterp[f] = filter_exposedvegp[f];

  oracle_photosyns_timestepinit(x, bounds);                            /* :663 */

  for (int f = 0; f < fn; ++f) {                                       /* :700-712 */
    const int p = filterp[f];
    A(del, p) = 0.0; A(efeb, p) = 0.0; A(wtlq0, p) = 0.0; A(wtalq, p) = 0.0; A(wtgq, p) = 0.0; A(wtaq0, p) = 0.0;
    A(obuold, p) = 0.0; A(btran, p) = btran0; P1(dhsdt_canopy, p) = 0.0; P1(eflx_sh_stem, p) = 0.0;
  }
  if (prm->use_biomass_heat_storage) {                                 /* :716-805 */
    for (int f = 0; f < fn; ++f) {
      const int p = filterp[f], ivt = P1(itype, p);
      const double elai = P1(elai, p), esai = P1(esai, p), htop = P1(htop, p), fbw = PFT(pft_fbw, ivt);
      A(frac_rad_abs_by_stem, p) = (esai) / (elai + esai);
      if (elai > 0.0) A(frac_rad_abs_by_stem, p) = k_vert * A(frac_rad_abs_by_stem, p);
      A(dbh, p) = PFT(pft_dbh, ivt);
      A(sa_leaf, p) = elai;
      A(sa_leaf, p) = 2.0 * A(sa_leaf, p);
      A(sa_stem, p) = PFT(pft_nstem, ivt) * (htop * rpi * A(dbh, p));
      A(sa_stem, p) = k_cyl_area * A(sa_stem, p);
      if (!(PFT(pft_is_tree, ivt) || PFT(pft_is_shrub, ivt)) || A(dbh, p) < min_stem_diameter) {
        A(frac_rad_abs_by_stem, p) = 0.0;
        A(sa_stem, p) = 0.0;
        A(sa_leaf, p) = A(sa_leaf, p) + esai;
      } else {
        if (elai < min_lai) A(sa_leaf, p) = A(sa_leaf, p) + esai;
      }
      P1(leaf_biomass, p) = (1.e-3 * c_to_b / PFT(pft_slatop, ivt)) * fmax(0.01, 0.5 * A(sa_leaf, p)) / (1.0 - fbw);
      const double carea_stem = rpi * ((A(dbh, p) * 0.5) * (A(dbh, p) * 0.5));
      P1(stem_biomass, p) = carea_stem * htop * k_cyl_vol * PFT(pft_nstem, ivt) * PFT(pft_wood_density, ivt) / (1.0 - fbw);
      A(sa_internal, p) = fmin(A(sa_leaf, p), A(sa_stem, p));
      A(sa_internal, p) = k_internal * A(sa_internal, p);
      A(cp_leaf, p) = P1(leaf_biomass, p) * (c_dry_biomass * (1.0 - fbw) + (fbw)*c_water);
      A(cp_stem, p) = P1(stem_biomass, p) * (c_dry_biomass * (1.0 - fbw) + (fbw)*c_water);
      A(cp_stem, p) = k_cyl_vol * A(cp_stem, p);
      A(rstem, p) = PFT(pft_rstem_per_dbh, ivt) * A(dbh, p);
    }
  } else {                                                             /* :806-818 */
    for (int f = 0; f < fn; ++f) {
      const int p = filterp[f];
      A(sa_leaf, p) = (P1(elai, p) + P1(esai, p));
      A(frac_rad_abs_by_stem, p) = 0.0; A(sa_stem, p) = 0.0; A(sa_internal, p) = 0.0;
      A(cp_leaf, p) = 0.0; A(cp_stem, p) = 0.0; A(rstem, p) = 0.0;
    }
  }
  for (int f = 0; f < fn; ++f) {                                       /* :822-828 */
    const int p = filterp[f], g = P1(gridcell, p);
    A(dayl_factor, p) = fmin(1.0, fmax(0.01, (G1(dayl, g) * G1(dayl, g)) / (G1(max_dayl, g) * G1(max_dayl, g))));
  }
  for (int p = begp; p <= endp; ++p) P1(rb1, p) = 0.0;                 /* :830 */
  for (int f = 0; f < fn; ++f) filterc_tmp[f] = P1(column, filterp[f]);   /* :833-836 */

  /* calc_effective_soilporosity, SoilMoistStressMod.F90:104-113 */
  for (int j = 1; j <= NLEVGRND; ++j)
    for (int fc = 0; fc < fn; ++fc) {
      const int c = filterc_tmp[fc];
      const double vol_ice = fmin(C2(watsat, c, j, 1), C2(h2osoi_ice, c, j, SNOSOI_LO) / (denice * C2(dz, c, j, SNOSOI_LO)));
      C2(eff_porosity, c, j, 1) = C2(watsat, c, j, 1) - vol_ice;
    }
  /* calc_volumetric_h2oliq with jtop = 1, :205-215 */
  for (int j = 1; j <= NLEVGRND; ++j)
    for (int fc = 0; fc < fn; ++fc) {
      const int c = filterc_tmp[fc];
      C2(h2osoi_liqvol, c, j, SNOSOI_LO) =
          fmin(C2(eff_porosity, c, j, 1), C2(h2osoi_liq, c, j, SNOSOI_LO) / (C2(dz, c, j, SNOSOI_LO) * denh2o));
    }
  /* calc_root_moist_stress -> calc_root_moist_stress_clm45default, :377-431 (perchroot off) */
  for (int j = 1; j <= NLEVGRND; ++j)
    for (int f = 0; f < fn; ++f) {
      const int p = filterp[f], c = P1(column, p), ivt = P1(itype, p);
      if (C2(h2osoi_liqvol, c, j, SNOSOI_LO) <= 0.0 || C2(t_soisno, c, j, SNOSOI_LO) <= tfrz - 2.0) {
        P2(rootr, p, j, 1) = 0.0;
      } else {
        const double s_node = fmax(C2(h2osoi_liqvol, c, j, SNOSOI_LO) / C2(eff_porosity, c, j, 1), 0.01);
        double smp_node = -C2(sucsat, c, j, 1) * pow(s_node, -C2(bsw, c, j, 1));      /* soil_suction */
        smp_node = fmax(PFT(pft_smpsc, ivt), smp_node);
        P2(rresis, p, j, 1) = fmin((C2(eff_porosity, c, j, 1) / C2(watsat, c, j, 1)) * (smp_node - PFT(pft_smpsc, ivt)) /
                                       (PFT(pft_smpso, ivt) - PFT(pft_smpsc, ivt)), 1.0);
        P2(rootr, p, j, 1) = P2(rootfr, p, j, 1) * P2(rresis, p, j, 1);
        A(btran, p) = A(btran, p) + fmax(P2(rootr, p, j, 1), 0.0);
      }
    }
  for (int j = 1; j <= NLEVGRND; ++j)
    for (int f = 0; f < fn; ++f) {
      const int p = filterp[f];
      if (A(btran, p) > btran0) P2(rootr, p, j, 1) = P2(rootr, p, j, 1) / A(btran, p);
      else P2(rootr, p, j, 1) = 0.0;
    }

  for (int f = 0; f < fn; ++f) {                                       /* :900-948 */
    const int p = filterp[f], c = P1(column, p), g = P1(gridcell, p), ivt = P1(itype, p);
    double lt;
    if (prm->z0param_method == 1) {
      lt = fmin(P1(elai, p) + P1(esai, p), tlsai_crit);
      const double egvf = (1.0 - alpha_aero * exp(-lt)) / (1.0 - alpha_aero * exp(-tlsai_crit));
      P1(displa, p) = egvf * P1(displa, p);
      P1(z0mv, p) = exp(egvf * log(P1(z0mv, p)) + (1.0 - egvf) * log(C1(z0mg, c)));
    } else {
      lt = fmax(1.e-5, P1(elai, p) + P1(esai, p));
      P1(displa, p) = P1(htop, p) * (1.0 - (1.0 - exp(-pow(cd1_param * lt, 0.5))) / pow(cd1_param * lt, 0.5));
      lt = fmin(lt, PFT(pft_z0v_LAImax, ivt));
      double delt = 2.0;
      const double U_ustar_ini = pow(PFT(pft_z0v_Cs, ivt) + PFT(pft_z0v_Cr, ivt) * lt * 0.5, -0.5) * PFT(pft_z0v_c, ivt) * lt * 0.25;
      double U_ustar = U_ustar_ini;
      while (delt > 1.e-4) {
        const double U_ustar_prev = U_ustar;
        U_ustar = U_ustar_ini * exp(U_ustar_prev);
        delt = fabs(U_ustar - U_ustar_prev);
      }
      U_ustar = 4.0 * U_ustar / lt / PFT(pft_z0v_c, ivt);
      P1(z0mv, p) = P1(htop, p) * (1.0 - P1(displa, p) / P1(htop, p)) *
                    exp(-vkc * U_ustar + log(PFT(pft_z0v_cw, ivt)) - 1.0 + 1.0 / PFT(pft_z0v_cw, ivt));
    }
    P1(z0hv, p) = P1(z0mv, p);
    P1(z0qv, p) = P1(z0mv, p);
    P1(forc_hgt_u_patch, p) = G1(forc_hgt_u, g) + P1(z0mv, p) + P1(displa, p);
    P1(forc_hgt_t_patch, p) = G1(forc_hgt_t, g) + P1(z0hv, p) + P1(displa, p);
    P1(forc_hgt_q_patch, p) = G1(forc_hgt_q, g) + P1(z0qv, p) + P1(displa, p);
  }

  int found = 0, index = 0;
  for (int f = 0; f < fn; ++f) {                                       /* :951-995 */
    const int p = filterp[f], c = P1(column, p), g = P1(gridcell, p);
    const double emv = P1(emv, p), emg = C1(emg, c);
    A(air, p) = emv * (1.0 + (1.0 - emv) * (1.0 - emg)) * C1(forc_lwrad, c);
    A(bir, p) = -(2.0 - emv * (1.0 - emg)) * emv * sb;
    A(cir, p) = emv * emg * sb;
    oracle_qsat(P1(t_veg, p), C1(forc_pbot, c), &A(qsatl, p), &A(el, p), &A(qsatldT, p));
    A(co2, p) = G1(forc_pco2, g);
    A(o2, p) = G1(forc_po2, g);
    nmozsgn[p - o] = 0;
    P1(taf, p) = (C1(t_grnd, c) + P1(thm, p)) / 2.0;
    P1(qaf, p) = (C1(forc_q, c) + C1(qg, c)) / 2.0;
    A(ur, p) = fmax(prm->wind_min, sqrt(G1(forc_u, g) * G1(forc_u, g) + G1(forc_v, g) * G1(forc_v, g)));
    A(dth, p) = P1(thm, p) - P1(taf, p);
    A(dqh, p) = C1(forc_q, c) - P1(qaf, p);
    A(delq, p) = C1(qg, c) - P1(qaf, p);
    A(dthv, p) = A(dth, p) * (1.0 + 0.61 * C1(forc_q, c)) + 0.61 * C1(forc_th, c) * A(dqh, p);
    A(zldis, p) = P1(forc_hgt_u_patch, p) - P1(displa, p);
    if (A(zldis, p) < 0.0) { found = 1; index = p; }
  }
  if (found) {                                                         /* :997-1002 */
    ctx.err_code = CTSM_ERR_FORC_HGT; ctx.err_index = index;
    goto done;
  }
  for (int f = 0; f < fn; ++f) {                                       /* :1004-1017 */
    const int p = filterp[f], c = P1(column, p);
    oracle_moninobukini(prm->zetamaxstable, A(ur, p), C1(thv, c), A(dthv, p), A(zldis, p), P1(z0mv, p), &P1(um, p), &P1(obu, p));
    P1(num_iter, p) = 0.0;
    A(tl_ini, p) = P1(t_veg, p);
    A(ts_ini, p) = P1(t_stem, p);
  }

  int itlef = 0;
  const int fnorig = fn;
  for (int f = 0; f < fn; ++f) fporig[f] = filterp[f];

  while (itlef <= prm->itmax_canopy_fluxes && fn > 0) {                /* :1028 ITERATION */
    FrictionVelocity(x, fn, filterp, fld->displa, fld->z0mv, fld->z0hv, fld->z0qv, fld->obu, itlef + 1, ur, fld->um,
                     fld->ustar, temp1, temp2, temp12m, temp22m, fm);
    for (int f = 0; f < fn; ++f) {                                     /* :1038-1122 */
      const int p = filterp[f], c = P1(column, p), ivt = P1(itype, p);
      A(tlbef, p) = P1(t_veg, p);
      A(del2, p) = A(del, p);
      P1(ram1, p) = 1.0 / (P1(ustar, p) * P1(ustar, p) / P1(um, p));
      A(rah_a, p) = 1.0 / (A(temp1, p) * P1(ustar, p));
      A(raw_a, p) = 1.0 / (A(temp2, p) * P1(ustar, p));
      P1(uaf, p) = P1(um, p) * sqrt(1.0 / (P1(ram1, p) * P1(um, p)));
      A(uuc, p) = fmin(0.4, (0.03 * P1(um, p) / P1(ustar, p)));
      P1(dleaf_patch, p) = PFT(pft_dleaf, ivt);
      const double cf = prm->cv / (sqrt(P1(uaf, p)) * sqrt(P1(dleaf_patch, p)));
      A(rb, p) = 1.0 / (cf * P1(uaf, p));
      P1(rb1, p) = A(rb, p);
      const double w = exp(-(P1(elai, p) + P1(esai, p)));
      const double csoilb = vkc / (prm->a_coef * pow(C1(z0mg, c) * P1(uaf, p) / nu_param, prm->a_exp));
      const double ri = (grav * P1(htop, p) * (P1(taf, p) - C1(t_grnd, c))) / (P1(taf, p) * (P1(uaf, p) * P1(uaf, p)));
      double csoilcn;
      if (prm->use_undercanopy_stability && (P1(taf, p) - C1(t_grnd, c)) > 0.0) {
        const double ricsoilc = prm->csoilc / (1.00 + ria * fmin(ri, 10.0));
        csoilcn = csoilb * w + ricsoilc * (1.0 - w);
      } else {
        csoilcn = csoilb * w + prm->csoilc * (1.0 - w);
      }
      if (prm->use_biomass_heat_storage) A(rah_b, p) = 1.0 / (csoilcn * A(uuc, p));
      else A(rah_b, p) = 1.0 / (csoilcn * P1(uaf, p));
      A(raw_b, p) = A(rah_b, p);
      A(svpts, p) = A(el, p);
      A(eah, p) = C1(forc_pbot, c) * P1(qaf, p) / 0.622;
      P1(rh_af, p) = A(eah, p) / A(svpts, p);
      P1(rah1, p) = A(rah_a, p);
      P1(raw1, p) = A(raw_a, p);
      P1(rah2, p) = A(rah_b, p);
      P1(raw2, p) = A(raw_b, p);
      P1(vpd, p) = fmax((A(svpts, p) - A(eah, p)), 50.0) * 0.001;
    }

    if (prm->use_hydrstress) {
      oracle_photosynthesis_hydraulic_stress(x, fn, filterp, svpts, eah, o2, co2, rb, bsun, bsha, btran, dayl_factor, qsatl,
                                             fld->qaf);   /* :1134-1141 */
    } else {                                              /* :1143-1166: sunlit, then shaded leaves */
      oracle_photosynthesis(x, fn, filterp, svpts, eah, o2, co2, rb, btran, dayl_factor, 0);
      oracle_photosynthesis(x, fn, filterp, svpts, eah, o2, co2, rb, btran, dayl_factor, 1);
    }
    if (ctx.err_code) goto done;

    for (int f = 0; f < fn; ++f) {                                     /* :1174-1435 */
      const int p = filterp[f], c = P1(column, p);
      const double forc_rho = C1(forc_rho, c), forc_q = C1(forc_q, c), t_grnd = C1(t_grnd, c), thm = P1(thm, p);
      const double elai = P1(elai, p), esai = P1(esai, p), emv = P1(emv, p);
      const double fvn = (double)P1(frac_veg_nosno, p);
      const double wta = 1.0 / A(rah_a, p);
      const double wtl = A(sa_leaf, p) / A(rb, p);
      A(wtg, p) = 1.0 / A(rah_b, p);
      const double wtstem = A(sa_stem, p) / (A(rstem, p) + A(rb, p));
      const double wtshi = 1.0 / (wta + wtl + wtstem + A(wtg, p));
      A(wtl0, p) = wtl * wtshi;
      const double wtg0 = A(wtg, p) * wtshi;
      A(wta0, p) = wta * wtshi;
      A(wtstem0, p) = wtstem * wtshi;
      A(wtga, p) = A(wta0, p) + wtg0 + A(wtstem0, p);
      A(wtal, p) = A(wta0, p) + A(wtl0, p) + A(wtstem0, p);
      const double tv = P1(t_veg, p), tstem = P1(t_stem, p);
      A(lw_stem, p) = A(sa_internal, p) * emv * sb * ((tstem * tstem) * (tstem * tstem));
      A(lw_leaf, p) = A(sa_internal, p) * emv * sb * ((tv * tv) * (tv * tv));
      double rppdry;
      if (P1(fdry, p) > 0.0)
        rppdry = P1(fdry, p) * A(rb, p) * (P1(laisun, p) / (A(rb, p) + P1(rssun, p)) + P1(laisha, p) / (A(rb, p) + P1(rssha, p))) / elai;
      else
        rppdry = 0.0;
      double efpot = forc_rho * ((elai + esai) / A(rb, p)) * (A(qsatl, p) - P1(qaf, p));
      const double h2ocan = P1(liqcan, p) + P1(snocan, p);
      double rpp;
      if (prm->use_hydrstress) {
        if (efpot > 0.0) {                                             /* use_hydrstress branch :1219-1230 */
          if (A(btran, p) > btran0) rpp = rppdry + P1(fwet, p);
          else rpp = P1(fwet, p);
          rpp = fmin(rpp, (P1(qflx_tran_veg, p) + h2ocan / dtime) / efpot);
        } else {
          rpp = 1.0;
        }
      } else {                                                         /* :1231-1248: transpiration follows the potential */
        if (efpot > 0.0) {
          if (A(btran, p) > btran0) {
            P1(qflx_tran_veg, p) = efpot * rppdry;
            rpp = rppdry + P1(fwet, p);
          } else {
            rpp = P1(fwet, p);
            P1(qflx_tran_veg, p) = 0.0;
          }
          rpp = fmin(rpp, (P1(qflx_tran_veg, p) + h2ocan / dtime) / efpot);
        } else {
          rpp = 1.0;
          P1(qflx_tran_veg, p) = 0.0;
        }
      }
      const double wtaq = fvn / A(raw_a, p);
      const double wtlq = fvn * (elai + esai) / A(rb, p) * rpp;
      const double snow_depth_c = prm->z_dl;
      const double fsno_dl = C1(snow_depth, c) / snow_depth_c;
      const double elai_dl = prm->lai_dl * (1.0 - fmin(fsno_dl, 1.0));
      const double rdl = (1.0 - exp(-elai_dl)) / (0.004 * P1(uaf, p));
      if (A(delq, p) < 0.0) {
        A(wtgq, p) = fvn / (A(raw_b, p) + rdl);
      } else {
        if (prm->soil_resis_method == 0) A(wtgq, p) = C1(soilbeta, c) * fvn / (A(raw_b, p) + rdl);
        if (prm->soil_resis_method == 1) A(wtgq, p) = fvn / (A(raw_b, p) + C1(soilresis, c));
      }
      const double wtsqi = 1.0 / (wtaq + wtlq + A(wtgq, p));
      const double wtgq0 = A(wtgq, p) * wtsqi;
      A(wtlq0, p) = wtlq * wtsqi;
      A(wtaq0, p) = wtaq * wtsqi;
      const double wtgaq = A(wtaq0, p) + wtgq0;
      A(wtalq, p) = A(wtaq0, p) + A(wtlq0, p);
      const double dc1 = forc_rho * cpair * wtl;
      const double dc2 = hvap * forc_rho * wtlq;
      const double efsh = dc1 * (A(wtga, p) * tv - wtg0 * t_grnd - A(wta0, p) * thm - A(wtstem0, p) * tstem);
      P1(eflx_sh_stem, p) = forc_rho * cpair * wtstem *
                            ((A(wta0, p) + wtg0 + A(wtl0, p)) * tstem - wtg0 * t_grnd - A(wta0, p) * thm - A(wtl0, p) * tv);
      A(efe, p) = dc2 * (wtgaq * A(qsatl, p) - wtgq0 * C1(qg, c) - A(wtaq0, p) * forc_q);
      double erre = 0.0;
      if (A(efe, p) * A(efeb, p) < 0.0) {
        const double efeold = A(efe, p);
        A(efe, p) = 0.1 * efeold;
        erre = A(efe, p) - efeold;
      }
      const int snl = C1(snl, c);
      const double tsn = C2(t_soisno, c, snl + 1, SNOSOI_LO), ts1 = C2(t_soisno, c, 1, SNOSOI_LO), th2o = C1(t_h2osfc, c);
      const double frac_sno = C1(frac_sno_eff, c), frac_h2osfc = C1(frac_h2osfc, c);
      const double lw_grnd = (frac_sno * ((tsn * tsn) * (tsn * tsn)) + (1.0 - frac_sno - frac_h2osfc) * ((ts1 * ts1) * (ts1 * ts1))
                              + frac_h2osfc * ((th2o * th2o) * (th2o * th2o)));
      const double frs = A(frac_rad_abs_by_stem, p);
      const double tv3 = (tv * tv) * tv, tv4 = (tv * tv) * (tv * tv);
      A(dt_veg, p) = ((1.0 - frs) * (P1(sabv, p) + A(air, p) + A(bir, p) * tv4 + A(cir, p) * lw_grnd)
                      - efsh - A(efe, p) - A(lw_leaf, p) + A(lw_stem, p) - (A(cp_leaf, p) / dtime) * (tv - A(tl_ini, p)))
                     / ((1.0 - frs) * (-4.0 * A(bir, p) * tv3) + 4.0 * A(sa_internal, p) * emv * sb * tv3
                        + dc1 * A(wtga, p) + dc2 * wtgaq * A(qsatldT, p) + A(cp_leaf, p) / dtime);
      P1(t_veg, p) = A(tlbef, p) + A(dt_veg, p);
      const double dels = A(dt_veg, p);
      A(del, p) = fabs(dels);
      A(err, p) = 0.0;
      const double tb = A(tlbef, p), tb3 = (tb * tb) * tb;
      if (A(del, p) > delmax) {
        A(dt_veg, p) = delmax * dels / A(del, p);
        P1(t_veg, p) = A(tlbef, p) + A(dt_veg, p);
        A(err, p) = (1.0 - frs) * (P1(sabv, p) + A(air, p) + A(bir, p) * tb3 * (tb + 4.0 * A(dt_veg, p)) + A(cir, p) * lw_grnd)
                    - A(sa_internal, p) * emv * sb * tb3 * (tb + 4.0 * A(dt_veg, p)) + A(lw_stem, p)
                    - (efsh + dc1 * A(wtga, p) * A(dt_veg, p)) - (A(efe, p) + dc2 * wtgaq * A(qsatldT, p) * A(dt_veg, p))
                    - (A(cp_leaf, p) / dtime) * (P1(t_veg, p) - A(tl_ini, p));
      }
      efpot = forc_rho * ((elai + esai) / A(rb, p)) *
              (wtgaq * (A(qsatl, p) + A(qsatldT, p) * A(dt_veg, p)) - wtgq0 * C1(qg, c) - A(wtaq0, p) * forc_q);
      P1(qflx_evap_veg, p) = rpp * efpot;
      if (!prm->use_hydrstress) {                                      /* :1357-1362 */
        if (efpot > 0.0 && A(btran, p) > btran0) P1(qflx_tran_veg, p) = efpot * rppdry;
        else P1(qflx_tran_veg, p) = 0.0;
      }
      const double ecidif = fmax(0.0, P1(qflx_evap_veg, p) - P1(qflx_tran_veg, p) - h2ocan / dtime);   /* :1353-1355, :1363 */
      P1(qflx_evap_veg, p) = fmin(P1(qflx_evap_veg, p), P1(qflx_tran_veg, p) + h2ocan / dtime);
      P1(eflx_sh_veg, p) = efsh + dc1 * A(wtga, p) * A(dt_veg, p) + A(err, p) + erre + hvap * ecidif;
      P1(eflx_sh_stem, p) = P1(eflx_sh_stem, p) + forc_rho * cpair * wtstem * (-A(wtl0, p) * A(dt_veg, p));
      A(lw_leaf, p) = A(sa_internal, p) * emv * sb * tb3 * (tb + 4.0 * A(dt_veg, p));
      oracle_qsat(P1(t_veg, p), C1(forc_pbot, c), &A(qsatl, p), &A(el, p), &A(qsatldT, p));
      P1(taf, p) = wtg0 * t_grnd + A(wta0, p) * thm + A(wtl0, p) * P1(t_veg, p) + A(wtstem0, p) * tstem;
      P1(qaf, p) = A(wtlq0, p) * A(qsatl, p) + wtgq0 * C1(qg, c) + forc_q * A(wtaq0, p);
      A(dth, p) = thm - P1(taf, p);
      A(dqh, p) = forc_q - P1(qaf, p);
      A(delq, p) = A(wtalq, p) * C1(qg, c) - A(wtlq0, p) * A(qsatl, p) - A(wtaq0, p) * forc_q;
      const double tstar = A(temp1, p) * A(dth, p);
      const double qstar = A(temp2, p) * A(dqh, p);
      const double thvstar = tstar * (1.0 + 0.61 * forc_q) + 0.61 * C1(forc_th, c) * qstar;
      P1(zeta, p) = A(zldis, p) * vkc * grav * thvstar / ((P1(ustar, p) * P1(ustar, p)) * C1(thv, c));
      if (P1(zeta, p) >= 0.0) {
        P1(zeta, p) = fmin(prm->zetamaxstable, fmax(P1(zeta, p), 0.01));
        P1(um, p) = fmax(A(ur, p), 0.1);
      } else {
        P1(zeta, p) = fmax(-100.0, fmin(P1(zeta, p), -0.01));
        double wc;
        if (P1(ustar, p) * thvstar > 0.0) { wc = 0.0; ctx.n_warnings++; }
        else wc = beta * pow(-grav * P1(ustar, p) * thvstar * zii / C1(thv, c), 0.333);
        P1(um, p) = sqrt(A(ur, p) * A(ur, p) + wc * wc);
      }
      P1(obu, p) = A(zldis, p) / P1(zeta, p);
      if (A(obuold, p) * P1(obu, p) < 0.0) nmozsgn[p - o] = nmozsgn[p - o] + 1;
      if (nmozsgn[p - o] >= 4) P1(obu, p) = A(zldis, p) / (-0.01);
      A(obuold, p) = P1(obu, p);
    }

    itlef = itlef + 1;                                                 /* :1439-1457 */
    if (itlef > itmin) {
      for (int f = 0; f < fn; ++f) {
        const int p = filterp[f];
        A(dele, p) = fabs(A(efe, p) - A(efeb, p));
        A(efeb, p) = A(efe, p);
        A(det, p) = fmax(A(del, p), A(del2, p));
        P1(num_iter, p) = (double)itlef;
      }
      const int fnold = fn;
      fn = 0;
      for (int f = 0; f < fnold; ++f) {
        const int p = filterp[f];
        if (!(A(det, p) < dtmin && A(dele, p) < dlemin)) { filterp[fn] = p; fn = fn + 1; }
      }
    }
  }

  fn = fnorig;                                                         /* :1461-1462 */
  for (int f = 0; f < fn; ++f) filterp[f] = fporig[f];

  for (int f = 0; f < fn; ++f) {                                       /* :1464-1634 */
    const int p = filterp[f], c = P1(column, p), g = P1(gridcell, p);
    const double forc_rho = C1(forc_rho, c), forc_q = C1(forc_q, c), t_grnd = C1(t_grnd, c), thm = P1(thm, p);
    const double emv = P1(emv, p), emg = C1(emg, c), forc_lwrad = C1(forc_lwrad, c);
    const int snl = C1(snl, c);
    const double tsn = C2(t_soisno, c, snl + 1, SNOSOI_LO), ts1 = C2(t_soisno, c, 1, SNOSOI_LO), th2o = C1(t_h2osfc, c);
    const double frac_sno = C1(frac_sno_eff, c), frac_h2osfc = C1(frac_h2osfc, c);
    const double lw_grnd = (frac_sno * ((tsn * tsn) * (tsn * tsn)) + (1.0 - frac_sno - frac_h2osfc) * ((ts1 * ts1) * (ts1 * ts1))
                            + frac_h2osfc * ((th2o * th2o) * (th2o * th2o)));
    const double frs = A(frac_rad_abs_by_stem, p);
    const double tb = A(tlbef, p), tb3 = (tb * tb) * tb;
    const double tsi = A(ts_ini, p), tsi3 = (tsi * tsi) * tsi, tsi4 = (tsi * tsi) * (tsi * tsi);
    A(err, p) = (1.0 - frs) * (P1(sabv, p) + A(air, p) + A(bir, p) * tb3 * (tb + 4.0 * A(dt_veg, p)) + A(cir, p) * lw_grnd)
                - A(lw_leaf, p) + A(lw_stem, p) - P1(eflx_sh_veg, p) - hvap * P1(qflx_evap_veg, p)
                - ((P1(t_veg, p) - A(tl_ini, p)) * A(cp_leaf, p) / dtime);
    if (prm->use_biomass_heat_storage) {
      if (P1(stem_biomass, p) > 0.0) {
        A(dt_stem, p) = (frs * (P1(sabv, p) + A(air, p) + A(bir, p) * tsi4 + A(cir, p) * lw_grnd) - P1(eflx_sh_stem, p)
                         + A(lw_leaf, p) - A(lw_stem, p)) / (A(cp_stem, p) / dtime - frs * A(bir, p) * 4.0 * tsi3);
      } else {
        A(dt_stem, p) = 0.0;
      }
      P1(dhsdt_canopy, p) = A(dt_stem, p) * A(cp_stem, p) / dtime + (P1(t_veg, p) - A(tl_ini, p)) * A(cp_leaf, p) / dtime;
      P1(t_stem, p) = P1(t_stem, p) + A(dt_stem, p);
    } else {
      A(dt_stem, p) = 0.0;
    }
    const double tv = P1(t_veg, p), tstem = P1(t_stem, p);
    const double delt = A(wtal, p) * t_grnd - A(wtl0, p) * tv - A(wta0, p) * thm - A(wtstem0, p) * tstem;
    P1(taux, p) = -forc_rho * G1(forc_u, g) / P1(ram1, p);
    P1(tauy, p) = -forc_rho * G1(forc_v, g) / P1(ram1, p);
    P1(eflx_sh_grnd, p) = cpair * forc_rho * A(wtg, p) * delt;
    const double delt_snow = A(wtal, p) * tsn - A(wtl0, p) * tv - A(wta0, p) * thm - A(wtstem0, p) * tstem;
    const double delt_soil = A(wtal, p) * ts1 - A(wtl0, p) * tv - A(wta0, p) * thm - A(wtstem0, p) * tstem;
    const double delt_h2osfc = A(wtal, p) * th2o - A(wtl0, p) * tv - A(wta0, p) * thm - A(wtstem0, p) * tstem;
    P1(eflx_sh_snow, p) = cpair * forc_rho * A(wtg, p) * delt_snow;
    P1(eflx_sh_soil, p) = cpair * forc_rho * A(wtg, p) * delt_soil;
    P1(eflx_sh_h2osfc, p) = cpair * forc_rho * A(wtg, p) * delt_h2osfc;
    P1(qflx_evap_soi, p) = forc_rho * A(wtgq, p) * A(delq, p);
    const double delq_snow = A(wtalq, p) * C1(qg_snow, c) - A(wtlq0, p) * A(qsatl, p) - A(wtaq0, p) * forc_q;
    P1(qflx_ev_snow, p) = forc_rho * A(wtgq, p) * delq_snow;
    const double delq_soil = A(wtalq, p) * C1(qg_soil, c) - A(wtlq0, p) * A(qsatl, p) - A(wtaq0, p) * forc_q;
    P1(qflx_ev_soil, p) = forc_rho * A(wtgq, p) * delq_soil;
    const double delq_h2osfc = A(wtalq, p) * C1(qg_h2osfc, c) - A(wtlq0, p) * A(qsatl, p) - A(wtaq0, p) * forc_q;
    P1(qflx_ev_h2osfc, p) = forc_rho * A(wtgq, p) * delq_h2osfc;
    P1(t_ref2m, p) = thm + A(temp1, p) * A(dth, p) * (1.0 / A(temp12m, p) - 1.0 / A(temp1, p));
    P1(t_ref2m_r, p) = P1(t_ref2m, p);
    P1(q_ref2m, p) = forc_q + A(temp2, p) * A(dqh, p) * (1.0 / A(temp22m, p) - 1.0 / A(temp2, p));
    double qsat_ref2m, e_ref2m;
    oracle_qsat(P1(t_ref2m, p), C1(forc_pbot, c), &qsat_ref2m, &e_ref2m, NULL);
    P1(rh_ref2m, p) = fmin(100.0, P1(q_ref2m, p) / qsat_ref2m * 100.0);
    P1(rh_ref2m_r, p) = P1(rh_ref2m, p);
    P1(vpd_ref2m, p) = e_ref2m * (1.0 - P1(rh_ref2m, p) / 100.0);
    if (prm->calc_human_stress_indices == 1) {                          /* fast_human_stress_indices :1550-1570 */
      const double rh = P1(rh_ref2m, p);
      const double tc = P1(t_ref2m, p) - tfrz;                          /* KtoC, HumanIndexMod.F90:1205 */
      P1(tc_ref2m, p) = tc;
      const double vap = (rh / 100.0) * e_ref2m;                        /* VaporPres :1243 */
      P1(vap_ref2m, p) = vap;
      if ((rh < 0.0 || rh > 100.0) && !ctx.err_code) { ctx.err_code = CTSM_ERR_RH; ctx.err_index = p; }   /* Wet_BulbS :1016-1022 endrun */
      const double wbt = oracle_wet_bulbs(tc, rh);                      /* :1024-1027 */
      P1(wbt_ref2m, p) = wbt;
      const double tf = (tc) * 9.0 / 5.0 + 32.0;                        /* HeatIndex :1039-1095 */
      double hi;
      if (tf < 68.0) hi = tf;
      else hi = -42.379 + 2.04901523 * tf + 10.14333127 * rh + (-0.22475541 * tf * rh) + (-6.83783e-3 * (tf * tf))
                + (-5.481717e-2 * (rh * rh)) + 1.22874e-3 * (tf * tf) * rh + 8.5282e-4 * tf * (rh * rh)
                + (-1.99e-6 * (tf * tf) * (rh * rh));
      hi = (hi - 32.0) * 5.0 / 9.0;
      P1(nws_hi_ref2m, p) = hi;
      P1(appar_temp_ref2m, p) = tc + 3.30 * vap / 1000.0 - 0.70 * P1(u10_clm, p) - 4.0;               /* AppTemp :555 */
      P1(swbgt_ref2m, p) = 0.567 * (tc) + 0.393 * vap / 100.0 + 3.94;                                 /* swbgt :596 */
      P1(humidex_ref2m, p) = tc + ((5.0 / 9.0) * (vap / 100.0 - 10.0));                               /* hmdex :637 */
      {                                                                                               /* dis_coiS :715-761 */
        const double Tc = fmin(tc, 50.0);
        double rhl = fmin(rh, 99.0);
        rhl = fmax(rhl, 5.0);
        const double rh_min = Tc * (-2.27) + 27.7;
        if (Tc < -20.0 || rhl < rh_min) P1(discomf_index_ref2mS, p) = Tc;
        else P1(discomf_index_ref2mS, p) = 0.5 * wbt + 0.5 * Tc;
      }
      P1(wbt_ref2m_r, p) = P1(wbt_ref2m, p);
      P1(nws_hi_ref2m_r, p) = P1(nws_hi_ref2m, p);
      P1(appar_temp_ref2m_r, p) = P1(appar_temp_ref2m, p);
      P1(swbgt_ref2m_r, p) = P1(swbgt_ref2m, p);
      P1(humidex_ref2m_r, p) = P1(humidex_ref2m, p);
      P1(discomf_index_ref2mS_r, p) = P1(discomf_index_ref2mS, p);
    }
    P1(dlrad, p) = (1.0 - emv) * emg * forc_lwrad + emv * emg * sb * tb3 * (tb + 4.0 * A(dt_veg, p)) * (1.0 - frs)
                   + emv * emg * sb * tsi3 * (tsi + 4.0 * A(dt_stem, p)) * frs;
    P1(ulrad, p) = ((1.0 - emg) * (1.0 - emv) * (1.0 - emv) * forc_lwrad
                    + emv * (1.0 + (1.0 - emg) * (1.0 - emv)) * sb * tb3 * (tb + 4.0 * A(dt_veg, p)) * (1.0 - frs)
                    + emv * (1.0 + (1.0 - emg) * (1.0 - emv)) * sb * tsi3 * (tsi + 4.0 * A(dt_stem, p)) * frs
                    + emg * (1.0 - emv) * sb * lw_grnd);
    P1(t_skin, p) = emv * tv + (1.0 - emv) * sqrt(sqrt(lw_grnd));
    P1(cgrnds, p) = P1(cgrnds, p) + cpair * forc_rho * A(wtg, p) * A(wtal, p);
    P1(cgrndl, p) = P1(cgrndl, p) + forc_rho * A(wtgq, p) * A(wtalq, p) * C1(dqgdT, c);
    P1(cgrnd, p) = P1(cgrnds, p) + P1(cgrndl, p) * C1(htvp, c);
    A(snocan_baseline, p) = P1(snocan, p);
    if (tv > tfrz) {                                                   /* :1618-1632 */
      if ((P1(qflx_evap_veg, p) - P1(qflx_tran_veg, p)) * dtime > P1(liqcan, p))
        P1(snocan, p) = fmax(0.0, P1(snocan, p) + P1(liqcan, p) + (P1(qflx_tran_veg, p) - P1(qflx_evap_veg, p)) * dtime);
      P1(liqcan, p) = fmax(0.0, P1(liqcan, p) + (P1(qflx_tran_veg, p) - P1(qflx_evap_veg, p)) * dtime);
    } else if (tv <= tfrz) {
      if ((P1(qflx_evap_veg, p) - P1(qflx_tran_veg, p)) * dtime > P1(snocan, p))
        P1(liqcan, p) = P1(liqcan, p) + P1(snocan, p) + (P1(qflx_tran_veg, p) - P1(qflx_evap_veg, p)) * dtime;
      P1(snocan, p) = fmax(0.0, P1(snocan, p) + (P1(qflx_tran_veg, p) - P1(qflx_evap_veg, p)) * dtime);
    }
  }
  /* truncate_small_values(custom_rel_epsilon = 1e-10), :1639-1641 */
  for (int f = 0; f < fn; ++f) {
    const int p = filterp[f];
    if (fabs(P1(snocan, p)) < 1.e-10 * fabs(A(snocan_baseline, p))) P1(snocan, p) = 0.0;
  }
  oracle_photosynthesis_total(x, fn, filterp);                         /* :1655 */
  for (int f = 0; f < fn; ++f) {                                       /* :1660-1676 */
    const int p = filterp[f], g = P1(gridcell, p);
    if (G1(near_local_noon, g) && P1(fpsn, p) > 0.0) {
      const double gs = 1.e-6 * (P1(laisun, p) * P2(gs_mol_sun, p, 1, 1) + P1(laisha, p) * P2(gs_mol_sha, p, 1, 1));
      if (gs > 0.0) P1(iwue_ln, p) = P1(fpsn, p) / gs;
      else P1(iwue_ln, p) = spval;
    } else {
      P1(iwue_ln, p) = spval;
    }
  }
  if (prm->use_luna) {                                                 /* Acc24_Climate_LUNA, LunaMod.F90:695-724 (call :1704) */
    for (int f = 0; f < fn; ++f) {
      const int p = filterp[f];
      if (P1(t_veg_day, p) != spval) {                                 /* not the first day */
        if (P1(sabv, p) > 0) {
          P1(t_veg_day, p) = P1(t_veg_day, p) + P1(t_veg, p);
          P1(ndaysteps, p) = P1(ndaysteps, p) + 1;
        } else {
          P1(t_veg_night, p) = P1(t_veg_night, p) + P1(t_veg, p);
          P1(nnightsteps, p) = P1(nnightsteps, p) + 1;
        }
        for (int z = 1; z <= P1(nrad, p); ++z) {
          const double tlaii = P2(laisun_z, p, z, 1) + P2(laisha_z, p, z, 1);
          if (tlaii > 0.0) {
            const double TRad = P2(parsun_z, p, z, 1);                 /* :715 overrides the lai-weighted mean of :714 */
            P2(par24d_z, p, z, 1) = P2(par24d_z, p, z, 1) + dtime * TRad;
            if (TRad > P2(par24x_z, p, z, 1)) P2(par24x_z, p, z, 1) = TRad;
          }
        }
        P1(fpsn24, p) = P1(fpsn24, p) + dtime * P1(fpsn, p);
      }
    }
  }
  for (int f = 0; f < fn; ++f) {                                       /* :1746-1760 */
    const int p = filterp[f];
    if (fabs(A(err, p)) > 0.1) ctx.n_warnings++;
  }

done:
  if (st) {
    st->code = ctx.err_code;
    st->subgrid_index = ctx.err_index;
    st->subgrid_level = ctx.err_code ? CTSM_SUBGRID_PATCH : 0;
    st->n_warnings = ctx.n_warnings;
  }
  free(zldis); free(dth); free(dthv); free(dqh); free(ur); free(temp1); free(temp12m); free(temp2); free(temp22m);
  free(rb); free(rah_a); free(rah_b); free(raw_a); free(raw_b); free(wtg); free(wta0); free(wtl0); free(wtstem0);
  free(wtal); free(wtga); free(wtgq); free(wtaq0); free(wtlq0); free(wtalq); free(el); free(qsatl); free(qsatldT);
  free(air); free(bir); free(cir); free(delq); free(del); free(del2); free(dele); free(det); free(efeb); free(efe);
  free(obuold); free(tlbef); free(tl_ini); free(ts_ini); free(err); free(co2); free(o2); free(svpts); free(eah);
  free(fm); free(dayl_factor); free(dt_veg); free(dbh); free(cp_leaf); free(cp_stem); free(rstem); free(dt_stem);
  free(frac_rad_abs_by_stem); free(lw_stem); free(lw_leaf); free(sa_stem); free(sa_leaf); free(sa_internal); free(uuc);
  free(snocan_baseline); free(bbb_); free(mbb_); free(nmozsgn); free(filterp); free(fporig); free(filterc_tmp);
  return ctx.err_code;
#undef A
#undef NEWA
}

/* truncate_small_values, src/utils/NumericsMod.F90:50-99 (data_baseline/data are (lb:ub)) */
void oracle_truncate_small_values(int num_f, const int32_t* filter_f, int lb, const double* data_baseline, double* data,
                                  double rel_epsilon) {
  for (int fn = 0; fn < num_f; ++fn) {
    const int n = filter_f[fn];
    if (fabs(data[n - lb]) < rel_epsilon * fabs(data_baseline[n - lb])) data[n - lb] = 0.0;
  }
}

/* BalanceCheckInit, src/biogeophys/BalanceCheckMod.F90:74-95: skip_steps = max(2, nint(3600/dtime)) + 1 */
int oracle_balancecheck_skip_steps(double dtime) {
  const double skip_size = 3600.0;
  int n = (int)lround(skip_size / dtime);
  if (n < 2) n = 2;
  return n + 1;
}
