/* oracle_preflux.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the three routines clm_drv runs immediately before CanopyFluxes:
 *   BiogeophysPreFluxCalcs      src/biogeophys/BiogeophysPreFluxCalcsMod.F90:58-406 (clm_driver.F90:680)
 *     SetZ0mDisp :120-219, SetRoughnessLengthsAndForcHeightsNonLake FrictionVelocityMod.F90:543-685,
 *     CalcInitialTemperatureAndEnergyVars :223-405, calc_soilevap_resis SurfaceResistanceMod.F90:192-426
 *   CalculateSurfaceHumidity    src/biogeophys/SurfaceHumidityMod.F90:41-239 (clm_driver.F90:702)
 *   BareGroundFluxes            src/biogeophys/BareGroundFluxesMod.F90:63-579 (clm_driver.F90:711)
 * Configuration: non-urban landunits, use_fates = use_lch4 = .false.  Loops and statement order follow the Fortran.
 * Default-kind REAL literals of the source (SurfaceResistanceMod.F90:398-401) are rounded through float like gfortran does.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_canopy.h"
#include "oracle_pert.h"
#undef P1
#undef P2
#undef C1
#undef C2
#undef G1

static const double vkc = 0.4, grav = 9.80616, cpair = 1.00464e3, hvap = 2.501e6, hsub = 2.501e6 + 3.337e5;
static const double denh2o = 1.000e3, denice = 0.917e3, spval = 1.e36;
static const double roverg = 6.02214e26 * 1.38065e-23 / 18.016 / 9.80616 * 1000.0;      /* clm_varcon.F90:  rwat/grav*1000 */
static const double cd1_param = 7.5, nu_param = 1.5e-5, beta_param = 7.2, b1_param = 1.4, b4_param = -0.31;
static const double meier_param1 = 0.23, meier_param2 = 0.08, meier_param3 = 70.0;
static const double rpi_ = 3.14159265358979323846;
#define R4(x) ((double)(float)(x))

#define ISTSOIL CTSM_ISTSOIL
#define ISTCROP CTSM_ISTCROP
#define ISTICE CTSM_ISTICE
#define ISTWET CTSM_ISTWET

static int is_urban(int lt) { return lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX; }

/* ------------------------------------------------------------------------------------------------------------------- */
int oracle_biogeophys_pre_flux_calcs(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_nolakec,
                                     const int32_t* filter_nolakec, int num_nolakep, const int32_t* filter_nolakep,
                                     int num_urbanc, int time_flags, const ctsm_preflux_fields_t* f, ctsm_status_t* st) {
  (void)bounds;
  if (st) memset(st, 0, sizeof *st);
  const int begc0 = f->alloc.begc, begp0 = f->alloc.begp, begg0 = f->alloc.begg;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1);
#define CC(name, c) f->name[(c) - begc0]
#define C2(name, c, j, lo) f->name[(size_t)((j) - (lo)) * ldc + ((c) - begc0)]
#define PP(name, p) f->name[(p) - begp0]
#define GG(name, g) f->name[(g) - begg0]
  if (num_urbanc != 0) { if (st) st->code = CTSM_ERR_URBAN; return CTSM_ERR_URBAN; }
  for (int fc = 0; fc < num_nolakec; ++fc)
    if (is_urban(CC(lun_itype, filter_nolakec[fc]))) {
      if (st) { st->code = CTSM_ERR_URBAN; st->subgrid_index = filter_nolakec[fc]; }
      return CTSM_ERR_URBAN;
    }

  /* SetZ0mDisp :120-219 */
  for (int fp = 0; fp < num_nolakep; ++fp) {
    const int p = filter_nolakep[fp], c = PP(column, p), ivt = PP(itype, p);
    if (prm->z0param_method == 1) {                                            /* ZengWang2007 */
      PP(z0m, p) = f->pft_z0mr[ivt] * PP(htop, p);
      PP(displa, p) = f->pft_displar[ivt] * PP(htop, p);
    } else {                                                                   /* Meier2022 */
      if (time_flags & CTSM_TIME_FIRST_STEPS) {
        PP(z0m, p) = 0.0;
        PP(displa, p) = 0.0;
        continue;
      } else if ((time_flags & CTSM_TIME_BEG_CURR_YEAR) && f->pft_crop[ivt] != 0.0) {
        PP(z0m, p) = 0.0;
        PP(displa, p) = 0.0;
      }
      if (ivt == 0) {                                                          /* noveg */
        PP(z0m, p) = 0.0;
        PP(displa, p) = 0.0;
      } else {
        const double lm = f->pft_z0v_LAImax[ivt];
        PP(displa, p) = PP(htop, p) * (1.0 - (1.0 - exp(-pow(cd1_param * lm, 0.5))) / pow(cd1_param * lm, 0.5));
        const double U_ustar = 4.0 * pow(f->pft_z0v_Cs[ivt] + f->pft_z0v_Cr[ivt] * lm / 2.0, -0.5) / lm / f->pft_z0v_c[ivt];
        if (PP(htop, p) <= 1.e-10) {
          PP(z0m, p) = CC(z0mg, c);
        } else {
          PP(z0m, p) = PP(htop, p) * (1.0 - PP(displa, p) / PP(htop, p)) *
                       exp(-0.4 * U_ustar + log(f->pft_z0v_cw[ivt]) - 1.0 + 1.0 / f->pft_z0v_cw[ivt]);   /* x**(-1._r8): gcc folds pow(x,-1) to 1/x */
        }
      }
    }
  }

  /* SetRoughnessLengthsAndForcHeightsNonLake, FrictionVelocityMod.F90:601-681 */
  for (int fc = 0; fc < num_nolakec; ++fc) {
    const int c = filter_nolakec[fc];
    if (prm->z0param_method == 1) {
      if (CC(frac_sno, c) > 0.0) CC(z0mg, c) = prm->zsno;
      else CC(z0mg, c) = prm->zlnd;
    } else {
      if (CC(frac_sno, c) > 0.0) {
        if (prm->use_z0m_snowmelt) {
          if (CC(snomelt_accum, c) < 1.e-5) CC(z0mg, c) = exp(-b1_param * rpi_ * 0.5 + b4_param) * 1.e-3;
          else CC(z0mg, c) = exp(b1_param * (atan((log10(CC(snomelt_accum, c)) + meier_param1) / meier_param2)) + b4_param) * 1.e-3;
        } else {
          CC(z0mg, c) = prm->zsno;
        }
      } else if (CC(lun_itype, c) == ISTICE) {
        CC(z0mg, c) = prm->zglc;
      } else {
        CC(z0mg, c) = prm->zlnd;
      }
    }
    CC(z0hg, c) = CC(z0mg, c);
    CC(z0qg, c) = CC(z0mg, c);
  }
  for (int fp = 0; fp < num_nolakep; ++fp) {
    const int p = filter_nolakep[fp];
    PP(z0mv, p) = PP(z0m, p);
    PP(z0hv, p) = PP(z0mv, p);
    PP(z0qv, p) = PP(z0mv, p);
    PP(z0mg_p, p) = spval;
    PP(z0hg_p, p) = spval;
    PP(z0qg_p, p) = spval;
    PP(kbm1, p) = spval;
  }
  for (int fp = 0; fp < num_nolakep; ++fp) {
    const int p = filter_nolakep[fp], g = PP(gridcell, p), c = PP(column, p), lt = CC(lun_itype, c);
    if (lt == ISTSOIL || lt == ISTCROP) {
      if (PP(frac_veg_nosno, p) == 0) {
        PP(forc_hgt_u_patch, p) = GG(forc_hgt_u, g) + CC(z0mg, c) + PP(displa, p);
        PP(forc_hgt_t_patch, p) = GG(forc_hgt_t, g) + CC(z0hg, c) + PP(displa, p);
        PP(forc_hgt_q_patch, p) = GG(forc_hgt_q, g) + CC(z0qg, c) + PP(displa, p);
      } else {
        PP(forc_hgt_u_patch, p) = GG(forc_hgt_u, g) + PP(z0mv, p) + PP(displa, p);
        PP(forc_hgt_t_patch, p) = GG(forc_hgt_t, g) + PP(z0hv, p) + PP(displa, p);
        PP(forc_hgt_q_patch, p) = GG(forc_hgt_q, g) + PP(z0qv, p) + PP(displa, p);
      }
    } else if (lt == ISTWET || lt == ISTICE) {
      PP(forc_hgt_u_patch, p) = GG(forc_hgt_u, g) + CC(z0mg, c) + PP(displa, p);
      PP(forc_hgt_t_patch, p) = GG(forc_hgt_t, g) + CC(z0hg, c) + PP(displa, p);
      PP(forc_hgt_q_patch, p) = GG(forc_hgt_q, g) + CC(z0qg, c) + PP(displa, p);
    }
  }

  /* CalcInitialTemperatureAndEnergyVars :304-401 */
  for (int j = -CTSM_NLEVSNO + 1; j <= CTSM_NLEVGRND; ++j)
    for (int fc = 0; fc < num_nolakec; ++fc) {
      const int c = filter_nolakec[fc];
      C2(t_ssbef, c, j, -CTSM_NLEVSNO + 1) = C2(t_soisno, c, j, -CTSM_NLEVSNO + 1);
    }
  for (int fc = 0; fc < num_nolakec; ++fc) {
    const int c = filter_nolakec[fc], snl = CC(snl, c);
    CC(t_h2osfc_bef, c) = CC(t_h2osfc, c);
    if (snl < 0) {
      CC(t_grnd, c) = CC(frac_sno_eff, c) * C2(t_soisno, c, snl + 1, -CTSM_NLEVSNO + 1)
                      + (1.0 - CC(frac_sno_eff, c) - CC(frac_h2osfc, c)) * C2(t_soisno, c, 1, -CTSM_NLEVSNO + 1)
                      + CC(frac_h2osfc, c) * CC(t_h2osfc, c);
    } else {
      CC(t_grnd, c) = (1 - CC(frac_h2osfc, c)) * C2(t_soisno, c, 1, -CTSM_NLEVSNO + 1) + CC(frac_h2osfc, c) * CC(t_h2osfc, c);
    }
    if (CC(lun_itype, c) == ISTICE) CC(emg, c) = 0.97;
    else CC(emg, c) = (1.0 - CC(frac_sno, c)) * 0.96 + CC(frac_sno, c) * 0.97;
    if (C2(h2osoi_liq, c, snl + 1, -CTSM_NLEVSNO + 1) <= 0.0 && C2(h2osoi_ice, c, snl + 1, -CTSM_NLEVSNO + 1) > 0.0) CC(htvp, c) = hsub;
    else CC(htvp, c) = hvap;
    CC(beta, c) = 1.0;
    CC(zii, c) = 1000.0;
    CC(thv, c) = CC(forc_th, c) * (1.0 + 0.61 * CC(forc_q, c));
  }
  for (int fp = 0; fp < num_nolakep; ++fp) {
    const int p = filter_nolakep[fp], c = PP(column, p), lt = CC(lun_itype, c);
    PP(eflx_sh_tot, p) = 0.0;
    if (lt == ISTSOIL || lt == ISTCROP) PP(eflx_sh_tot_r, p) = 0.0;
    PP(eflx_lh_tot, p) = 0.0;
    if (lt == ISTSOIL || lt == ISTCROP) PP(eflx_lh_tot_r, p) = 0.0;
    PP(eflx_sh_veg, p) = 0.0;
    PP(cgrnd, p) = 0.0;
    PP(cgrnds, p) = 0.0;
    PP(cgrndl, p) = 0.0;
    const double avmuir = 1.0;
    PP(emv, p) = 1.0 - exp(-(PP(elai, p) + PP(esai, p)) / avmuir);
    PP(thm, p) = CC(forc_t, c) + 0.0098 * PP(forc_hgt_t_patch, p);
  }

  /* calc_soilevap_resis, SurfaceResistanceMod.F90:192-426 */
  for (int fc = 0; fc < num_nolakec; ++fc) {
    const int c = filter_nolakec[fc], lt = CC(lun_itype, c);
    const double liq1 = C2(h2osoi_liq, c, 1, -CTSM_NLEVSNO + 1), ice1 = C2(h2osoi_ice, c, 1, -CTSM_NLEVSNO + 1);
    const double dz1 = C2(dz, c, 1, -CTSM_NLEVSNO + 1), watsat1 = C2(watsat, c, 1, 1);
    if (prm->soil_resis_method == 0) {                                         /* calc_beta_leepielke1992 :279-313 */
      if (lt != ISTWET && lt != ISTICE) {
        if (lt == ISTSOIL || lt == ISTCROP) {
          const double wx = (liq1 / denh2o + ice1 / denice) / dz1;
          const double watfc1 = C2(watfc, c, 1, 1);
          if (wx < watfc1) {
            double fac_fc = fmin(1.0, wx / watfc1);
            fac_fc = fmax(fac_fc, 0.01);
            const double t = (1.0 - cos(rpi_ * fac_fc));
            CC(soilbeta, c) = (1.0 - CC(frac_sno, c) - CC(frac_h2osfc, c)) * 0.25 * (t * t) + CC(frac_sno, c) + CC(frac_h2osfc, c);
          } else {
            CC(soilbeta, c) = 1.0;
          }
        }
      } else {
        CC(soilbeta, c) = 1.0;
      }
    } else {                                                                   /* calc_soil_resistance_sl14 :388-424 */
      if (lt != ISTWET && lt != ISTICE) {
        if (lt == ISTSOIL || lt == ISTCROP) {
          const double bsw1 = C2(bsw, c, 1, 1), sucsat1 = C2(sucsat, c, 1, 1);
          const double vwc_liq = fmax(liq1, 1.0e-6) / (dz1 * denh2o);
          const double eff_por_top = fmax(0.01, watsat1 - fmin(watsat1, ice1 / (dz1 * denice)));
          const double aird = watsat1 * pow(sucsat1 / 1.e7, R4(1.) / bsw1);
          const double d0 = R4(2.12e-5) * pow(C2(t_soisno, c, 1, -CTSM_NLEVSNO + 1) / R4(273.15), R4(1.75));
          const double eps = watsat1 - aird;
          const double dg = eps * d0 * pow(eps / watsat1, 3.0 / fmax(3.0, bsw1));
          double dsl = prm->d_max * fmax(0.001, (prm->frac_sat_soil_dsl_init * eff_por_top - vwc_liq))
                       / fmax(0.001, (prm->frac_sat_soil_dsl_init * watsat1 - aird));
          dsl = fmax(dsl, 0.0);
          dsl = fmin(dsl, 200.0);
          CC(dsl, c) = dsl;
          double sr = dsl / (dg * eps * R4(1.e3)) + 20.0;
          sr = fmin(1.e6, sr);
          CC(soilresis, c) = sr;
        }
      } else {
        CC(soilresis, c) = 0.0;
      }
    }
  }
  return 0;
#undef CC
#undef C2
#undef PP
#undef GG
}

/* ------------------------------------------------------------------------------------------------------------------- */
int oracle_calculate_surface_humidity(const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                                      const ctsm_surfacehumidity_fields_t* f, ctsm_status_t* st) {
  (void)bounds;
  if (st) memset(st, 0, sizeof *st);
  const int begc0 = f->alloc.begc;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1);
#define CC(name, c) f->name[(c) - begc0]
#define C2(name, c, j, lo) f->name[(size_t)((j) - (lo)) * ldc + ((c) - begc0)]
  for (int fc = 0; fc < num_nolakec; ++fc) {
    const int c = filter_nolakec[fc], lt = CC(lun_itype, c), snl = CC(snl, c);
    if (is_urban(lt)) { if (st) { st->code = CTSM_ERR_URBAN; st->subgrid_index = c; } return CTSM_ERR_URBAN; }
    double qred = 1.0, hr = 0.0, qsatg, qsatgdT, qsatgdT_snow, qsatgdT_soil, qsatgdT_h2osfc;
    const double t1 = C2(t_soisno, c, 1, -CTSM_NLEVSNO + 1);
    if (lt != ISTWET && lt != ISTICE) {
      if (lt == ISTSOIL || lt == ISTCROP) {                                    /* :125-135 */
        const double wx = (C2(h2osoi_liq, c, 1, -CTSM_NLEVSNO + 1) / denh2o + C2(h2osoi_ice, c, 1, -CTSM_NLEVSNO + 1) / denice)
                          / C2(dz, c, 1, -CTSM_NLEVSNO + 1);
        double fac = fmin(1.0, wx / C2(watsat, c, 1, 1));
        fac = fmax(fac, 0.01);
        double psit = -C2(sucsat, c, 1, 1) * pow(fac, -C2(bsw, c, 1, 1));
        psit = fmax(CC(smpmin, c), psit);
        hr = exp(psit / roverg / t1);
        qred = (1.0 - CC(frac_sno_eff, c) - CC(frac_h2osfc, c)) * hr + CC(frac_sno_eff, c) + CC(frac_h2osfc, c);
        CC(soilalpha, c) = qred;
      }
    } else {
      CC(soilalpha, c) = spval;
    }
    if (lt == ISTSOIL || lt == ISTCROP) {                                      /* :179-216 */
      oracle_qsat(t1, CC(forc_pbot, c), &qsatg, NULL, &qsatgdT_soil);
      if (qsatg > CC(forc_q, c) && CC(forc_q, c) > hr * qsatg) {
        qsatg = CC(forc_q, c);
        qsatgdT_soil = 0.0;
      }
      CC(qg_soil, c) = hr * qsatg;
      if (snl < 0) {
        oracle_qsat(C2(t_soisno, c, snl + 1, -CTSM_NLEVSNO + 1), CC(forc_pbot, c), &qsatg, NULL, &qsatgdT_snow);
        CC(qg_snow, c) = qsatg;
        CC(dqgdT, c) = CC(frac_sno_eff, c) * qsatgdT_snow + (1.0 - CC(frac_sno_eff, c) - CC(frac_h2osfc, c)) * hr * qsatgdT_soil;
      } else {
        CC(qg_snow, c) = CC(qg_soil, c);
        CC(dqgdT, c) = (1.0 - CC(frac_h2osfc, c)) * hr * qsatgdT_soil;
      }
      if (CC(frac_h2osfc, c) > 0.0) {
        oracle_qsat(CC(t_h2osfc, c), CC(forc_pbot, c), &qsatg, NULL, &qsatgdT_h2osfc);
        CC(qg_h2osfc, c) = qsatg;
        CC(dqgdT, c) = CC(dqgdT, c) + CC(frac_h2osfc, c) * qsatgdT_h2osfc;
      } else {
        CC(qg_h2osfc, c) = CC(qg_soil, c);
      }
      CC(qg, c) = CC(frac_sno_eff, c) * CC(qg_snow, c) + (1.0 - CC(frac_sno_eff, c) - CC(frac_h2osfc, c)) * CC(qg_soil, c)
                  + CC(frac_h2osfc, c) * CC(qg_h2osfc, c);
    } else {                                                                   /* :218-234 */
      oracle_qsat(CC(t_grnd, c), CC(forc_pbot, c), &qsatg, NULL, &qsatgdT);
      CC(qg, c) = qred * qsatg;
      CC(dqgdT, c) = qred * qsatgdT;
      if (qsatg > CC(forc_q, c) && CC(forc_q, c) > qred * qsatg) {
        CC(qg, c) = CC(forc_q, c);
        CC(dqgdT, c) = 0.0;
      }
      CC(qg_snow, c) = CC(qg, c);
      CC(qg_soil, c) = CC(qg, c);
      CC(qg_h2osfc, c) = CC(qg, c);
    }
  }
  return 0;
#undef CC
#undef C2
}

/* ------------------------------------------------------------------------------------------------------------------- */
/* dewpoint, BareGroundFluxesMod.F90:531-577 */
static double dewpoint(double e, double t) {
  const double A1_liq = 17.625, B1_liq = 243.04, C1_liq = 610.94, A1_ice = 22.587, B1_ice = 273.86, C1_ice = 611.21;
  double d;
  if (t < tfrz) d = B1_ice * log(e / C1_ice) / (A1_ice - log(e / C1_ice));
  else d = B1_liq * log(e / C1_liq) / (A1_liq - log(e / C1_liq));
  return d + tfrz;
}

int oracle_bare_ground_fluxes(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_noexposedvegp,
                              const int32_t* filter_noexposedvegp, const ctsm_baregroundfluxes_fields_t* f, ctsm_status_t* st) {
  (void)bounds;
  if (st) memset(st, 0, sizeof *st);
  const int niters = 3;                                                        /* :82 */
  const int begc0 = f->alloc.begc, begp0 = f->alloc.begp, begg0 = f->alloc.begg;
  const size_t ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1), ldp = (size_t)(f->alloc.endp - f->alloc.begp + 1);
  const int np = f->alloc.endp - f->alloc.begp + 1;
#define CC(name, c) f->name[(c) - begc0]
#define C2(name, c, j, lo) f->name[(size_t)((j) - (lo)) * ldc + ((c) - begc0)]
#define PP(name, p) f->name[(p) - begp0]
#define P2(name, p, j, lo) f->name[(size_t)((j) - (lo)) * ldp + ((p) - begp0)]
#define GG(name, g) f->name[(g) - begg0]
#define A(arr, p) arr[(p) - begp0]
  double* wk = (double*)calloc((size_t)(np > 0 ? np : 1) * 9, sizeof(double));
  double *ur = wk, *dth = wk + np, *dqh = wk + 2 * (size_t)np, *zldis = wk + 3 * (size_t)np, *temp1 = wk + 4 * (size_t)np,
         *temp2 = wk + 5 * (size_t)np, *temp12m = wk + 6 * (size_t)np, *temp22m = wk + 7 * (size_t)np, *fm = wk + 8 * (size_t)np;
  int rc = 0;
  for (int fi = 0; fi < num_noexposedvegp; ++fi) {                             /* :281-294 */
    const int p = filter_noexposedvegp[fi], c = PP(column, p);
    if (is_urban(CC(lun_itype, c))) { rc = CTSM_ERR_URBAN; if (st) { st->code = rc; st->subgrid_index = p; } goto done; }
    PP(btran, p) = 0.0;
    PP(t_veg, p) = CC(forc_t, c);
    const double cf_bare = CC(forc_pbot, c) / (rgas * 0.001 * PP(thm, p)) * 1.e06;
    PP(rssun, p) = 1.0 / 1.e15 * cf_bare;
    PP(rssha, p) = 1.0 / 1.e15 * cf_bare;
    for (int j = 1; j <= CTSM_NLEVGRND; ++j) {
      P2(rootr, p, j, 1) = 0.0;
      P2(rresis, p, j, 1) = 0.0;
    }
  }
  for (int fi = 0; fi < num_noexposedvegp; ++fi) {                             /* :299-335 */
    const int p = filter_noexposedvegp[fi], c = PP(column, p), g = PP(gridcell, p);
    PP(displa, p) = 0.0;
    PP(z0mv, p) = 0.0;
    PP(z0hv, p) = 0.0;
    PP(z0qv, p) = 0.0;
    PP(dlrad, p) = 0.0;
    PP(ulrad, p) = 0.0;
    PP(dhsdt_canopy, p) = 0.0;
    PP(eflx_sh_stem, p) = 0.0;
    A(ur, p) = fmax(prm->wind_min, sqrt(GG(forc_u, g) * GG(forc_u, g) + GG(forc_v, g) * GG(forc_v, g)));
    A(dth, p) = PP(thm, p) - CC(t_grnd, c);
    A(dqh, p) = CC(forc_q, c) - CC(qg, c);
    const double dthv = A(dth, p) * (1.0 + 0.61 * CC(forc_q, c)) + 0.61 * CC(forc_th, c) * A(dqh, p);
    A(zldis, p) = PP(forc_hgt_u_patch, p);
    PP(z0mg_p, p) = CC(z0mg, c);
    PP(z0hg_p, p) = CC(z0hg, c);
    PP(z0qg_p, p) = CC(z0qg, c);
    double um, obu;
    oracle_moninobukini(prm->zetamaxstable, A(ur, p), CC(thv, c), dthv, A(zldis, p), PP(z0mg_p, p), &um, &obu);
    PP(um, p) = um;
    PP(obu, p) = obu;
    PP(num_iter, p) = 0.0;
  }
  for (int iter = 1; iter <= niters; ++iter) {                                 /* :341-398 */
    for (int fi = 0; fi < num_noexposedvegp; ++fi) {                           /* FrictionVelocity over the filter */
      const int p = filter_noexposedvegp[fi];
      oracle_fricvel_t r;
      r.fm = A(fm, p);
      oracle_friction_velocity_point(PP(forc_hgt_u_patch, p), PP(forc_hgt_t_patch, p), PP(forc_hgt_q_patch, p), PP(displa, p),
                                     PP(z0mg_p, p), PP(z0hg_p, p), PP(z0qg_p, p), PP(obu, p), iter, A(ur, p), PP(um, p), &r);
      PP(ustar, p) = r.ustar; A(temp1, p) = r.temp1; A(temp2, p) = r.temp2; A(temp12m, p) = r.temp12m; A(temp22m, p) = r.temp22m;
      A(fm, p) = r.fm;
      PP(vds, p) = r.vds; PP(u10_clm, p) = r.u10_clm; PP(va, p) = r.va; PP(u10, p) = r.u10; PP(fv, p) = r.fv;
    }
    for (int fi = 0; fi < num_noexposedvegp; ++fi) {
      const int p = filter_noexposedvegp[fi], c = PP(column, p), g = PP(gridcell, p);
      const double tstar = A(temp1, p) * A(dth, p);
      const double qstar = A(temp2, p) * A(dqh, p);
      if (prm->z0param_method == 1)
        PP(z0hg_p, p) = PP(z0mg_p, p) / exp(prm->a_coef * pow(PP(ustar, p) * PP(z0mg_p, p) / nu_param, prm->a_exp));
      else
        PP(z0hg_p, p) = meier_param3 * nu_param / PP(ustar, p) * exp(-beta_param * pow(PP(ustar, p), 0.5) * pow(fabs(tstar), 0.25));
      PP(z0qg_p, p) = PP(z0hg_p, p);
      PP(forc_hgt_u_patch, p) = GG(forc_hgt_u, g) + PP(z0mg_p, p) + PP(displa, p);
      PP(forc_hgt_t_patch, p) = GG(forc_hgt_t, g) + PP(z0hg_p, p) + PP(displa, p);
      PP(forc_hgt_q_patch, p) = GG(forc_hgt_q, g) + PP(z0qg_p, p) + PP(displa, p);
      const double thvstar = tstar * (1.0 + 0.61 * CC(forc_q, c)) + 0.61 * CC(forc_th, c) * qstar;
      double zeta = A(zldis, p) * vkc * grav * thvstar / ((PP(ustar, p) * PP(ustar, p)) * CC(thv, c));
      if (zeta >= 0.0) {
        zeta = fmin(prm->zetamaxstable, fmax(zeta, 0.01));
        PP(um, p) = fmax(A(ur, p), 0.1);
      } else {
        zeta = fmax(-100.0, fmin(zeta, -0.01));
        const double wc = CC(beta, c) * pow(-grav * PP(ustar, p) * thvstar * CC(zii, c) / CC(thv, c), 0.333);
        PP(um, p) = sqrt(A(ur, p) * A(ur, p) + wc * wc);
      }
      PP(zeta, p) = zeta;
      PP(obu, p) = A(zldis, p) / zeta;
      PP(num_iter, p) = (double)iter;
    }
  }
  for (int fi = 0; fi < num_noexposedvegp; ++fi) {                             /* :402-525 */
    const int p = filter_noexposedvegp[fi], c = PP(column, p), g = PP(gridcell, p), lt = CC(lun_itype, c), snl = CC(snl, c);
    const double ram = 1.0 / (PP(ustar, p) * PP(ustar, p) / PP(um, p));
    const double rah = 1.0 / (A(temp1, p) * PP(ustar, p));
    const double raw = 1.0 / (A(temp2, p) * PP(ustar, p));
    const double raih = CC(forc_rho, c) * cpair / rah;
    double www = (C2(h2osoi_liq, c, 1, -CTSM_NLEVSNO + 1) / denh2o + C2(h2osoi_ice, c, 1, -CTSM_NLEVSNO + 1) / denice)
                 / C2(dz, c, 1, -CTSM_NLEVSNO + 1) / C2(watsat, c, 1, 1);
    www = fmin(fmax(www, 0.0), 1.0);
    (void)www;
    double forc_qs, forc_esat;
    oracle_qsat(CC(forc_t, c), CC(forc_pbot, c), &forc_qs, &forc_esat, NULL);
    const double forc_e = fmax((CC(forc_q, c) * CC(forc_pbot, c)) / (CC(forc_q, c) + 0.622), 0.01 * forc_esat);
    const double forc_dewpoint = dewpoint(forc_e, CC(t_grnd, c));
    double raiw = 0.0;
    if (A(dqh, p) > 0.0) {
      if (CC(t_grnd, c) > forc_dewpoint) raiw = 0.0;
      else raiw = CC(forc_rho, c) / (raw);
    } else {
      if (prm->soil_resis_method == 0) {
        if (CC(t_grnd, c) > forc_dewpoint) raiw = 0.0;
        else raiw = CC(soilbeta, c) * CC(forc_rho, c) / (raw);
      }
      if (prm->soil_resis_method == 1) raiw = CC(forc_rho, c) / (raw + CC(soilresis, c));
    }
    PP(ram1, p) = ram;
    PP(cgrnds, p) = raih;
    PP(cgrndl, p) = raiw * CC(dqgdT, c);
    PP(cgrnd, p) = PP(cgrnds, p) + CC(htvp, c) * PP(cgrndl, p);
    PP(taux, p) = -CC(forc_rho, c) * GG(forc_u, g) / ram;
    PP(tauy, p) = -CC(forc_rho, c) * GG(forc_v, g) / ram;
    PP(eflx_sh_grnd, p) = -raih * A(dth, p);
    PP(eflx_sh_tot, p) = PP(eflx_sh_grnd, p);
    PP(eflx_sh_snow, p) = -raih * (PP(thm, p) - C2(t_soisno, c, snl + 1, -CTSM_NLEVSNO + 1));
    PP(eflx_sh_soil, p) = -raih * (PP(thm, p) - C2(t_soisno, c, 1, -CTSM_NLEVSNO + 1));
    PP(eflx_sh_h2osfc, p) = -raih * (PP(thm, p) - CC(t_h2osfc, c));
    PP(qflx_tran_veg, p) = 0.0;
    PP(qflx_evap_veg, p) = 0.0;
    PP(qflx_evap_soi, p) = -raiw * A(dqh, p);
    PP(qflx_evap_tot_patch, p) = PP(qflx_evap_soi, p);
    PP(qflx_ev_snow, p) = -raiw * (CC(forc_q, c) - CC(qg_snow, c));
    PP(qflx_ev_soil, p) = -raiw * (CC(forc_q, c) - CC(qg_soil, c));
    PP(qflx_ev_h2osfc, p) = -raiw * (CC(forc_q, c) - CC(qg_h2osfc, c));
    PP(t_ref2m, p) = PP(thm, p) + A(temp1, p) * A(dth, p) * (1.0 / A(temp12m, p) - 1.0 / A(temp1, p));
    PP(q_ref2m, p) = CC(forc_q, c) + A(temp2, p) * A(dqh, p) * (1.0 / A(temp22m, p) - 1.0 / A(temp2, p));
    double qsat_ref2m, e_ref2m;
    oracle_qsat(PP(t_ref2m, p), CC(forc_pbot, c), &qsat_ref2m, &e_ref2m, NULL);
    PP(rh_ref2m, p) = fmin(100.0, PP(q_ref2m, p) / qsat_ref2m * 100.0);
    const int rural = (lt == ISTSOIL || lt == ISTCROP);
    if (rural) {
      PP(rh_ref2m_r, p) = PP(rh_ref2m, p);
      PP(t_ref2m_r, p) = PP(t_ref2m, p);
    }
    PP(kbm1, p) = log(PP(z0mg_p, p) / PP(z0hg_p, p));
    CC(z0hg, c) = PP(z0hg_p, p);
    CC(z0qg, c) = PP(z0qg_p, p);
    if (prm->calc_human_stress_indices == 1) {                                 /* fast indices :475-485, HumanIndexMod.F90 */
      const double rh = PP(rh_ref2m, p);
      const double tc = PP(t_ref2m, p) - tfrz;                                 /* KtoC :1205 */
      PP(tc_ref2m, p) = tc;
      const double vap = (rh / 100.0) * e_ref2m;                               /* VaporPres :1243 */
      PP(vap_ref2m, p) = vap;
      if (rh < 0.0 || rh > 100.0) { rc = CTSM_ERR_RH; if (st) { st->code = rc; st->subgrid_index = p; } goto done; }
      const double wbt = oracle_wet_bulbs(tc, rh);                             /* Wet_BulbS :1024-1027 */
      PP(wbt_ref2m, p) = wbt;
      const double tf = (tc) * 9.0 / 5.0 + 32.0;                               /* HeatIndex :1039-1095 */
      double hi;
      if (tf < 68.0) hi = tf;
      else hi = -42.379 + 2.04901523 * tf + 10.14333127 * rh + (-0.22475541 * tf * rh) + (-6.83783e-3 * (tf * tf))
                + (-5.481717e-2 * (rh * rh)) + 1.22874e-3 * (tf * tf) * rh + 8.5282e-4 * tf * (rh * rh)
                + (-1.99e-6 * (tf * tf) * (rh * rh));
      hi = (hi - 32.0) * 5.0 / 9.0;
      PP(nws_hi_ref2m, p) = hi;
      PP(appar_temp_ref2m, p) = tc + 3.30 * vap / 1000.0 - 0.70 * PP(u10_clm, p) - 4.0;             /* AppTemp :555 */
      PP(swbgt_ref2m, p) = 0.567 * (tc) + 0.393 * vap / 100.0 + 3.94;                              /* swbgt :596 */
      PP(humidex_ref2m, p) = tc + ((5.0 / 9.0) * (vap / 100.0 - 10.0));                            /* hmdex :637 */
      const double Tc = fmin(tc, 50.0);                                        /* dis_coiS :715-761 */
      double rhl = fmin(rh, 99.0);
      rhl = fmax(rhl, 5.0);
      const double rh_min = Tc * (-2.27) + 27.7;
      PP(discomf_index_ref2mS, p) = (Tc < -20.0 || rhl < rh_min) ? Tc : 0.5 * wbt + 0.5 * Tc;
      if (rural) {
        PP(wbt_ref2m_r, p) = PP(wbt_ref2m, p);
        PP(nws_hi_ref2m_r, p) = PP(nws_hi_ref2m, p);
        PP(appar_temp_ref2m_r, p) = PP(appar_temp_ref2m, p);
        PP(swbgt_ref2m_r, p) = PP(swbgt_ref2m, p);
        PP(humidex_ref2m_r, p) = PP(humidex_ref2m, p);
        PP(discomf_index_ref2mS_r, p) = PP(discomf_index_ref2mS, p);
      }
    }
  }
done:
  free(wk);
  return rc;
#undef CC
#undef C2
#undef PP
#undef P2
#undef GG
#undef A
}
