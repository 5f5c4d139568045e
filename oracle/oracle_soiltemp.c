/* oracle_soiltemp.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of SoilTemperature and its callees
 * (src/biogeophys/SoilTemperatureMod.F90: SoilTemperature :92-599, SoilThermProp
 * :602-901, PhaseChangeH2osfc :904-1130, Phasechange :1133-1540,
 * ComputeGroundHeatFluxAndDeriv :1543-1796, ComputeHeatDiffFluxAndFactor
 * :1799-1910, SetRHSVec* :1913-2353, SetMatrix* :2356-2926) for non-urban
 * columns (istsoil, istcrop, istice, istwet landunits), use_excess_ice=.false.
 * (excess_ice == 0 everywhere, so every `+ excess_ice` term is a bit-exact
 * no-op and is omitted).  Loop order is the Fortran's: level-outer /
 * filter-inner, clump-sized temporaries, one dgbsv per column.
 *
 * Un-suffixed REAL(4) literals are reproduced as (double)(float) constants
 * (SURVEY.md F9): 1.0e-3 at :352,:835 and the Sturm-1997 constants :753-775.
 *
 * PARITY UNPINNED by the reference's own tests; invariants
 * (steady state, energy conservation) in tests/test_oracle_soiltemp.py.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define NLEVSNO CTSM_NLEVSNO
#define NLEVGRND CTSM_NLEVGRND
#define NLEVSOI CTSM_NLEVSOI
#define SNOSOI_LO (-NLEVSNO + 1)
#define SNOSOI0_LO (-NLEVSNO)

/* shr_const_mod.F90:16-53, clm_varcon.F90:50-122 */
static const double tfrz = 273.15, denh2o = 1.000e3, denice = 0.917e3, cpliq = 4.188e3, cpice = 2.11727e3;
static const double hfus = 3.337e5, sb = 5.67e-8, grav = 9.80616;
static const double tkair = 0.023, tkice = 2.290, tkwat = 0.57, capr = 0.34, cnfac = 0.5;
static const double thk_bedrock = 3.0, csol_bedrock = 2.0e6;
static const double thin_sfclayer = 1.0e-6;   /* SoilTemperatureMod.F90:84 */
/* REAL(4) literals promoted to double (F9) */
#define R4(x) ((double)(float)(x))

typedef struct {
  const ctsm_soiltemperature_fields_t* f;
  const ctsm_params_t* prm;
  int begc0, begp0;       /* allocation lower bounds */
  size_t ldc, ldp;
  int begc, endc;         /* call bounds (size the temporaries) */
  size_t ldt;
  int numc; const int32_t* filterc;
  int nump; const int32_t* filterp;
  /* clump-sized temporaries, SoilTemperatureMod.F90:153-180 */
  double *cv, *tk, *fn, *fn1, *sabg_lyr_col;
  double *hs_top, *hs_soil, *hs_top_snow, *hs_h2osfc, *dhsdT, *fn_h2osfc, *dz_h2osfc, *tk_h2osfc;
  double *bmatrix, *tvector, *rvector;
  int32_t *jtop, *jbot;
} st_ctx;

/* boundary fields */
#define C1(name, c) (x->f->name[(c) - x->begc0])
#define C2(name, c, j, lo) (x->f->name[(size_t)((j) - (lo)) * x->ldc + ((c) - x->begc0)])
#define P1(name, p) (x->f->name[(p) - x->begp0])
#define P2(name, p, j, lo) (x->f->name[(size_t)((j) - (lo)) * x->ldp + ((p) - x->begp0)])
/* temporaries: (begc:endc, lo:...) */
#define T1(name, c) (x->name[(c) - x->begc])
#define T2(name, c, j, lo) (x->name[(size_t)((j) - (lo)) * x->ldt + ((c) - x->begc)])
#define BM(c, k, r) (x->bmatrix[((size_t)((r) - SNOSOI0_LO) * 5 + ((k) - 1)) * x->ldt + ((c) - x->begc)])

#define T_SOISNO(c, j) C2(t_soisno, c, j, SNOSOI_LO)
#define LIQ(c, j) C2(h2osoi_liq, c, j, SNOSOI_LO)
#define ICE(c, j) C2(h2osoi_ice, c, j, SNOSOI_LO)
#define DZ(c, j) C2(dz, c, j, SNOSOI_LO)
#define Z(c, j) C2(z, c, j, SNOSOI_LO)
#define ZI(c, j) C2(zi, c, j, SNOSOI0_LO)
#define FACT(c, j) C2(fact, c, j, SNOSOI_LO)
#define THK(c, j) C2(thk, c, j, SNOSOI_LO)

static int is_soil_or_crop(int lt) { return lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP; }

/* snow thermal conductivity, :750-781 */
static double snow_thk(int method, double bw) {
  if (method == 1) {   /* Jordan1991 */
    return tkair + (7.75e-5 * bw + 1.105e-6 * bw * bw) * (tkice - tkair);
  }
  /* Sturm1997: all literals are default-kind REAL, integer 156 and 1000 */
  if (bw <= 156.0) return R4(0.023) + R4(0.234) * (bw / 1000.0);
  return R4(0.138) - R4(1.01) * (bw / 1000.0) + (R4(3.233) * ((bw / 1000.0) * (bw / 1000.0)));
}

/* SoilThermProp :602-901 */
static void SoilThermProp(st_ctx* x) {
  const int numc = x->numc; const int32_t* fl = x->filterc;
  for (int j = -NLEVSNO + 1; j <= NLEVGRND; ++j) {          /* :695-783 */
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      const int lt = C1(lun_itype, c);
      const int snl = C1(snl, c);
      if (j >= 1) {
        if (lt != CTSM_ISTWET && lt != CTSM_ISTICE) {
          double satw = (LIQ(c, j) / denh2o + ICE(c, j) / denice) / (DZ(c, j) * C2(watsat, c, j, 1));
          satw = fmin(1.0, satw);
          if (satw > .1e-6) {
            double dke;
            if (T_SOISNO(c, j) >= tfrz) dke = fmax(0.0, log10(satw) + 1.0);
            else dke = satw;
            const double flq = (LIQ(c, j) / (denh2o * DZ(c, j))) /
                               (LIQ(c, j) / (denh2o * DZ(c, j)) + ICE(c, j) / (denice * DZ(c, j)));
            const double dksat = C2(tkmg, c, j, 1) * pow(tkwat, flq * C2(watsat, c, j, 1)) *
                                 pow(tkice, (1.0 - flq) * C2(watsat, c, j, 1));
            THK(c, j) = dke * dksat + (1.0 - dke) * C2(tkdry, c, j, 1);
          } else {
            THK(c, j) = C2(tkdry, c, j, 1);
          }
          if (j > C1(nbedrock, c)) THK(c, j) = thk_bedrock;
        } else if (lt == CTSM_ISTICE) {
          THK(c, j) = tkwat;
          if (T_SOISNO(c, j) < tfrz) THK(c, j) = tkice;
        } else if (lt == CTSM_ISTWET) {
          if (j > NLEVSOI) {
            THK(c, j) = thk_bedrock;
          } else {
            THK(c, j) = tkwat;
            if (T_SOISNO(c, j) < tfrz) THK(c, j) = tkice;
          }
        }
      }
      if (snl + 1 < 1 && j >= snl + 1 && j <= 0) {          /* :741 */
        C2(bw, c, j, SNOSOI_LO) = (ICE(c, j) + LIQ(c, j)) / (C1(frac_sno_eff, c) * DZ(c, j));
        const int method = (lt == CTSM_ISTICE) ? x->prm->snow_thermal_cond_glc_method
                                               : x->prm->snow_thermal_cond_method;
        THK(c, j) = snow_thk(method, C2(bw, c, j, SNOSOI_LO));
      }
    }
  }
  for (int j = -NLEVSNO + 1; j <= NLEVGRND; ++j) {          /* :803-826 */
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      if (j >= C1(snl, c) + 1 && j <= NLEVGRND - 1) {
        T2(tk, c, j, SNOSOI_LO) = THK(c, j) * THK(c, j + 1) * (Z(c, j + 1) - Z(c, j)) /
                                  (THK(c, j) * (Z(c, j + 1) - ZI(c, j)) + THK(c, j + 1) * (ZI(c, j) - Z(c, j)));
      } else if (j == NLEVGRND) {
        T2(tk, c, j, SNOSOI_LO) = 0.0;
      }
    }
  }
  for (int fc = 0; fc < numc; ++fc) {                       /* :829-835 */
    const int c = fl[fc];
    const double zh2osfc = R4(1.0e-3) * (0.5 * C1(h2osfc, c));
    T1(tk_h2osfc, c) = tkwat * THK(c, 1) * (Z(c, 1) + zh2osfc) / (tkwat * Z(c, 1) + THK(c, 1) * zh2osfc);
  }
  for (int j = 1; j <= NLEVGRND; ++j) {                     /* :839-858 */
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      const int lt = C1(lun_itype, c);
      if (lt != CTSM_ISTWET && lt != CTSM_ISTICE) {
        T2(cv, c, j, SNOSOI_LO) = C2(csol, c, j, 1) * (1.0 - C2(watsat, c, j, 1)) * DZ(c, j) +
                                  (ICE(c, j) * cpice + LIQ(c, j) * cpliq);
        if (j > C1(nbedrock, c)) T2(cv, c, j, SNOSOI_LO) = csol_bedrock * DZ(c, j);
      } else if (lt == CTSM_ISTWET) {
        T2(cv, c, j, SNOSOI_LO) = (ICE(c, j) * cpice + LIQ(c, j) * cpliq);
        if (j > C1(nbedrock, c)) T2(cv, c, j, SNOSOI_LO) = csol_bedrock * DZ(c, j);
      } else if (lt == CTSM_ISTICE) {
        T2(cv, c, j, SNOSOI_LO) = (ICE(c, j) * cpice + LIQ(c, j) * cpliq);
      }
    }
  }
  for (int fc = 0; fc < numc; ++fc) {                       /* :875-880 */
    const int c = fl[fc];
    if (C1(h2osno_no_layers, c) > 0.0) T2(cv, c, 1, SNOSOI_LO) = T2(cv, c, 1, SNOSOI_LO) + cpice * C1(h2osno_no_layers, c);
  }
  for (int j = -NLEVSNO + 1; j <= 0; ++j) {                 /* :884-895 */
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      if (C1(snl, c) + 1 < 1 && j >= C1(snl, c) + 1) {
        if (C1(frac_sno_eff, c) > 0.0)
          T2(cv, c, j, SNOSOI_LO) = fmax(thin_sfclayer, (cpliq * LIQ(c, j) + cpice * ICE(c, j)) / C1(frac_sno_eff, c));
        else
          T2(cv, c, j, SNOSOI_LO) = thin_sfclayer;
      }
    }
  }
}

static double pow4(double t) { const double t2 = t * t; return t2 * t2; }   /* t**4 as gfortran expands it */
static double pow3(double t) { return t * t * t; }


/* ComputeGroundHeatFluxAndDeriv :1543-1796 (non-urban branch) */
static void ComputeGroundHeatFluxAndDeriv(st_ctx* x) {
  const int numc = x->numc; const int32_t* fl = x->filterc;
  const size_t nc = x->ldt;
  double* lwrad_emit = malloc(sizeof(double) * nc), *dlwrad_emit = malloc(sizeof(double) * nc);
  double* lwrad_emit_snow = malloc(sizeof(double) * nc), *lwrad_emit_soil = malloc(sizeof(double) * nc);
  double* lwrad_emit_h2osfc = malloc(sizeof(double) * nc), *hs = malloc(sizeof(double) * nc);
#define L1(a, c) a[(c) - x->begc]
  for (int fc = 0; fc < numc; ++fc) {                       /* :1664-1673 */
    const int c = fl[fc];
    L1(lwrad_emit, c) = C1(emg, c) * sb * pow4(C1(t_grnd, c));
    L1(dlwrad_emit, c) = 4.0 * C1(emg, c) * sb * pow3(C1(t_grnd, c));
    L1(lwrad_emit_snow, c) = C1(emg, c) * sb * pow4(T_SOISNO(c, C1(snl, c) + 1));
    L1(lwrad_emit_soil, c) = C1(emg, c) * sb * pow4(T_SOISNO(c, 1));
    L1(lwrad_emit_h2osfc, c) = C1(emg, c) * sb * pow4(C1(t_h2osfc, c));
  }
  for (size_t i = 0; i < nc; ++i) { x->hs_soil[i] = 0; x->hs_h2osfc[i] = 0; hs[i] = 0; x->dhsdT[i] = 0; }
  for (int fp = 0; fp < x->nump; ++fp) {                    /* :1679-1745 */
    const int p = x->filterp[fp];
    const int c = P1(column, p);
    const double fv = (double)P1(frac_veg_nosno, p);
    P1(eflx_gnet, p) = P1(sabg, p) + P1(dlrad, p) + (1.0 - fv) * C1(emg, c) * C1(forc_lwrad, c) - L1(lwrad_emit, c)
                       - (P1(eflx_sh_grnd, p) + P1(qflx_evap_soi, p) * C1(htvp, c));
    P1(sabg_chk, p) = C1(frac_sno_eff, c) * P1(sabg_snow, p) + (1.0 - C1(frac_sno_eff, c)) * P1(sabg_soil, p);
    const double eflx_gnet_soil = P1(sabg_soil, p) + P1(dlrad, p) + (1.0 - fv) * C1(emg, c) * C1(forc_lwrad, c)
                                  - L1(lwrad_emit_soil, c) - (P1(eflx_sh_soil, p) + P1(qflx_ev_soil, p) * C1(htvp, c));
    const double eflx_gnet_h2osfc = P1(sabg_soil, p) + P1(dlrad, p) + (1.0 - fv) * C1(emg, c) * C1(forc_lwrad, c)
                                    - L1(lwrad_emit_h2osfc, c) - (P1(eflx_sh_h2osfc, p) + P1(qflx_ev_h2osfc, p) * C1(htvp, c));
    P1(dgnetdT, p) = -P1(cgrnd, p) - L1(dlwrad_emit, c);
    L1(hs, c) = L1(hs, c) + P1(eflx_gnet, p) * P1(wtcol, p);
    T1(dhsdT, c) = T1(dhsdT, c) + P1(dgnetdT, p) * P1(wtcol, p);
    T1(hs_soil, c) = T1(hs_soil, c) + eflx_gnet_soil * P1(wtcol, p);
    T1(hs_h2osfc, c) = T1(hs_h2osfc, c) + eflx_gnet_h2osfc * P1(wtcol, p);
  }
  for (size_t i = 0; i < nc * (NLEVSNO + 1); ++i) x->sabg_lyr_col[i] = 0.0;   /* :1757 */
  for (size_t i = 0; i < nc; ++i) { x->hs_top[i] = 0; x->hs_top_snow[i] = 0; }
  for (int fp = 0; fp < x->nump; ++fp) {                    /* :1760-1792 */
    const int p = x->filterp[fp];
    const int c = P1(column, p);
    const int lyr_top = C1(snl, c) + 1;
    const double fv = (double)P1(frac_veg_nosno, p);
    const double eflx_gnet_top = P2(sabg_lyr, p, lyr_top, SNOSOI_LO) + P1(dlrad, p) + (1.0 - fv) * C1(emg, c) * C1(forc_lwrad, c)
                                 - L1(lwrad_emit, c) - (P1(eflx_sh_grnd, p) + P1(qflx_evap_soi, p) * C1(htvp, c));
    T1(hs_top, c) = T1(hs_top, c) + eflx_gnet_top * P1(wtcol, p);
    const double eflx_gnet_snow = P2(sabg_lyr, p, lyr_top, SNOSOI_LO) + P1(dlrad, p) + (1.0 - fv) * C1(emg, c) * C1(forc_lwrad, c)
                                  - L1(lwrad_emit_snow, c) - (P1(eflx_sh_snow, p) + P1(qflx_ev_snow, p) * C1(htvp, c));
    T1(hs_top_snow, c) = T1(hs_top_snow, c) + eflx_gnet_snow * P1(wtcol, p);
    for (int j = lyr_top; j <= 1; ++j)
      T2(sabg_lyr_col, c, j, SNOSOI_LO) = T2(sabg_lyr_col, c, j, SNOSOI_LO) + P2(sabg_lyr, p, j, SNOSOI_LO) * P1(wtcol, p);
  }
  free(lwrad_emit); free(dlwrad_emit); free(lwrad_emit_snow); free(lwrad_emit_soil); free(lwrad_emit_h2osfc); free(hs);
#undef L1
}

/* ComputeHeatDiffFluxAndFactor :1799-1910 (non-urban branch) */
static void ComputeHeatDiffFluxAndFactor(st_ctx* x) {
  const double dtime = x->prm->dtime;
  for (int j = -NLEVSNO + 1; j <= NLEVGRND; ++j) {
    for (int fc = 0; fc < x->numc; ++fc) {
      const int c = x->filterc[fc];
      const int snl = C1(snl, c);
      if (j >= snl + 1) {
        if (j == snl + 1) {
          FACT(c, j) = dtime / T2(cv, c, j, SNOSOI_LO) * DZ(c, j) /
                       (0.5 * (Z(c, j) - ZI(c, j - 1) + capr * (Z(c, j + 1) - ZI(c, j - 1))));
          T2(fn, c, j, SNOSOI_LO) = T2(tk, c, j, SNOSOI_LO) * (T_SOISNO(c, j + 1) - T_SOISNO(c, j)) / (Z(c, j + 1) - Z(c, j));
        } else if (j <= NLEVGRND - 1) {
          FACT(c, j) = dtime / T2(cv, c, j, SNOSOI_LO);
          T2(fn, c, j, SNOSOI_LO) = T2(tk, c, j, SNOSOI_LO) * (T_SOISNO(c, j + 1) - T_SOISNO(c, j)) / (Z(c, j + 1) - Z(c, j));
        } else if (j == NLEVGRND) {
          FACT(c, j) = dtime / T2(cv, c, j, SNOSOI_LO);
          T2(fn, c, j, SNOSOI_LO) = C1(eflx_bot, c);
        }
      }
    }
  }
}

/* SetRHSVec, SetRHSVec_Snow, _StandingSurfaceWater, _Soil :1913-2353 (non-urban).
 * rvector rows: snow layer j -> row j-1, row 0 = standing water, rows 1..nlevgrnd soil. */
#define RV(c, r) T2(rvector, c, r, SNOSOI0_LO)
static void SetRHSVec(st_ctx* x) {
  const double dtime = x->prm->dtime;
  const int numc = x->numc; const int32_t* fl = x->filterc;
  for (size_t i = 0; i < x->ldt * (NLEVSNO + NLEVGRND + 1); ++i) x->rvector[i] = NAN;   /* :2003 */
  /* SetRHSVec_Snow :2060-2150 */
  for (int j = -NLEVSNO + 1; j <= 0; ++j) {
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      const int snl = C1(snl, c);
      const double hs_top_lev = T1(hs_top_snow, c);
      if (j == snl + 1) {
        RV(c, j - 1) = T_SOISNO(c, j) + FACT(c, j) * (hs_top_lev - T1(dhsdT, c) * T_SOISNO(c, j) + cnfac * T2(fn, c, j, SNOSOI_LO));
      } else if (j > snl + 1) {
        RV(c, j - 1) = T_SOISNO(c, j) + cnfac * FACT(c, j) * (T2(fn, c, j, SNOSOI_LO) - T2(fn, c, j - 1, SNOSOI_LO));
        RV(c, j - 1) = RV(c, j - 1) + FACT(c, j) * T2(sabg_lyr_col, c, j, SNOSOI_LO);
      }
    }
  }
  /* SetRHSVec_StandingSurfaceWater :2153-2207 */
  for (int fc = 0; fc < numc; ++fc) {
    const int c = fl[fc];
    const double dzm = (0.5 * T1(dz_h2osfc, c) + Z(c, 1));
    T1(fn_h2osfc, c) = T1(tk_h2osfc, c) * (T_SOISNO(c, 1) - C1(t_h2osfc, c)) / dzm;
    RV(c, 0) = C1(t_h2osfc, c) + (dtime / C1(c_h2osfc, c)) *
               (T1(hs_h2osfc, c) - T1(dhsdT, c) * C1(t_h2osfc, c) + cnfac * T1(fn_h2osfc, c));
  }
  /* SetRHSVec_Soil :2210-2353 */
  for (int j = 1; j <= NLEVGRND; ++j) {
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      const int snl = C1(snl, c);
      if (j == snl + 1) {
        RV(c, j) = T_SOISNO(c, j) + FACT(c, j) * (T1(hs_top_snow, c) - T1(dhsdT, c) * T_SOISNO(c, j) + cnfac * T2(fn, c, j, SNOSOI_LO));
      } else if (j == 1) {
        RV(c, j) = T_SOISNO(c, j) + FACT(c, j) *
                   ((1.0 - C1(frac_sno_eff, c)) * (T1(hs_soil, c) - T1(dhsdT, c) * T_SOISNO(c, j)) +
                    cnfac * (T2(fn, c, j, SNOSOI_LO) - C1(frac_sno_eff, c) * T2(fn, c, j - 1, SNOSOI_LO)));
        RV(c, j) = RV(c, j) + C1(frac_sno_eff, c) * FACT(c, j) * T2(sabg_lyr_col, c, j, SNOSOI_LO);
      } else if (j <= NLEVGRND - 1) {
        RV(c, j) = T_SOISNO(c, j) + cnfac * FACT(c, j) * (T2(fn, c, j, SNOSOI_LO) - T2(fn, c, j - 1, SNOSOI_LO));
      } else if (j == NLEVGRND) {
        RV(c, j) = T_SOISNO(c, j) - cnfac * FACT(c, j) * T2(fn, c, j - 1, SNOSOI_LO) + FACT(c, j) * T2(fn, c, j, SNOSOI_LO);
      }
    }
  }
  for (int fc = 0; fc < numc; ++fc) {                       /* :2342-2349 */
    const int c = fl[fc];
    if (C1(frac_h2osfc, c) != 0.0) {
      RV(c, 1) = RV(c, 1) - C1(frac_h2osfc, c) * FACT(c, 1) *
                 ((T1(hs_soil, c) - T1(dhsdT, c) * T_SOISNO(c, 1)) + cnfac * T1(fn_h2osfc, c));
    }
  }
}

/* SetMatrix, SetMatrix_Snow, _Soil, _StandingSurfaceWater, AssembleMatrixFromSubmatrices
 * :2356-2926 (non-urban).  The seven sub-matrices of the Fortran are zero-filled
 * and then copied entry by entry into bmatrix (:2492-2548); writing each entry
 * straight into a zero-filled bmatrix is the same assignment. */
static void SetMatrix(st_ctx* x) {
  const double dtime = x->prm->dtime;
  const int numc = x->numc; const int32_t* fl = x->filterc;
  for (size_t i = 0; i < x->ldt * 5 * (NLEVSNO + NLEVGRND + 1); ++i) x->bmatrix[i] = 0.0;   /* :2520 */
  /* SetMatrix_Snow :2551-2639 */
  for (int j = -NLEVSNO + 1; j <= 0; ++j) {
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      const int snl = C1(snl, c);
      if (j >= snl + 1) {
        const double dzp = Z(c, j + 1) - Z(c, j);
        if (j == snl + 1) {
          BM(c, 4, j - 1) = 0.0;
          BM(c, 3, j - 1) = 1.0 + (1.0 - cnfac) * FACT(c, j) * T2(tk, c, j, SNOSOI_LO) / dzp - FACT(c, j) * T1(dhsdT, c);
        } else {
          const double dzm = (Z(c, j) - Z(c, j - 1));
          BM(c, 4, j - 1) = -(1.0 - cnfac) * FACT(c, j) * T2(tk, c, j - 1, SNOSOI_LO) / dzm;
          BM(c, 3, j - 1) = 1.0 + (1.0 - cnfac) * FACT(c, j) * (T2(tk, c, j, SNOSOI_LO) / dzp + T2(tk, c, j - 1, SNOSOI_LO) / dzm);
        }
        if (j != 0) BM(c, 2, j - 1) = -(1.0 - cnfac) * FACT(c, j) * T2(tk, c, j, SNOSOI_LO) / dzp;
        else        BM(c, 1, j - 1) = -(1.0 - cnfac) * FACT(c, j) * T2(tk, c, j, SNOSOI_LO) / dzp;   /* snow_soil */
      }
    }
  }
  /* SetMatrix_Soil :2642-2810 */
  for (int j = 1; j <= NLEVGRND; ++j) {
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      const int snl = C1(snl, c);
      if (j == snl + 1) {
        const double dzp = Z(c, j + 1) - Z(c, j);
        /* j == 1 here: bmatrix_soil_snow(c,5,1) = 0 */
        BM(c, 5, j) = 0.0;
        BM(c, 3, j) = 1.0 + (1.0 - cnfac) * FACT(c, j) * T2(tk, c, j, SNOSOI_LO) / dzp - FACT(c, j) * T1(dhsdT, c);
        BM(c, 2, j) = -(1.0 - cnfac) * FACT(c, j) * T2(tk, c, j, SNOSOI_LO) / dzp;
      } else if (j == 1) {
        const double dzm = (Z(c, j) - Z(c, j - 1));
        const double dzp = (Z(c, j + 1) - Z(c, j));
        BM(c, 2, j) = -(1.0 - cnfac) * FACT(c, j) * T2(tk, c, j, SNOSOI_LO) / dzp;
        BM(c, 3, j) = 1.0 + (1.0 - cnfac) * FACT(c, j) * (T2(tk, c, j, SNOSOI_LO) / dzp + C1(frac_sno_eff, c) * T2(tk, c, j - 1, SNOSOI_LO) / dzm)
                      - (1.0 - C1(frac_sno_eff, c)) * FACT(c, j) * T1(dhsdT, c);
        BM(c, 5, j) = -C1(frac_sno_eff, c) * (1.0 - cnfac) * FACT(c, j) * T2(tk, c, j - 1, SNOSOI_LO) / dzm;
      } else if (j <= NLEVGRND - 1) {
        const double dzm = (Z(c, j) - Z(c, j - 1));
        const double dzp = (Z(c, j + 1) - Z(c, j));
        BM(c, 2, j) = -(1.0 - cnfac) * FACT(c, j) * T2(tk, c, j, SNOSOI_LO) / dzp;
        BM(c, 3, j) = 1.0 + (1.0 - cnfac) * FACT(c, j) * (T2(tk, c, j, SNOSOI_LO) / dzp + T2(tk, c, j - 1, SNOSOI_LO) / dzm);
        BM(c, 4, j) = -(1.0 - cnfac) * FACT(c, j) * T2(tk, c, j - 1, SNOSOI_LO) / dzm;
      } else if (j == NLEVGRND) {
        const double dzm = (Z(c, j) - Z(c, j - 1));
        BM(c, 2, j) = 0.0;
        BM(c, 3, j) = 1.0 + (1.0 - cnfac) * FACT(c, j) * T2(tk, c, j - 1, SNOSOI_LO) / dzm;
        BM(c, 4, j) = -(1.0 - cnfac) * FACT(c, j) * T2(tk, c, j - 1, SNOSOI_LO) / dzm;
      }
    }
  }
  for (int fc = 0; fc < numc; ++fc) {                       /* :2797-2807 */
    const int c = fl[fc];
    const double dzm = (0.5 * T1(dz_h2osfc, c) + Z(c, 1));
    if (C1(frac_h2osfc, c) != 0.0) {
      BM(c, 3, 1) = BM(c, 3, 1) + C1(frac_h2osfc, c) *
                    ((1.0 - cnfac) * FACT(c, 1) * T1(tk_h2osfc, c) / dzm + FACT(c, 1) * T1(dhsdT, c));
    }
  }
  /* SetMatrix_StandingSurfaceWater :2813-2926 */
  for (int fc = 0; fc < numc; ++fc) {
    const int c = fl[fc];
    const double dzm = (0.5 * T1(dz_h2osfc, c) + Z(c, 1));
    BM(c, 3, 0) = 1.0 + (1.0 - cnfac) * (dtime / C1(c_h2osfc, c)) * T1(tk_h2osfc, c) / dzm - (dtime / C1(c_h2osfc, c)) * T1(dhsdT, c);
    BM(c, 2, 0) = -(1.0 - cnfac) * (dtime / C1(c_h2osfc, c)) * T1(tk_h2osfc, c) / dzm;
    if (C1(frac_h2osfc, c) != 0.0)
      BM(c, 4, 1) = -C1(frac_h2osfc, c) * (1.0 - cnfac) * FACT(c, 1) * T1(tk_h2osfc, c) / dzm;
  }
}

/* PhaseChangeH2osfc :904-1130 */
static void PhaseChangeH2osfc(st_ctx* x) {
  const double dtime = x->prm->dtime;
  const int numc = x->numc; const int32_t* fl = x->filterc;
  for (int fc = 0; fc < numc; ++fc) {
    const int c = fl[fc];
    C1(xmf_h2osfc, c) = 0.0;
    C1(qflx_h2osfc_to_ice, c) = 0.0;
    C1(eflx_h2osfc_to_snow, c) = 0.0;
  }
  for (int fc = 0; fc < numc; ++fc) {
    const int c = fl[fc];
    const int snl = C1(snl, c);
    const double frac_sno = C1(frac_sno_eff, c);           /* :953 frac_sno => frac_sno_eff_col */
    const double frac_h2osfc = C1(frac_h2osfc, c);
    const double dhsdT = T1(dhsdT, c);
    /* CalculateTotalH2osno, WaterStateType.F90:887-897 */
    double h2osno_total = C1(h2osno_no_layers, c);
    for (int j = snl + 1; j <= 0; ++j) h2osno_total = h2osno_total + ICE(c, j) + LIQ(c, j);

    if (frac_h2osfc > 0.0 && C1(t_h2osfc, c) <= tfrz) {    /* :993 */
      const double tinc = tfrz - C1(t_h2osfc, c);
      C1(t_h2osfc, c) = tfrz;
      const double hm = frac_h2osfc * (dhsdT * tinc - tinc * C1(c_h2osfc, c) / dtime);
      const double xm = hm * dtime / hfus;
      const double temp1 = C1(h2osfc, c) + xm;
      const double z_avg = frac_sno * C1(snow_depth, c);
      double rho_avg;
      if (z_avg > 0.0) rho_avg = fmin(800.0, h2osno_total / z_avg);
      else rho_avg = 200.0;
      if (temp1 >= 0.0) {                                   /* :1013 */
        C1(int_snow, c) = C1(int_snow, c) - xm;
        if (snl == 0) C1(h2osno_no_layers, c) = C1(h2osno_no_layers, c) - xm;
        else ICE(c, 0) = ICE(c, 0) - xm;
        h2osno_total = h2osno_total - xm;
        C1(h2osfc, c) = C1(h2osfc, c) + xm;
        C1(xmf_h2osfc, c) = hm;
        C1(qflx_h2osfc_to_ice, c) = -xm / dtime;
        if (frac_sno > 0 && snl < 0) C1(snow_depth, c) = h2osno_total / (rho_avg * frac_sno);
        else C1(snow_depth, c) = h2osno_total / denice;
        if (snl == 0) {
          T_SOISNO(c, 0) = C1(t_h2osfc, c);
          C1(eflx_h2osfc_to_snow, c) = 0.;
        } else {
          double c1, c2;
          if (snl == -1) c1 = frac_sno * (dtime / FACT(c, 0) - dhsdT * dtime);
          else c1 = frac_sno / FACT(c, 0) * dtime;
          if (frac_h2osfc != 0.0) c2 = (-cpliq * xm - frac_h2osfc * dhsdT * dtime);
          else c2 = 0.0;
          T_SOISNO(c, 0) = (c1 * T_SOISNO(c, 0) + c2 * C1(t_h2osfc, c)) / (c1 + c2);
          C1(eflx_h2osfc_to_snow, c) = (C1(t_h2osfc, c) - T_SOISNO(c, 0)) * c2 / dtime;
        }
      } else {                                              /* :1062 all h2osfc converted to ice */
        rho_avg = (h2osno_total * rho_avg + C1(h2osfc, c) * denice) / (h2osno_total + C1(h2osfc, c));
        C1(int_snow, c) = C1(int_snow, c) + C1(h2osfc, c);
        if (snl == 0) C1(h2osno_no_layers, c) = C1(h2osno_no_layers, c) + C1(h2osfc, c);
        else ICE(c, 0) = ICE(c, 0) + C1(h2osfc, c);
        h2osno_total = h2osno_total + C1(h2osfc, c);
        C1(qflx_h2osfc_to_ice, c) = C1(h2osfc, c) / dtime;
        C1(t_h2osfc, c) = C1(t_h2osfc, c) - temp1 * hfus / (dtime * dhsdT - C1(c_h2osfc, c));
        C1(xmf_h2osfc, c) = (hm - frac_h2osfc * temp1 * hfus / dtime);
        if (snl == 0) {
          T_SOISNO(c, 0) = C1(t_h2osfc, c);
        } else {
          double c1, c2;
          if (snl == -1) c1 = frac_sno * (dtime / FACT(c, 0) - dhsdT * dtime);
          else c1 = frac_sno / FACT(c, 0) * dtime;
          if (frac_h2osfc != 0.0) c2 = frac_h2osfc * (C1(c_h2osfc, c) - dtime * dhsdT);
          else c2 = 0.0;
          T_SOISNO(c, 0) = (c1 * T_SOISNO(c, 0) + c2 * C1(t_h2osfc, c)) / (c1 + c2);
          C1(t_h2osfc, c) = T_SOISNO(c, 0);
        }
        C1(h2osfc, c) = 0.0;
        if (frac_sno > 0 && snl < 0) C1(snow_depth, c) = h2osno_total / (rho_avg * frac_sno);
        else C1(snow_depth, c) = h2osno_total / denice;
      }
    }
  }
}

/* Phasechange :1133-1540 (non-urban, excess_ice == 0) */
static void Phasechange(st_ctx* x) {
  const double dtime = x->prm->dtime;
  const int numc = x->numc; const int32_t* fl = x->filterc;
  const size_t nc = x->ldt;
  const int NL = NLEVSNO + NLEVGRND;
  double* hm = calloc(nc * NL, sizeof(double)), *xm = calloc(nc * NL, sizeof(double));
  double* wmass0 = calloc(nc * NL, sizeof(double)), *wice0 = calloc(nc * NL, sizeof(double));
  double* wliq0 = calloc(nc * NL, sizeof(double)), *tinc = calloc(nc * NL, sizeof(double));
  double* supercool = calloc(nc * NLEVGRND, sizeof(double));
#define W2(a, c, j) a[(size_t)((j) - SNOSOI_LO) * nc + ((c) - x->begc)]
#define SC(c, j) supercool[(size_t)((j) - 1) * nc + ((c) - x->begc)]
#define IMELT(c, j) C2(imelt, c, j, SNOSOI_LO)
  for (int fc = 0; fc < numc; ++fc) {                       /* :1236-1245 */
    const int c = fl[fc];
    C1(xmf, c) = 0.0; C1(qflx_snomelt, c) = 0.0; C1(qflx_snofrz, c) = 0.0; C1(qflx_snow_drain, c) = 0.0;
  }
  for (int j = -NLEVSNO + 1; j <= NLEVGRND; ++j) {          /* :1247-1273 */
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      if (j >= C1(snl, c) + 1) {
        IMELT(c, j) = 0;
        W2(hm, c, j) = 0.0; W2(xm, c, j) = 0.0;
        W2(wice0, c, j) = ICE(c, j);
        W2(wliq0, c, j) = LIQ(c, j);
        W2(wmass0, c, j) = ICE(c, j) + LIQ(c, j);           /* + wexice0 == 0 */
      }
      if (j <= 0) {
        C2(qflx_snomelt_lyr, c, j, SNOSOI_LO) = 0.0;
        C2(qflx_snofrz_lyr, c, j, SNOSOI_LO) = 0.0;
      }
    }
  }
  for (int j = -NLEVSNO + 1; j <= 0; ++j) {                 /* :1276-1299 snow layers */
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      if (j >= C1(snl, c) + 1) {
        if (ICE(c, j) > 0.0 && T_SOISNO(c, j) > tfrz) {
          IMELT(c, j) = 1; W2(tinc, c, j) = tfrz - T_SOISNO(c, j); T_SOISNO(c, j) = tfrz;
        }
        if (LIQ(c, j) > 0.0 && T_SOISNO(c, j) < tfrz) {
          IMELT(c, j) = 2; W2(tinc, c, j) = tfrz - T_SOISNO(c, j); T_SOISNO(c, j) = tfrz;
        }
      }
    }
  }
  for (int j = 1; j <= NLEVGRND; ++j) {                     /* :1302-1357 soil layers */
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      const int lt = C1(lun_itype, c);
      SC(c, j) = 0.0;
      if (ICE(c, j) > 0. && T_SOISNO(c, j) > tfrz) {
        IMELT(c, j) = 1; W2(tinc, c, j) = tfrz - T_SOISNO(c, j); T_SOISNO(c, j) = tfrz;
      }
      SC(c, j) = 0.0;
      if (is_soil_or_crop(lt)) {
        if (T_SOISNO(c, j) < tfrz) {
          const double smp = hfus * (tfrz - T_SOISNO(c, j)) / (grav * T_SOISNO(c, j)) * 1000.0;
          SC(c, j) = C2(watsat, c, j, 1) * pow(smp / C2(sucsat, c, j, 1), -1.0 / C2(bsw, c, j, 1));
          SC(c, j) = SC(c, j) * DZ(c, j) * 1000.0;
        }
      }
      if (LIQ(c, j) > SC(c, j) && T_SOISNO(c, j) < tfrz) {
        IMELT(c, j) = 2; W2(tinc, c, j) = tfrz - T_SOISNO(c, j); T_SOISNO(c, j) = tfrz;
      }
      if (C1(h2osno_no_layers, c) > 0.0 && j == 1) {
        if (T_SOISNO(c, j) > tfrz) {
          IMELT(c, j) = 1; W2(tinc, c, j) = tfrz - T_SOISNO(c, j); T_SOISNO(c, j) = tfrz;
        }
      }
    }
  }
  for (int j = -NLEVSNO + 1; j <= NLEVGRND; ++j) {          /* :1360-1518 */
    for (int fc = 0; fc < numc; ++fc) {
      const int c = fl[fc];
      const int snl = C1(snl, c);
      const double frac_sno_eff = C1(frac_sno_eff, c), frac_h2osfc = C1(frac_h2osfc, c);
      const double dhsdT = T1(dhsdT, c);
      if (j >= snl + 1) {
        if (IMELT(c, j) > 0) {
          const double ti = W2(tinc, c, j);
          if (j == snl + 1) {
            if (j > 0) W2(hm, c, j) = dhsdT * ti - ti / FACT(c, j);
            else W2(hm, c, j) = frac_sno_eff * (dhsdT * ti - ti / FACT(c, j));
            if (j == 1 && frac_h2osfc != 0.0) W2(hm, c, j) = W2(hm, c, j) - frac_h2osfc * (dhsdT * ti);
          } else if (j == 1) {
            W2(hm, c, j) = (1.0 - frac_sno_eff - frac_h2osfc) * dhsdT * ti - ti / FACT(c, j);
          } else {
            if (j < 1) W2(hm, c, j) = -frac_sno_eff * (ti / FACT(c, j));
            else W2(hm, c, j) = -ti / FACT(c, j);
          }
        }
        if (IMELT(c, j) == 1 && W2(hm, c, j) < 0.0) { W2(hm, c, j) = 0.0; IMELT(c, j) = 0; }
        if (IMELT(c, j) == 2 && W2(hm, c, j) > 0.0) { W2(hm, c, j) = 0.0; IMELT(c, j) = 0; }

        if (IMELT(c, j) > 0 && fabs(W2(hm, c, j)) > 0.0) {
          W2(xm, c, j) = W2(hm, c, j) * dtime / hfus;
          if (j == 1) {
            if (C1(h2osno_no_layers, c) > 0.0 && W2(xm, c, j) > 0.0) {
              const double temp1 = C1(h2osno_no_layers, c);
              C1(h2osno_no_layers, c) = fmax(0.0, temp1 - W2(xm, c, j));
              const double propor = C1(h2osno_no_layers, c) / temp1;
              C1(snow_depth, c) = propor * C1(snow_depth, c);
              const double heatr0 = W2(hm, c, j) - hfus * (temp1 - C1(h2osno_no_layers, c)) / dtime;
              if (heatr0 > 0.0) { W2(xm, c, j) = heatr0 * dtime / hfus; W2(hm, c, j) = heatr0; }
              else { W2(xm, c, j) = 0.0; W2(hm, c, j) = 0.0; }
              C1(qflx_snomelt, c) = fmax(0.0, (temp1 - C1(h2osno_no_layers, c))) / dtime;
              C1(xmf, c) = hfus * C1(qflx_snomelt, c);
              C1(qflx_snow_drain, c) = C1(qflx_snomelt, c);
            }
          }
          double heatr = 0.0;
          if (W2(xm, c, j) > 0.0) {
            ICE(c, j) = fmax(0.0, W2(wice0, c, j) - W2(xm, c, j));
            /* excess_ice == 0: wexice0 - excess_ice + wice0 - h2osoi_ice == wice0 - h2osoi_ice bit for bit */
            heatr = W2(hm, c, j) - hfus * (W2(wice0, c, j) - ICE(c, j)) / dtime;
          } else if (W2(xm, c, j) < 0.0) {
            if (j <= 0) {
              ICE(c, j) = fmin(W2(wmass0, c, j), W2(wice0, c, j) - W2(xm, c, j));
            } else {
              if (W2(wmass0, c, j) < SC(c, j)) ICE(c, j) = 0.0;
              else ICE(c, j) = fmin(W2(wmass0, c, j) - SC(c, j), W2(wice0, c, j) - W2(xm, c, j));
            }
            heatr = W2(hm, c, j) - hfus * (W2(wice0, c, j) - ICE(c, j)) / dtime;
          }
          LIQ(c, j) = fmax(0.0, W2(wmass0, c, j) - ICE(c, j));

          if (fabs(heatr) > 0.0) {
            if (j == snl + 1) {
              if (j == 1) T_SOISNO(c, j) = T_SOISNO(c, j) + FACT(c, j) * heatr / (1.0 - (1.0 - frac_h2osfc) * FACT(c, j) * dhsdT);
              else T_SOISNO(c, j) = T_SOISNO(c, j) + (FACT(c, j) / frac_sno_eff) * heatr / (1.0 - FACT(c, j) * dhsdT);
            } else if (j == 1) {
              T_SOISNO(c, j) = T_SOISNO(c, j) + FACT(c, j) * heatr / (1.0 - (1.0 - frac_sno_eff - frac_h2osfc) * FACT(c, j) * dhsdT);
            } else {
              if (j > 0) T_SOISNO(c, j) = T_SOISNO(c, j) + FACT(c, j) * heatr;
              else if (frac_sno_eff > 0.0) T_SOISNO(c, j) = T_SOISNO(c, j) + (FACT(c, j) / frac_sno_eff) * heatr;
            }
            if (j <= 0) {
              if (LIQ(c, j) * ICE(c, j) > 0.0) T_SOISNO(c, j) = tfrz;
            }
          }
          if (j >= 1) {
            /* + hfus*(wexice0-excess_ice)/dtime == + 0.0 */
            C1(xmf, c) = C1(xmf, c) + hfus * (W2(wice0, c, j) - ICE(c, j)) / dtime + 0.0;
          } else {
            C1(xmf, c) = C1(xmf, c) + hfus * (W2(wice0, c, j) - ICE(c, j)) / dtime;
          }
          if (IMELT(c, j) == 1 && j < 1) {
            C2(qflx_snomelt_lyr, c, j, SNOSOI_LO) = fmax(0.0, (W2(wice0, c, j) - ICE(c, j))) / dtime;
            C1(qflx_snomelt, c) = C1(qflx_snomelt, c) + C2(qflx_snomelt_lyr, c, j, SNOSOI_LO);
            C1(snomelt_accum, c) = C1(snomelt_accum, c) + C2(qflx_snomelt_lyr, c, j, SNOSOI_LO) * dtime * 1.e-3;
          }
          if (IMELT(c, j) == 2 && j < 1) {
            C2(qflx_snofrz_lyr, c, j, SNOSOI_LO) = fmax(0.0, (ICE(c, j) - W2(wice0, c, j))) / dtime;
            C1(qflx_snofrz, c) = C1(qflx_snofrz, c) + C2(qflx_snofrz_lyr, c, j, SNOSOI_LO);
          }
        }
      }
    }
  }
  for (int fc = 0; fc < numc; ++fc) {                       /* :1523-1534 */
    const int c = fl[fc];
    C1(eflx_snomelt, c) = C1(qflx_snomelt, c) * hfus;
    if (is_soil_or_crop(C1(lun_itype, c))) C1(eflx_snomelt_r, c) = C1(eflx_snomelt, c);
  }
  free(hm); free(xm); free(wmass0); free(wice0); free(wliq0); free(tinc); free(supercool);
#undef W2
#undef SC
#undef IMELT
}

/* SoilTemperature :92-599 */
int oracle_soiltemperature(const ctsm_params_t* prm, const ctsm_bounds_t* bounds, int num_nolakep,
                           const int32_t* filter_nolakep, int num_nolakec, const int32_t* filter_nolakec,
                           const ctsm_soiltemperature_fields_t* f, ctsm_status_t* st) {
  st_ctx xs; st_ctx* x = &xs;
  memset(x, 0, sizeof *x);
  if (st) memset(st, 0, sizeof *st);
  x->f = f; x->prm = prm;
  x->begc0 = f->alloc.begc; x->begp0 = f->alloc.begp;
  x->ldc = (size_t)(f->alloc.endc - f->alloc.begc + 1);
  x->ldp = (size_t)(f->alloc.endp - f->alloc.begp + 1);
  x->begc = bounds->begc; x->endc = bounds->endc;
  x->ldt = (size_t)(bounds->endc - bounds->begc + 1);
  x->numc = num_nolakec; x->filterc = filter_nolakec;
  x->nump = num_nolakep; x->filterp = filter_nolakep;
  const size_t nc = x->ldt;
  const int NL = NLEVSNO + NLEVGRND;
  const double dtime = prm->dtime;
  int rc = 0;

  for (int fc = 0; fc < num_nolakec; ++fc) {   /* urban columns are outside the hot path */
    const int lt = C1(lun_itype, filter_nolakec[fc]);
    if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) {
      if (st) { st->code = CTSM_ERR_URBAN; st->subgrid_level = CTSM_SUBGRID_COLUMN; st->subgrid_index = filter_nolakec[fc];
                snprintf(st->msg, sizeof st->msg, "SoilTemperature: urban column in filter_nolakec is outside the ctsm_b200 hot path"); }
      return CTSM_ERR_URBAN;
    }
  }

  x->cv = malloc(sizeof(double) * nc * NL); x->tk = malloc(sizeof(double) * nc * NL);
  x->fn = malloc(sizeof(double) * nc * NL); x->fn1 = malloc(sizeof(double) * nc * NL);
  x->sabg_lyr_col = malloc(sizeof(double) * nc * (NLEVSNO + 1));
  x->hs_top = malloc(sizeof(double) * nc); x->hs_soil = malloc(sizeof(double) * nc);
  x->hs_top_snow = malloc(sizeof(double) * nc); x->hs_h2osfc = malloc(sizeof(double) * nc);
  x->dhsdT = malloc(sizeof(double) * nc); x->fn_h2osfc = malloc(sizeof(double) * nc);
  x->dz_h2osfc = malloc(sizeof(double) * nc); x->tk_h2osfc = malloc(sizeof(double) * nc);
  x->bmatrix = malloc(sizeof(double) * nc * 5 * (NL + 1));
  x->tvector = malloc(sizeof(double) * nc * (NL + 1)); x->rvector = malloc(sizeof(double) * nc * (NL + 1));
  x->jtop = malloc(sizeof(int32_t) * nc); x->jbot = malloc(sizeof(int32_t) * nc);

  for (size_t i = 0; i < nc; ++i) x->jtop[i] = -9999;       /* :272 */
  for (int fc = 0; fc < num_nolakec; ++fc) {                /* :273-283 */
    const int c = filter_nolakec[fc];
    T1(jtop, c) = C1(snl, c);
    T1(jbot, c) = NLEVGRND;
  }
  for (size_t i = 0; i < nc; ++i) x->tk_h2osfc[i] = NAN;    /* :312 */
  SoilThermProp(x);                                         /* :313 */
  ComputeGroundHeatFluxAndDeriv(x);                         /* :323 */
  ComputeHeatDiffFluxAndFactor(x);                          /* :338 */
  for (int fc = 0; fc < num_nolakec; ++fc) {                /* :348-357 */
    const int c = filter_nolakec[fc];
    if ((C1(h2osfc, c) > thin_sfclayer) && (C1(frac_h2osfc, c) > thin_sfclayer)) {
      C1(c_h2osfc, c) = fmax(thin_sfclayer, cpliq * C1(h2osfc, c) / C1(frac_h2osfc, c));
      T1(dz_h2osfc, c) = fmax(thin_sfclayer, R4(1.0e-3) * C1(h2osfc, c) / C1(frac_h2osfc, c));
    } else {
      C1(c_h2osfc, c) = thin_sfclayer;
      T1(dz_h2osfc, c) = thin_sfclayer;
    }
  }
  SetRHSVec(x);                                             /* :362 */
  SetMatrix(x);                                             /* :382 */
#define TV(c, r) T2(tvector, c, r, SNOSOI0_LO)
  for (size_t i = 0; i < nc * (NL + 1); ++i) x->tvector[i] = NAN;   /* :396 */
  for (int fc = 0; fc < num_nolakec; ++fc) {                /* :397-409 */
    const int c = filter_nolakec[fc];
    for (int j = C1(snl, c) + 1; j <= 0; ++j) TV(c, j - 1) = T_SOISNO(c, j);
    TV(c, 0) = C1(t_h2osfc, c);
    for (int j = 1; j <= NLEVGRND; ++j) TV(c, j) = T_SOISNO(c, j);
  }
  {                                                         /* :415-417 BandDiagonal */
    ctsm_bounds_t tb = *bounds;
    rc = oracle_banddiagonal(&tb, -NLEVSNO, NLEVGRND, x->jtop, x->jbot, num_nolakec, filter_nolakec, 5,
                             x->bmatrix, x->rvector, x->tvector, st);
  }
  if (rc == 0) {
    for (int fc = 0; fc < num_nolakec; ++fc) {              /* :422-434 */
      const int c = filter_nolakec[fc];
      for (int j = C1(snl, c) + 1; j <= 0; ++j) T_SOISNO(c, j) = TV(c, j - 1);
      for (int j = 1; j <= NLEVGRND; ++j) T_SOISNO(c, j) = TV(c, j);
      if (C1(frac_h2osfc, c) == 0.0) C1(t_h2osfc, c) = T_SOISNO(c, 1);
      else C1(t_h2osfc, c) = TV(c, 0);
    }
    for (int j = -NLEVSNO + 1; j <= NLEVGRND; ++j) {        /* :438-483 */
      for (int fc = 0; fc < num_nolakec; ++fc) {
        const int c = filter_nolakec[fc];
        if (j >= C1(snl, c) + 1) {
          if (j <= NLEVGRND - 1)
            T2(fn1, c, j, SNOSOI_LO) = T2(tk, c, j, SNOSOI_LO) * (T_SOISNO(c, j + 1) - T_SOISNO(c, j)) / (Z(c, j + 1) - Z(c, j));
          else if (j == NLEVGRND)
            T2(fn1, c, j, SNOSOI_LO) = 0.0;
        }
      }
    }
    for (int fc = 0; fc < num_nolakec; ++fc) C1(xmf_h2osfc, filter_nolakec[fc]) = 0.;   /* :511-514 */
    PhaseChangeH2osfc(x);                                   /* :516 */
    Phasechange(x);                                         /* :520 */
    for (int fc = 0; fc < num_nolakec; ++fc) {              /* :546-568 */
      const int c = filter_nolakec[fc];
      const int snl = C1(snl, c);
      if (snl < 0) {
        if (C1(frac_h2osfc, c) != 0.0)
          C1(t_grnd, c) = C1(frac_sno_eff, c) * T_SOISNO(c, snl + 1) + (1.0 - C1(frac_sno_eff, c) - C1(frac_h2osfc, c)) * T_SOISNO(c, 1)
                          + C1(frac_h2osfc, c) * C1(t_h2osfc, c);
        else
          C1(t_grnd, c) = C1(frac_sno_eff, c) * T_SOISNO(c, snl + 1) + (1.0 - C1(frac_sno_eff, c)) * T_SOISNO(c, 1);
      } else {
        if (C1(frac_h2osfc, c) != 0.0)
          C1(t_grnd, c) = (1.0 - C1(frac_h2osfc, c)) * T_SOISNO(c, 1) + C1(frac_h2osfc, c) * C1(t_h2osfc, c);
        else
          C1(t_grnd, c) = T_SOISNO(c, 1);
      }
    }
    for (int fc = 0; fc < num_nolakec; ++fc) C1(eflx_fgr12, filter_nolakec[fc]) = 0.0;   /* :572-576 */
    for (int j = -NLEVSNO + 1; j <= NLEVGRND; ++j) {        /* :580-595 */
      for (int fc = 0; fc < num_nolakec; ++fc) {
        const int c = filter_nolakec[fc];
        const int lt = C1(lun_itype, c);
        if (j == 1) C1(eflx_fgr12, c) = -cnfac * T2(fn, c, 1, SNOSOI_LO) - (1.0 - cnfac) * T2(fn1, c, 1, SNOSOI_LO);
        if (j > 0 && j < NLEVGRND && is_soil_or_crop(lt))
          C2(eflx_fgr, c, j, 1) = -cnfac * T2(fn, c, j, SNOSOI_LO) - (1.0 - cnfac) * T2(fn1, c, j, SNOSOI_LO);
        else if (j == NLEVGRND && is_soil_or_crop(lt))
          C2(eflx_fgr, c, j, 1) = 0.0;
      }
    }
  }
#undef TV
  free(x->cv); free(x->tk); free(x->fn); free(x->fn1); free(x->sabg_lyr_col);
  free(x->hs_top); free(x->hs_soil); free(x->hs_top_snow); free(x->hs_h2osfc); free(x->dhsdT);
  free(x->fn_h2osfc); free(x->dz_h2osfc); free(x->tk_h2osfc);
  free(x->bmatrix); free(x->tvector); free(x->rvector); free(x->jtop); free(x->jbot);
  (void)dtime;
  return rc;
}
