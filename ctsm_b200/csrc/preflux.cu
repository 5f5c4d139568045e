// preflux.cu — the three routines clm_drv runs immediately before CanopyFluxes, on B200:
//   BiogeophysPreFluxCalcs     BiogeophysPreFluxCalcsMod.F90:58-406 (clm_driver.F90:680)
//     SetZ0mDisp :120-219, SetRoughnessLengthsAndForcHeightsNonLake FrictionVelocityMod.F90:543-685,
//     CalcInitialTemperatureAndEnergyVars :223-405, calc_soilevap_resis SurfaceResistanceMod.F90:192-426
//   CalculateSurfaceHumidity   SurfaceHumidityMod.F90:41-239 (clm_driver.F90:702)
//   BareGroundFluxes           BareGroundFluxesMod.F90:63-579 (clm_driver.F90:711)
//
// B200 mapping.  All three are maps over a filter with no coupling between points, so each is one thread per filter entry on
// the Fortran arrays (subgrid index fastest: coalesced where the filter is dense) and HBM-bound:
//   preflux_patch_a_kernel   SetZ0mDisp (reads the PREVIOUS step's z0mg of the patch's column, so it runs first)
//   preflux_col_kernel       ground roughness lengths, t_ssbef copy (37 levels, level-outer per thread = coalesced across the
//                            warp), t_grnd / emg / htvp / thv, soil evaporative resistance
//   preflux_patch_b_kernel   vegetation roughness lengths, patch forcing heights (need the NEW column roughness), zeroed fluxes,
//                            emv, thm
//   surface_humidity_kernel  one thread per column
//   bareground_kernel        one thread per patch without exposed vegetation: MoninObukIni, the three FrictionVelocity passes
//                            of the stability iteration and the fluxes, all in registers (the reference keeps ten clump-sized
//                            work arrays)
//   bareground_colcopy_kernel  z0hg / z0qg of a column = those of its LAST filter patch, what the reference's loop leaves
// Non-urban landunits only (CTSM_ERR_URBAN otherwise); use_fates = use_lch4 = .false.
#include "surface_layer.cuh"
#include "common.cuh"
#include <vector>

struct PreFluxDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_PREFLUX
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_PREFLUX
#undef CTSM_F
};
struct SurfHumDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_SURFACEHUMIDITY
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SURFACEHUMIDITY
#undef CTSM_F
};
struct BareDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_BAREGROUNDFLUXES
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_BAREGROUNDFLUXES
#undef CTSM_F
};

namespace {

constexpr double hsub = 2.501e6 + 3.337e5;                                     // SHR_CONST_LATICE + SHR_CONST_LATVAP
constexpr double roverg = 6.02214e26 * 1.38065e-23 / 18.016 / 9.80616 * 1000.0;   // clm_varcon.F90:55
constexpr double beta_param = 7.2, b1_param = 1.4, b4_param = -0.31;           // clm_varcon.F90:162-165
constexpr double meier_param1 = 0.23, meier_param2 = 0.08, meier_param3 = 70.0;   // clm_varcon.F90:167-169
constexpr int SNO_LO = -CTSM_NLEVSNO + 1;

struct PreGeo { int begc0, begp0, begg0, ldc, ldp; };
struct PrePrm {
  int z0param_method, soil_resis_method, use_z0m_snowmelt, time_flags, human_fast;
  double zlnd, zsno, zglc, d_max, frac_sat_soil_dsl_init, a_coef, a_exp, wind_min, zetamaxstable;
};
__device__ __forceinline__ bool is_urban(int lt) { return lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX; }

// SetZ0mDisp :120-219
__global__ void __launch_bounds__(256)
preflux_patch_a_kernel(PreFluxDev f, PrePrm prm, PreGeo g, int numf, const int32_t* __restrict__ filterp) {
  const int fp = blockIdx.x * blockDim.x + threadIdx.x;
  if (fp >= numf) return;
  const int pp = filterp[fp] - g.begp0;
  const int cc = f.column[pp] - g.begc0;
  const int ivt = f.itype[pp];
  const double htop = f.htop[pp];
  if (prm.z0param_method == 1) {                                               // ZengWang2007
    f.z0m[pp] = f.pft_z0mr[ivt] * htop;
    f.displa[pp] = f.pft_displar[ivt] * htop;
    return;
  }
  if (prm.time_flags & CTSM_TIME_FIRST_STEPS) { f.z0m[pp] = 0.0; f.displa[pp] = 0.0; return; }
  if (ivt == 0) { f.z0m[pp] = 0.0; f.displa[pp] = 0.0; return; }               // noveg (also covers the crop reset at new year)
  const double lm = f.pft_z0v_LAImax[ivt];
  const double displa = htop * (1.0 - (1.0 - dexp(-pw(cd1_param * lm, 0.5))) / pw(cd1_param * lm, 0.5));
  f.displa[pp] = displa;
  const double U_ustar = 4.0 * pw(f.pft_z0v_Cs[ivt] + f.pft_z0v_Cr[ivt] * lm / 2.0, -0.5) / lm / f.pft_z0v_c[ivt];
  if (htop <= 1.e-10) {
    f.z0m[pp] = f.z0mg[cc];
  } else {
    const double cw = f.pft_z0v_cw[ivt];
    f.z0m[pp] = htop * (1.0 - displa / htop) * dexp(-0.4 * U_ustar + dlog(cw) - 1.0 + 1.0 / cw);
  }
}

// SetRoughnessLengthsAndForcHeightsNonLake :601-637, CalcInitialTemperatureAndEnergyVars :304-359, calc_soilevap_resis
__global__ void __launch_bounds__(256)
preflux_col_kernel(PreFluxDev f, PrePrm prm, PreGeo g, int numf, const int32_t* __restrict__ filterc, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int c1 = filterc[fc];
  const int cc = c1 - g.begc0;
  const size_t ldc = (size_t)g.ldc;
  const int lt = f.lun_itype[cc];
  if (is_urban(lt)) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
  const double frac_sno = f.frac_sno[cc];
  double z0mg;
  if (prm.z0param_method == 1) {
    z0mg = (frac_sno > 0.0) ? prm.zsno : prm.zlnd;
  } else {
    if (frac_sno > 0.0) {
      if (prm.use_z0m_snowmelt) {
        const double sm = f.snomelt_accum[cc];
        if (sm < 1.e-5) z0mg = dexp(-b1_param * rpi * 0.5 + b4_param) * 1.e-3;
        else z0mg = dexp(b1_param * (atan((log10(sm) + meier_param1) / meier_param2)) + b4_param) * 1.e-3;
      } else {
        z0mg = prm.zsno;
      }
    } else if (lt == CTSM_ISTICE) {
      z0mg = prm.zglc;
    } else {
      z0mg = prm.zlnd;
    }
  }
  f.z0mg[cc] = z0mg; f.z0hg[cc] = z0mg; f.z0qg[cc] = z0mg;
  // tssbef(c,j) = t_soisno(c,j), all 37 levels
  double t1 = 0.0;
#pragma unroll 4
  for (int j = SNO_LO; j <= CTSM_NLEVGRND; ++j) {
    const double t = f.t_soisno[(size_t)(j - SNO_LO) * ldc + cc];
    f.t_ssbef[(size_t)(j - SNO_LO) * ldc + cc] = t;
    if (j == 1) t1 = t;
  }
  const int snl = f.snl[cc];
  const double t_h2osfc = f.t_h2osfc[cc], frac_h2osfc = f.frac_h2osfc[cc];
  f.t_h2osfc_bef[cc] = t_h2osfc;
  if (snl < 0) {
    const double fse = f.frac_sno_eff[cc];
    f.t_grnd[cc] = fse * f.t_soisno[(size_t)(snl + 1 - SNO_LO) * ldc + cc] + (1.0 - fse - frac_h2osfc) * t1 + frac_h2osfc * t_h2osfc;
  } else {
    f.t_grnd[cc] = (1 - frac_h2osfc) * t1 + frac_h2osfc * t_h2osfc;
  }
  if (lt == CTSM_ISTICE) f.emg[cc] = 0.97;
  else f.emg[cc] = (1.0 - frac_sno) * 0.96 + frac_sno * 0.97;
  const size_t otop = (size_t)(snl + 1 - SNO_LO) * ldc + cc;
  f.htvp[cc] = (f.h2osoi_liq[otop] <= 0.0 && f.h2osoi_ice[otop] > 0.0) ? hsub : hvap;
  f.beta[cc] = 1.0;
  f.zii[cc] = 1000.0;
  f.thv[cc] = f.forc_th[cc] * (1.0 + 0.61 * f.forc_q[cc]);

  // calc_soilevap_resis, SurfaceResistanceMod.F90:192-426
  const size_t o1 = (size_t)(1 - SNO_LO) * ldc + cc;
  const double liq1 = f.h2osoi_liq[o1], ice1 = f.h2osoi_ice[o1], dz1 = f.dz[o1], watsat1 = f.watsat[cc];
  if (prm.soil_resis_method == 0) {                                            // calc_beta_leepielke1992 :279-313
    if (lt != CTSM_ISTWET && lt != CTSM_ISTICE) {
      if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP) {
        const double wx = (liq1 / denh2o + ice1 / denice) / dz1;
        const double watfc1 = f.watfc[cc];
        if (wx < watfc1) {
          double fac_fc = fmin(1.0, wx / watfc1);
          fac_fc = fmax(fac_fc, 0.01);
          const double t = (1.0 - cos(rpi * fac_fc));
          f.soilbeta[cc] = (1.0 - frac_sno - frac_h2osfc) * 0.25 * (t * t) + frac_sno + frac_h2osfc;
        } else {
          f.soilbeta[cc] = 1.0;
        }
      }
    } else {
      f.soilbeta[cc] = 1.0;
    }
  } else {                                                                     // calc_soil_resistance_sl14 :388-424
    if (lt != CTSM_ISTWET && lt != CTSM_ISTICE) {
      if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP) {
        const double bsw1 = f.bsw[cc], sucsat1 = f.sucsat[cc];
        const double vwc_liq = fmax(liq1, 1.0e-6) / (dz1 * denh2o);
        const double eff_por_top = fmax(0.01, watsat1 - fmin(watsat1, ice1 / (dz1 * denice)));
        const double aird = watsat1 * pw(sucsat1 / 1.e7, R4(1.) / bsw1);
        const double d0 = R4(2.12e-5) * pw(t1 / R4(273.15), R4(1.75));
        const double eps = watsat1 - aird;
        const double dg = eps * d0 * pw(eps / watsat1, 3.0 / fmax(3.0, bsw1));
        double dsl = prm.d_max * fmax(0.001, (prm.frac_sat_soil_dsl_init * eff_por_top - vwc_liq))
                     / fmax(0.001, (prm.frac_sat_soil_dsl_init * watsat1 - aird));
        dsl = fmax(dsl, 0.0);
        dsl = fmin(dsl, 200.0);
        f.dsl[cc] = dsl;
        double sr = dsl / (dg * eps * R4(1.e3)) + 20.0;
        sr = fmin(1.e6, sr);
        f.soilresis[cc] = sr;
      }
    } else {
      f.soilresis[cc] = 0.0;
    }
  }
}

// SetRoughnessLengthsAndForcHeightsNonLake :639-681, CalcInitialTemperatureAndEnergyVars :361-401
__global__ void __launch_bounds__(256)
preflux_patch_b_kernel(PreFluxDev f, PreGeo g, int numf, const int32_t* __restrict__ filterp) {
  const int fp = blockIdx.x * blockDim.x + threadIdx.x;
  if (fp >= numf) return;
  const int pp = filterp[fp] - g.begp0;
  const int cc = f.column[pp] - g.begc0;
  const int gg = f.gridcell[pp] - g.begg0;
  const int lt = f.lun_itype[cc];
  const double z0mv = f.z0m[pp];
  f.z0mv[pp] = z0mv; f.z0hv[pp] = z0mv; f.z0qv[pp] = z0mv;
  f.z0mg_p[pp] = spval; f.z0hg_p[pp] = spval; f.z0qg_p[pp] = spval; f.kbm1[pp] = spval;
  const double displa = f.displa[pp];
  const bool rural = (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP);
  double hgt_t = f.forc_hgt_t_patch[pp];
  if (rural && f.frac_veg_nosno[pp] != 0) {
    f.forc_hgt_u_patch[pp] = f.forc_hgt_u[gg] + z0mv + displa;
    hgt_t = f.forc_hgt_t[gg] + z0mv + displa;
    f.forc_hgt_t_patch[pp] = hgt_t;
    f.forc_hgt_q_patch[pp] = f.forc_hgt_q[gg] + z0mv + displa;
  } else if (rural || lt == CTSM_ISTWET || lt == CTSM_ISTICE) {
    f.forc_hgt_u_patch[pp] = f.forc_hgt_u[gg] + f.z0mg[cc] + displa;
    hgt_t = f.forc_hgt_t[gg] + f.z0hg[cc] + displa;
    f.forc_hgt_t_patch[pp] = hgt_t;
    f.forc_hgt_q_patch[pp] = f.forc_hgt_q[gg] + f.z0qg[cc] + displa;
  }
  f.eflx_sh_tot[pp] = 0.0;
  if (rural) f.eflx_sh_tot_r[pp] = 0.0;
  f.eflx_lh_tot[pp] = 0.0;
  if (rural) f.eflx_lh_tot_r[pp] = 0.0;
  f.eflx_sh_veg[pp] = 0.0;
  f.cgrnd[pp] = 0.0; f.cgrnds[pp] = 0.0; f.cgrndl[pp] = 0.0;
  const double avmuir = 1.0;
  f.emv[pp] = 1.0 - dexp(-(f.elai[pp] + f.esai[pp]) / avmuir);
  f.thm[pp] = f.forc_t[cc] + 0.0098 * hgt_t;
}

// CalculateSurfaceHumidity :112-236
__global__ void __launch_bounds__(256)
surface_humidity_kernel(SurfHumDev f, int begc0, int ldc_, int numf, const int32_t* __restrict__ filterc, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int c1 = filterc[fc];
  const int cc = c1 - begc0;
  const size_t ldc = (size_t)ldc_;
  const int lt = f.lun_itype[cc], snl = f.snl[cc];
  if (is_urban(lt)) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
  const size_t o1 = (size_t)(1 - SNO_LO) * ldc + cc;
  const double t1 = f.t_soisno[o1], pbot = f.forc_pbot[cc], forc_q = f.forc_q[cc];
  const double fse = f.frac_sno_eff[cc], fh = f.frac_h2osfc[cc];
  double qred = 1.0, hr = 0.0;
  const bool rural = (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP);
  if (lt != CTSM_ISTWET && lt != CTSM_ISTICE) {
    if (rural) {
      const double wx = (f.h2osoi_liq[o1] / denh2o + f.h2osoi_ice[o1] / denice) / f.dz[o1];
      double fac = fmin(1.0, wx / f.watsat[cc]);
      fac = fmax(fac, 0.01);
      double psit = -f.sucsat[cc] * pw(fac, -f.bsw[cc]);
      psit = fmax(f.smpmin[cc], psit);
      hr = dexp(psit / roverg / t1);
      qred = (1.0 - fse - fh) * hr + fse + fh;
      f.soilalpha[cc] = qred;
    }
  } else {
    f.soilalpha[cc] = spval;
  }
  if (rural) {
    QS q = qsat(t1, pbot, true);
    double qsatg = q.qs, qsatgdT_soil = q.qsdT;
    if (qsatg > forc_q && forc_q > hr * qsatg) { qsatg = forc_q; qsatgdT_soil = 0.0; }
    const double qg_soil = hr * qsatg;
    double qg_snow, dqgdT, qg_h2osfc;
    if (snl < 0) {
      q = qsat(f.t_soisno[(size_t)(snl + 1 - SNO_LO) * ldc + cc], pbot, true);
      qg_snow = q.qs;
      dqgdT = fse * q.qsdT + (1.0 - fse - fh) * hr * qsatgdT_soil;
    } else {
      qg_snow = qg_soil;
      dqgdT = (1.0 - fh) * hr * qsatgdT_soil;
    }
    if (fh > 0.0) {
      q = qsat(f.t_h2osfc[cc], pbot, true);
      qg_h2osfc = q.qs;
      dqgdT = dqgdT + fh * q.qsdT;
    } else {
      qg_h2osfc = qg_soil;
    }
    f.qg_soil[cc] = qg_soil; f.qg_snow[cc] = qg_snow; f.qg_h2osfc[cc] = qg_h2osfc; f.dqgdT[cc] = dqgdT;
    f.qg[cc] = fse * qg_snow + (1.0 - fse - fh) * qg_soil + fh * qg_h2osfc;
  } else {
    const QS q = qsat(f.t_grnd[cc], pbot, true);
    double qg = qred * q.qs, dqgdT = qred * q.qsdT;
    if (q.qs > forc_q && forc_q > qred * q.qs) { qg = forc_q; dqgdT = 0.0; }
    f.qg[cc] = qg; f.dqgdT[cc] = dqgdT; f.qg_snow[cc] = qg; f.qg_soil[cc] = qg; f.qg_h2osfc[cc] = qg;
  }
}

// dewpoint, BareGroundFluxesMod.F90:531-577
__device__ __forceinline__ double dewpoint(double e, double t) {
  double d;
  if (t < tfrz) d = 273.86 * dlog(e / 611.21) / (22.587 - dlog(e / 611.21));
  else d = 243.04 * dlog(e / 610.94) / (17.625 - dlog(e / 610.94));
  return d + tfrz;
}

// BareGroundFluxes :281-525 for one patch
__global__ void __launch_bounds__(128)
bareground_kernel(BareDev f, PrePrm prm, PreGeo g, int numf, const int32_t* __restrict__ filterp, DevStatus* ds) {
  const int fi = blockIdx.x * blockDim.x + threadIdx.x;
  if (fi >= numf) return;
  const int p1 = filterp[fi];
  const int pp = p1 - g.begp0;
  const int c1 = f.column[pp];
  const int cc = c1 - g.begc0;
  const int gg = f.gridcell[pp] - g.begg0;
  const size_t ldc = (size_t)g.ldc, ldp = (size_t)g.ldp;
  const int lt = f.lun_itype[cc];
  if (is_urban(lt)) { report_failure(ds, p1, CTSM_ERR_URBAN, 0); return; }
  const double forc_t = f.forc_t[cc], forc_pbot = f.forc_pbot[cc], forc_q = f.forc_q[cc], forc_th = f.forc_th[cc];
  const double forc_rho = f.forc_rho[cc], thm = f.thm[pp], t_grnd = f.t_grnd[cc], thv = f.thv[cc];
  f.btran[pp] = 0.0;
  f.t_veg[pp] = forc_t;
  const double cf_bare = forc_pbot / (rgas * 0.001 * thm) * 1.e06;
  f.rssun[pp] = 1.0 / 1.e15 * cf_bare;
  f.rssha[pp] = 1.0 / 1.e15 * cf_bare;
#pragma unroll 5
  for (int j = 0; j < CTSM_NLEVGRND; ++j) { f.rootr[(size_t)j * ldp + pp] = 0.0; f.rresis[(size_t)j * ldp + pp] = 0.0; }
  const double displa = 0.0;
  f.displa[pp] = 0.0; f.z0mv[pp] = 0.0; f.z0hv[pp] = 0.0; f.z0qv[pp] = 0.0;
  f.dlrad[pp] = 0.0; f.ulrad[pp] = 0.0; f.dhsdt_canopy[pp] = 0.0; f.eflx_sh_stem[pp] = 0.0;
  const double fu = f.forc_u[gg], fvv = f.forc_v[gg];
  const double ur = fmax(prm.wind_min, sqrt(fu * fu + fvv * fvv));
  const double dth = thm - t_grnd;
  const double dqh = forc_q - f.qg[cc];
  const double dthv = dth * (1.0 + 0.61 * forc_q) + 0.61 * forc_th * dqh;
  double hgt_u = f.forc_hgt_u_patch[pp], hgt_t = f.forc_hgt_t_patch[pp], hgt_q = f.forc_hgt_q_patch[pp];
  const double zldis = hgt_u;
  const double z0mg = f.z0mg[cc];
  double z0hg = f.z0hg[cc], z0qg = f.z0qg[cc];
  double um, obu;
  {                                                                            // MoninObukIni, FrictionVelocityMod.F90:1187-1207
    if (dthv >= 0.0) um = fmax(ur, 0.1);
    else um = sqrt(ur * ur + 0.5 * 0.5);
    const double rib = grav * zldis * dthv / (thv * um * um);
    double zeta;
    if (rib >= 0.0) {
      zeta = rib * dlog(zldis / z0mg) / (1.0 - 5.0 * fmin(rib, 0.19));
      zeta = fmin(prm.zetamaxstable, fmax(zeta, 0.01));
    } else {
      zeta = rib * dlog(zldis / z0mg);
      zeta = fmax(-100.0, fmin(zeta, -0.01));
    }
    obu = zldis / zeta;
  }
  const double hu_g = f.forc_hgt_u[gg], ht_g = f.forc_hgt_t[gg], hq_g = f.forc_hgt_q[gg];
  const double beta = f.beta[cc], zii = f.zii[cc];
  FricOut fo;
  double fm = 0.0, zeta = 0.0;
  for (int iter = 1; iter <= 3; ++iter) {                                      // niters = 3, :341-398
    fo = friction_velocity(hgt_u, hgt_t, hgt_q, displa, z0mg, z0hg, z0qg, obu, iter, ur, um, fm);
    fm = fo.fm;
    const double umf = um;                                                     // va(n) = um of this pass (FrictionVelocity :904)
    const double ustar = fo.ustar;
    const double tstar = fo.temp1 * dth;
    const double qstar = fo.temp2 * dqh;
    if (prm.z0param_method == 1) z0hg = z0mg / dexp(prm.a_coef * pw(ustar * z0mg / nu_param, prm.a_exp));
    else z0hg = meier_param3 * nu_param / ustar * dexp(-beta_param * pw(ustar, 0.5) * pw(fabs(tstar), 0.25));
    z0qg = z0hg;
    hgt_u = hu_g + z0mg + displa;
    hgt_t = ht_g + z0hg + displa;
    hgt_q = hq_g + z0qg + displa;
    const double thvstar = tstar * (1.0 + 0.61 * forc_q) + 0.61 * forc_th * qstar;
    zeta = zldis * vkc * grav * thvstar / ((ustar * ustar) * thv);
    if (zeta >= 0.0) {
      zeta = fmin(prm.zetamaxstable, fmax(zeta, 0.01));
      um = fmax(ur, 0.1);
    } else {
      zeta = fmax(-100.0, fmin(zeta, -0.01));
      const double wc = beta * pw(-grav * ustar * thvstar * zii / thv, 0.333);
      um = sqrt(ur * ur + wc * wc);
    }
    obu = zldis / zeta;
    if (iter == 3) f.va[pp] = umf;
  }
  const double ustar = fo.ustar;
  f.forc_hgt_u_patch[pp] = hgt_u; f.forc_hgt_t_patch[pp] = hgt_t; f.forc_hgt_q_patch[pp] = hgt_q;
  f.z0mg_p[pp] = z0mg; f.z0hg_p[pp] = z0hg; f.z0qg_p[pp] = z0qg;
  f.um[pp] = um; f.obu[pp] = obu; f.zeta[pp] = zeta; f.ustar[pp] = ustar; f.num_iter[pp] = 3.0;
  f.vds[pp] = fo.vds; f.u10[pp] = fo.u10; f.u10_clm[pp] = fo.u10_clm; f.fv[pp] = ustar;

  // :402-525
  const double ram = 1.0 / (ustar * ustar / um);
  const double rah = 1.0 / (fo.temp1 * ustar);
  const double raw = 1.0 / (fo.temp2 * ustar);
  const double raih = forc_rho * cpair / rah;
  const QS qf = qsat(forc_t, forc_pbot, false);
  const double forc_e = fmax((forc_q * forc_pbot) / (forc_q + 0.622), 0.01 * qf.es);
  const double forc_dewpoint = dewpoint(forc_e, t_grnd);
  double raiw = 0.0;
  if (dqh > 0.0) {
    if (t_grnd > forc_dewpoint) raiw = 0.0;
    else raiw = forc_rho / (raw);
  } else {
    if (prm.soil_resis_method == 0) {
      if (t_grnd > forc_dewpoint) raiw = 0.0;
      else raiw = f.soilbeta[cc] * forc_rho / (raw);
    }
    if (prm.soil_resis_method == 1) raiw = forc_rho / (raw + f.soilresis[cc]);
  }
  f.ram1[pp] = ram;
  const double cgrnds = raih, cgrndl = raiw * f.dqgdT[cc];
  f.cgrnds[pp] = cgrnds; f.cgrndl[pp] = cgrndl;
  f.cgrnd[pp] = cgrnds + f.htvp[cc] * cgrndl;
  f.taux[pp] = -forc_rho * fu / ram;
  f.tauy[pp] = -forc_rho * fvv / ram;
  const double sh = -raih * dth;
  f.eflx_sh_grnd[pp] = sh; f.eflx_sh_tot[pp] = sh;
  const int snl = f.snl[cc];
  f.eflx_sh_snow[pp] = -raih * (thm - f.t_soisno[(size_t)(snl + 1 - SNO_LO) * ldc + cc]);
  f.eflx_sh_soil[pp] = -raih * (thm - f.t_soisno[(size_t)(1 - SNO_LO) * ldc + cc]);
  f.eflx_sh_h2osfc[pp] = -raih * (thm - f.t_h2osfc[cc]);
  f.qflx_tran_veg[pp] = 0.0; f.qflx_evap_veg[pp] = 0.0;
  const double ev = -raiw * dqh;
  f.qflx_evap_soi[pp] = ev; f.qflx_evap_tot_patch[pp] = ev;
  f.qflx_ev_snow[pp] = -raiw * (forc_q - f.qg_snow[cc]);
  f.qflx_ev_soil[pp] = -raiw * (forc_q - f.qg_soil[cc]);
  f.qflx_ev_h2osfc[pp] = -raiw * (forc_q - f.qg_h2osfc[cc]);
  const double t_ref2m = thm + fo.temp1 * dth * (1.0 / fo.temp12m - 1.0 / fo.temp1);
  f.t_ref2m[pp] = t_ref2m;
  const double q_ref2m = forc_q + fo.temp2 * dqh * (1.0 / fo.temp22m - 1.0 / fo.temp2);
  f.q_ref2m[pp] = q_ref2m;
  const QS q2 = qsat(t_ref2m, forc_pbot, false);
  const double rh = fmin(100.0, q_ref2m / q2.qs * 100.0);
  f.rh_ref2m[pp] = rh;
  const bool rural = (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP);
  if (rural) { f.rh_ref2m_r[pp] = rh; f.t_ref2m_r[pp] = t_ref2m; }
  f.kbm1[pp] = dlog(z0mg / z0hg);
  // z0hg_col(c) / z0qg_col(c) are copied back by bareground_colcopy_kernel once every patch has read them
  if (prm.human_fast) {                                                        // fast human-stress indices :475-506
    const double tc = t_ref2m - tfrz;                                          // KtoC, HumanIndexMod.F90:1205
    f.tc_ref2m[pp] = tc;
    const double vap = (rh / 100.0) * q2.es;                                   // VaporPres :1243
    f.vap_ref2m[pp] = vap;
    if (rh < 0.0 || rh > 100.0) { report_failure(ds, p1, CTSM_ERR_RH, 0); return; }    // Wet_BulbS :1016-1022
    const double wbt = tc * atan(0.151977 * sqrt(rh + 8.313659)) + atan(tc + rh) - atan(rh - 1.676331)
                       + 0.00391838 * pow(rh, (3.0 / 2.0)) * atan(0.023101 * rh) - 4.686035;
    const double tf = (tc) * 9.0 / 5.0 + 32.0;                                 // HeatIndex :1039-1095
    double hi;
    if (tf < 68.0) hi = tf;
    else hi = -42.379 + 2.04901523 * tf + 10.14333127 * rh + (-0.22475541 * tf * rh) + (-6.83783e-3 * (tf * tf))
              + (-5.481717e-2 * (rh * rh)) + 1.22874e-3 * (tf * tf) * rh + 8.5282e-4 * tf * (rh * rh)
              + (-1.99e-6 * (tf * tf) * (rh * rh));
    hi = (hi - 32.0) * 5.0 / 9.0;
    const double at = tc + 3.30 * vap / 1000.0 - 0.70 * fo.u10_clm - 4.0;      // AppTemp :555
    const double sw = 0.567 * (tc) + 0.393 * vap / 100.0 + 3.94;               // swbgt :596
    const double hx = tc + ((5.0 / 9.0) * (vap / 100.0 - 10.0));               // hmdex :637
    const double Tc = fmin(tc, 50.0);                                          // dis_coiS :715-761
    double rhl = fmin(rh, 99.0);
    rhl = fmax(rhl, 5.0);
    const double rh_min = Tc * (-2.27) + 27.7;
    const double dc = (Tc < -20.0 || rhl < rh_min) ? Tc : 0.5 * wbt + 0.5 * Tc;
    f.wbt_ref2m[pp] = wbt; f.nws_hi_ref2m[pp] = hi; f.appar_temp_ref2m[pp] = at; f.swbgt_ref2m[pp] = sw;
    f.humidex_ref2m[pp] = hx; f.discomf_index_ref2mS[pp] = dc;
    if (rural) {
      f.wbt_ref2m_r[pp] = wbt; f.nws_hi_ref2m_r[pp] = hi; f.appar_temp_ref2m_r[pp] = at; f.swbgt_ref2m_r[pp] = sw;
      f.humidex_ref2m_r[pp] = hx; f.discomf_index_ref2mS_r[pp] = dc;
    }
  }
}

// "Copy local patch ground roughness back to column arrays" :466-469.  The reference's sequential loop leaves the value of
// the column's LAST filter patch; a separate launch, so that no patch of the column reads the column value after it changed.
__global__ void __launch_bounds__(256)
bareground_colcopy_kernel(BareDev f, PreGeo g, int numf, const int32_t* __restrict__ filterp) {
  const int fi = blockIdx.x * blockDim.x + threadIdx.x;
  if (fi >= numf) return;
  const int pp = filterp[fi] - g.begp0;
  const int c1 = f.column[pp];
  const bool last_of_column = (fi + 1 >= numf) || (f.column[filterp[fi + 1] - g.begp0] != c1);
  if (last_of_column) { f.z0hg[c1 - g.begc0] = f.z0hg_p[pp]; f.z0qg[c1 - g.begc0] = f.z0qg_p[pp]; }
}

PrePrm make_prm(const ctsm_params_t& p, int time_flags) {
  PrePrm q;
  q.z0param_method = p.z0param_method; q.soil_resis_method = p.soil_resis_method; q.use_z0m_snowmelt = p.use_z0m_snowmelt;
  q.time_flags = time_flags; q.human_fast = (p.calc_human_stress_indices == 1);
  q.zlnd = p.zlnd; q.zsno = p.zsno; q.zglc = p.zglc; q.d_max = p.d_max; q.frac_sat_soil_dsl_init = p.frac_sat_soil_dsl_init;
  q.a_coef = p.a_coef; q.a_exp = p.a_exp; q.wind_min = p.wind_min; q.zetamaxstable = p.zetamaxstable;
  return q;
}
PreGeo make_geo(const ctsm_bounds_t& a) {
  PreGeo g;
  g.begc0 = a.begc; g.begp0 = a.begp; g.begg0 = a.begg; g.ldc = a.endc - a.begc + 1; g.ldp = a.endp - a.begp + 1;
  return g;
}

}  // namespace

extern "C" int ctsm_b200_biogeophys_pre_flux_calcs(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec,
                                                   const int32_t* filter_nolakec, int num_nolakep, const int32_t* filter_nolakep,
                                                   int num_urbanc, const int32_t* filter_urbanc, int time_flags,
                                                   const ctsm_preflux_fields_t* hf, int mem, ctsm_status_t* st) {
  (void)filter_urbanc;
  if (!ctx || !bounds || !hf || num_nolakec < 0 || num_nolakep < 0 || (num_nolakec > 0 && !filter_nolakec) ||
      (num_nolakep > 0 && !filter_nolakep))
    return CTSM_ERR_BAD_ARG;
  if (num_urbanc != 0) return CTSM_ERR_URBAN;                  // urban columns are outside the hot path
  CUDA_TRY(cudaSetDevice(ctx->device));
  PreFluxDev d;
  const int32_t *dfc = filter_nolakec, *dfp = filter_nolakep;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_PREFLUX
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_PREFLUX
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_nolakec, num_nolakec, &dfc);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter1, filter_nolakep, num_nolakep, &dfp);
    if (rc) return rc;
  }
  const PrePrm prm = make_prm(ctx->prm, time_flags);
  const PreGeo g = make_geo(hf->alloc);
  if (num_nolakep > 0) {
    preflux_patch_a_kernel<<<grid_for(num_nolakep, 256), 256, 0, ctx->stream>>>(d, prm, g, num_nolakep, dfp);
    ctx->launches++;
  }
  if (num_nolakec > 0) {
    preflux_col_kernel<<<grid_for(num_nolakec, 256), 256, 0, ctx->stream>>>(d, prm, g, num_nolakec, dfc, ctx->d_status);
    ctx->launches++;
  }
  if (num_nolakep > 0) {
    preflux_patch_b_kernel<<<grid_for(num_nolakep, 256), 256, 0, ctx->stream>>>(d, g, num_nolakep, dfp);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

extern "C" int ctsm_b200_calculate_surface_humidity(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec,
                                                    const int32_t* filter_nolakec, const ctsm_surfacehumidity_fields_t* hf,
                                                    int mem, ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_nolakec < 0 || (num_nolakec > 0 && !filter_nolakec)) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  SurfHumDev d;
  const int32_t* dfc = filter_nolakec;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_SURFACEHUMIDITY
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SURFACEHUMIDITY
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_nolakec, num_nolakec, &dfc);
    if (rc) return rc;
  }
  if (num_nolakec > 0) {
    surface_humidity_kernel<<<grid_for(num_nolakec, 256), 256, 0, ctx->stream>>>(d, hf->alloc.begc, hf->alloc.endc - hf->alloc.begc + 1,
                                                                                 num_nolakec, dfc, ctx->d_status);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

extern "C" int ctsm_b200_bare_ground_fluxes(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_noexposedvegp,
                                            const int32_t* filter_noexposedvegp, const ctsm_baregroundfluxes_fields_t* hf,
                                            int mem, ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_noexposedvegp < 0 || (num_noexposedvegp > 0 && !filter_noexposedvegp)) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  BareDev d;
  const int32_t* dfp = filter_noexposedvegp;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_BAREGROUNDFLUXES
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_BAREGROUNDFLUXES
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_noexposedvegp, num_noexposedvegp, &dfp);
    if (rc) return rc;
  }
  if (num_noexposedvegp > 0) {
    bareground_kernel<<<grid_for(num_noexposedvegp, 128), 128, 0, ctx->stream>>>(d, make_prm(ctx->prm, 0), make_geo(hf->alloc),
                                                                                 num_noexposedvegp, dfp, ctx->d_status);
    bareground_colcopy_kernel<<<grid_for(num_noexposedvegp, 256), 256, 0, ctx->stream>>>(d, make_geo(hf->alloc), num_noexposedvegp, dfp);
    ctx->launches += 2;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}
