// hydrology.cu — the soil-hydrology routines of HydrologyNoDrainage around SoilWater on B200:
//   ctsm_b200_hydrology_infiltration   SetSoilWaterFractions ... TotalSurfaceRunoff   HydrologyNoDrainageMod.F90:297-337
//   ctsm_b200_water_table              PerchedWaterTable, ThetaBasedWaterTable, RenewCondensation   :359-373
//   ctsm_b200_hydrology_diagnostics    the inline diagnostics that close the routine   :420-757
// All are maps over column filters with no coupling between columns: one thread per filter entry on the Fortran arrays
// (level-major, column fastest: a warp reads 32 consecutive doubles per level), HBM-bound.
#include "surface_layer.cuh"
#include "common.cuh"
#include <vector>

namespace {
constexpr int SNO_LO = -CTSM_NLEVSNO + 1;
__device__ __forceinline__ bool is_urban(int lt) { return lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX; }
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// The surface-water / infiltration chain of HydrologyNoDrainage (HydrologyNoDrainageMod.F90:297-337; SURVEY.md 8f rank 3,
// first part): nine small per-column routines of the reference, none of which reads another column, fused into one
// thread-per-column kernel (every intermediate stays in registers and is also stored, because all of them are history /
// balance fields of the reference).  HBM-bound: ~0.9 KB per hydrology column.
struct InfilDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_INFILTRATION
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_INFILTRATION
#undef CTSM_F
};

namespace {
struct InfilPrm { double dtime, fff, pc, mu, e_ice; int h2osfcflag, crop_fsat_equals_zero; };

// SetFloodc, SoilHydrologyMod.F90:282-291
__global__ void __launch_bounds__(256)
floodc_kernel(InfilDev f, int begc0, int begg0, int numf, const int32_t* __restrict__ filterc, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int c1 = filterc[fc];
  const int cc = c1 - begc0;
  if (is_urban(f.lun_itype[cc])) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
  f.qflx_floodc[cc] = f.forc_flood[f.col_gridcell[cc] - begg0];
}

// truncate_small_values, NumericsMod.F90:50
__device__ __forceinline__ double truncate_small(double data, double baseline) {
  return (fabs(data) < 1.e-13 * fabs(baseline)) ? 0.0 : data;
}

__global__ void __launch_bounds__(256)
infiltration_kernel(InfilDev f, InfilPrm prm, int begc0, int ldc_, int numf, const int32_t* __restrict__ filterc, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int c1 = filterc[fc];
  const int cc = c1 - begc0;
  const size_t ldc = (size_t)ldc_;
  const int lt = f.lun_itype[cc];
  if (is_urban(lt)) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
  const double dtime = prm.dtime;
  // SetSoilWaterFractions :239-252 (excess_ice = 0); the three top-level ice fractions feed qinmax below
  double qmin = 0.0;
#pragma unroll 4
  for (int j = 1; j <= CTSM_NLEVSOI; ++j) {
    const size_t o1 = (size_t)(j - 1) * ldc + cc, os = (size_t)(j - SNO_LO) * ldc + cc;
    const double watsat = f.watsat[o1];
    const double dz_ext = f.dz[os] + 0.0 / denice;
    const double vol_ice = fmin(watsat, (f.h2osoi_ice[os] + 0.0) / (dz_ext * denice));
    f.eff_porosity[o1] = fmax(0.01, watsat - vol_ice);
    const double icefrac = fmin(1.0, vol_ice / watsat);
    f.icefrac[o1] = icefrac;
    if (j <= 3) {                                                            // ComputeQinmaxHksat :296-300
      const double v = pow(10.0, -prm.e_ice * (icefrac)) * f.hksat[o1];
      if (j == 1 || v < qmin) qmin = v;
    }
  }
  // SaturatedExcessRunoff: ComputeFsatTopmodel :344-356
  const double frost_table = f.frost_table[cc], zwt = f.zwt[cc], zwt_perched = f.zwt_perched[cc];
  double fsat;
  if (frost_table > zwt_perched && frost_table <= zwt) fsat = f.wtfact[cc] * dexp(-0.5 * prm.fff * zwt_perched);
  else fsat = f.wtfact[cc] * dexp(-0.5 * prm.fff * zwt);
  if (prm.crop_fsat_equals_zero && lt == CTSM_ISTCROP) fsat = 0.0;
  f.fsat[cc] = fsat; f.fcov[cc] = fsat;
  const double rain = f.qflx_rain_plus_snomelt[cc];
  const double qflx_sat_excess_surf = fsat * rain;
  f.qflx_sat_excess_surf[cc] = qflx_sat_excess_surf;
  // SetQflxInputs :339-362
  const double frac_h2osfc = f.frac_h2osfc[cc];
  const double qflx_top_soil = rain + f.qflx_snow_h2osfc[cc] + f.qflx_floodc[cc];
  f.qflx_top_soil[cc] = qflx_top_soil;
  double fsno, qflx_evap;
  if (f.snl[cc] >= 0) { fsno = 0.0; qflx_evap = f.qflx_liqevap_from_top_layer[cc]; }
  else { fsno = f.frac_sno_eff[cc]; qflx_evap = f.qflx_ev_soil_col[cc]; }
  double qflx_in_soil = (1.0 - frac_h2osfc) * (qflx_top_soil - qflx_sat_excess_surf);
  double qflx_top_soil_to_h2osfc = frac_h2osfc * (qflx_top_soil - qflx_sat_excess_surf);
  qflx_in_soil = qflx_in_soil - (1.0 - fsno - frac_h2osfc) * qflx_evap;
  qflx_top_soil_to_h2osfc = qflx_top_soil_to_h2osfc - frac_h2osfc * f.qflx_ev_h2osfc_col[cc];
  f.qflx_in_soil[cc] = qflx_in_soil; f.qflx_top_soil_to_h2osfc[cc] = qflx_top_soil_to_h2osfc;
  // InfiltrationExcessRunoff :253-259
  const double qinmax = (1.0 - fsat) * qmin;
  f.qinmax[cc] = qinmax;
  const double qflx_infl_excess = fmax(0.0, (qflx_in_soil - (1.0 - frac_h2osfc) * qinmax));
  f.qflx_infl_excess[cc] = qflx_infl_excess;
  // RouteInfiltrationExcess :399-419
  double qflx_in_soil_limited, qflx_in_h2osfc, qflx_infl_excess_surf;
  if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP) {
    qflx_in_soil_limited = qflx_in_soil - qflx_infl_excess;
    if (prm.h2osfcflag != 0) { qflx_in_h2osfc = qflx_top_soil_to_h2osfc + qflx_infl_excess; qflx_infl_excess_surf = 0.0; }
    else { qflx_in_h2osfc = qflx_top_soil_to_h2osfc; qflx_infl_excess_surf = qflx_infl_excess; }
  } else {
    qflx_in_soil_limited = qflx_in_soil; qflx_in_h2osfc = 0.0; qflx_infl_excess_surf = 0.0;
  }
  f.qflx_in_soil_limited[cc] = qflx_in_soil_limited; f.qflx_in_h2osfc[cc] = qflx_in_h2osfc;
  f.qflx_infl_excess_surf[cc] = qflx_infl_excess_surf;
  // UpdateH2osfc, SurfaceWaterMod.F90:345-556
  const double h2osfc0 = f.h2osfc[cc], thresh = f.h2osfc_thresh[cc];
  double frac_infclust = 0.0;
  if (prm.h2osfcflag == 1) {
    const double fn = f.frac_h2osfc_nosnow[cc];
    if (fn <= prm.pc) frac_infclust = 0.0;
    else frac_infclust = pow(fn - prm.pc, prm.mu);
  }
  double qflx_h2osfc_surf;
  if (h2osfc0 > thresh && prm.h2osfcflag != 0) {
    const double k_wet = 1.0e-4 * sin((rpi / 180.0) * f.topo_slope[cc]);
    qflx_h2osfc_surf = k_wet * frac_infclust * (h2osfc0 - thresh);
    qflx_h2osfc_surf = fmin(qflx_h2osfc_surf, (h2osfc0 - thresh) / dtime);
  } else {
    qflx_h2osfc_surf = 0.0;
  }
  if (qflx_h2osfc_surf < R4(1.0e-8f)) qflx_h2osfc_surf = 0.0;              // SurfaceWaterMod.F90:499: a default-kind (REAL(4)) literal
  f.qflx_h2osfc_surf[cc] = qflx_h2osfc_surf;
  double h2osfc_partial = h2osfc0 + (qflx_in_h2osfc - qflx_h2osfc_surf) * dtime;
  h2osfc_partial = truncate_small(h2osfc_partial, h2osfc0);
  double qflx_h2osfc_drain;
  if (h2osfc_partial < 0.0) {
    qflx_h2osfc_drain = h2osfc_partial / dtime;
  } else {
    qflx_h2osfc_drain = fmin(frac_h2osfc * qinmax, h2osfc_partial / dtime);
    if (prm.h2osfcflag == 0) qflx_h2osfc_drain = fmax(0.0, h2osfc_partial / dtime);
  }
  f.qflx_h2osfc_drain[cc] = qflx_h2osfc_drain;
  double h2osfc = h2osfc_partial - qflx_h2osfc_drain * dtime;
  h2osfc = truncate_small(h2osfc, h2osfc_partial);
  f.h2osfc[cc] = h2osfc;
  // Infiltration :450-453, TotalSurfaceRunoff :511-515
  f.qflx_infl[cc] = qflx_in_soil_limited + qflx_h2osfc_drain;
  f.qflx_surf[cc] = qflx_sat_excess_surf + qflx_infl_excess_surf + qflx_h2osfc_surf;
}
}  // namespace

extern "C" int ctsm_b200_hydrology_infiltration(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec,
                                                const int32_t* filter_nolakec, int num_hydrologyc,
                                                const int32_t* filter_hydrologyc, int num_urbanc, const int32_t* filter_urbanc,
                                                const ctsm_infiltration_fields_t* hf, int mem, ctsm_status_t* st) {
  (void)filter_urbanc;
  if (!ctx || !bounds || !hf || num_nolakec < 0 || num_hydrologyc < 0 || (num_nolakec > 0 && !filter_nolakec) ||
      (num_hydrologyc > 0 && !filter_hydrologyc))
    return CTSM_ERR_BAD_ARG;
  if (num_urbanc != 0) return CTSM_ERR_URBAN;
  CUDA_TRY(cudaSetDevice(ctx->device));
  InfilDev d;
  const int32_t *dfn = filter_nolakec, *dfh = filter_hydrologyc;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_INFILTRATION
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_INFILTRATION
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_nolakec, num_nolakec, &dfn);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter1, filter_hydrologyc, num_hydrologyc, &dfh);
    if (rc) return rc;
  }
  const InfilPrm prm{ctx->prm.dtime, ctx->prm.fff, ctx->prm.pc, ctx->prm.mu, ctx->prm.e_ice, ctx->prm.h2osfcflag,
                     ctx->prm.crop_fsat_equals_zero};
  if (num_nolakec > 0) {
    floodc_kernel<<<grid_for(num_nolakec, 256), 256, 0, ctx->stream>>>(d, hf->alloc.begc, hf->alloc.begg, num_nolakec, dfn, ctx->d_status);
    ctx->launches++;
  }
  if (num_hydrologyc > 0) {
    infiltration_kernel<<<grid_for(num_hydrologyc, 256), 256, 0, ctx->stream>>>(d, prm, hf->alloc.begc, hf->alloc.endc - hf->alloc.begc + 1,
                                                                                num_hydrologyc, dfh, ctx->d_status);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// PerchedWaterTable SoilHydrologyMod.F90:1525-1641, ThetaBasedWaterTable :1933-2025, RenewCondensation :2569-2678
// (HydrologyNoDrainageMod.F90:359-373, use_aquifer_layer = .false.).  One thread per hydrology column; the three routines touch
// disjoint outputs of the same column (RenewCondensation changes layer-1 water AFTER both water tables have read it), so they
// run back to back in the thread.  Both water tables compare against sat_lev = 0.9 written as a default-kind literal (R4).
struct WaterTableDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_WATERTABLE
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_WATERTABLE
#undef CTSM_F
};
struct HydroDiagDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_HYDRODIAG
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_HYDRODIAG
#undef CTSM_F
};

namespace {
__global__ void __launch_bounds__(128)
water_table_kernel(WaterTableDev f, double dtime, int begc0, int ldc_, int numf, const int32_t* __restrict__ filterc, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int c1 = filterc[fc];
  const int cc = c1 - begc0;
  const size_t ldc = (size_t)ldc_;
  if (is_urban(f.lun_itype[cc])) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
  const double sat_lev = R4(0.9f);
  const double tfrz_ = 273.15;
#define SS(name, j) f.name[(size_t)((j) - SNO_LO) * ldc + cc]
#define GG(name, j) f.name[(size_t)((j) - 1) * ldc + cc]
#define ZI(j) f.zi[(size_t)((j) + CTSM_NLEVSNO) * ldc + cc]
#define VOL(k) (SS(h2osoi_liq, k) / (SS(dz, k) * denh2o) + SS(h2osoi_ice, k) / (SS(dz, k) * denice))
  // PerchedWaterTable :1581-1637
  {
    int k_frz = (SS(t_soisno, 1) > tfrz_) ? CTSM_NLEVSOI : 1;
    double t_prev = SS(t_soisno, 1);
    for (int k = 2; k <= CTSM_NLEVSOI; ++k) {
      const double tk = SS(t_soisno, k);
      if (t_prev > tfrz_ && tk <= tfrz_) { k_frz = k; break; }
      t_prev = tk;
    }
    const double frost_table = ZI(k_frz - 1);
    double zwt_perched = frost_table;
    const double t_frz = SS(t_soisno, k_frz);
    if (f.zwt[cc] < frost_table && t_frz <= tfrz_) {
    } else if (k_frz > 1) {
      int k_perch = 1;
      for (int k = k_frz; k >= 1; --k) {
        const double v = VOL(k);
        GG(h2osoi_vol, k) = v;
        if (v / GG(watsat, k) <= sat_lev) { k_perch = k; break; }
      }
      if (t_frz > tfrz_) k_perch = k_frz;
      if (k_frz > k_perch) {
        const double s1 = VOL(k_perch) / GG(watsat, k_perch);
        const double s2 = VOL(k_perch + 1) / GG(watsat, k_perch + 1);
        if (s1 > s2) {
          zwt_perched = ZI(k_perch - 1);
        } else {
          const double m = (SS(z, k_perch + 1) - SS(z, k_perch)) / (s2 - s1);
          const double b = SS(z, k_perch + 1) - m * s2;
          zwt_perched = fmax(0.0, m * sat_lev + b);
        }
      }
    }
    f.frost_table[cc] = frost_table;
    f.zwt_perched[cc] = zwt_perched;
  }
  // ThetaBasedWaterTable :1974-2021
  {
    const int nb = f.nbedrock[cc];
    double zwt;
    int k_zwt = nb, sat_flag = 1;
    for (int k = nb; k >= 1; --k) {
      const double v = VOL(k);
      GG(h2osoi_vol, k) = v;
      if (v / GG(watsat, k) <= sat_lev) { k_zwt = k; sat_flag = 0; break; }
    }
    if (sat_flag == 1) k_zwt = 1;
    if (k_zwt == 1) {
      zwt = ZI(1);
    } else if (k_zwt < nb) {
      const double s1 = VOL(k_zwt) / GG(watsat, k_zwt);
      const double s2 = VOL(k_zwt + 1) / GG(watsat, k_zwt + 1);
      const double m = (SS(z, k_zwt + 1) - SS(z, k_zwt)) / (s2 - s1);
      const double b = SS(z, k_zwt + 1) - m * s2;
      zwt = fmax(0.0, m * sat_lev + b);
    } else {
      zwt = ZI(nb);
    }
    f.zwt[cc] = zwt;
  }
  // RenewCondensation :2612-2674
  if (f.snl[cc] + 1 >= 1) {
    const double w = 1.0 - f.frac_h2osfc[cc];
    SS(h2osoi_liq, 1) = SS(h2osoi_liq, 1) + w * f.qflx_liqdew_to_top_layer[cc] * dtime;
    double ice = SS(h2osoi_ice, 1) + w * f.qflx_soliddew_to_top_layer[cc] * dtime;
    const double before = ice;
    ice = ice - w * f.qflx_solidevap_from_top_layer[cc] * dtime;
    if (fabs(ice) < 1.e-12 * fabs(before)) ice = 0.0;
    SS(h2osoi_ice, 1) = ice;
    if (ice < 0.0) report_failure(ds, c1, CTSM_ERR_SNOW_NEGATIVE, 2);
  }
}

// the non-lake loops of the tail, HydrologyNoDrainageMod.F90:432-570: one thread per non-lake column
__global__ void __launch_bounds__(128)
hydrodiag_nolake_kernel(HydroDiagDev f, int begc0, int ldc_, int numf, const int32_t* __restrict__ filterc, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int c1 = filterc[fc];
  const int cc = c1 - begc0;
  const size_t ldc = (size_t)ldc_;
  const int lt = f.lun_itype[cc];
  if (is_urban(lt)) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
  const int snl = f.snl[cc];
  // (snow totals: zero here; the snow-filter kernel, launched after this one, fills them on snow columns)
  f.snowice[cc] = 0.0; f.snowliq[cc] = 0.0; f.t_sno_mul_mss[cc] = 0.0;
  double t10 = 0.0, t17 = 0.0;
  const double t1 = SS(t_soisno, 1);
  double zi_above = ZI(0);
  for (int j = 1; j <= CTSM_NLEVSOI; ++j) {
    const double zij = ZI(j), tj = (j == 1) ? t1 : SS(t_soisno, j), dzj = SS(dz, j);
    if (zij <= 0.17) {
      t17 = t17 + tj * dzj * 1.0;
    } else if (zij > 0.17 && zi_above < 0.17) {
      const double fracl = (0.17 - zi_above) / dzj;
      t17 = t17 + tj * dzj * fracl;
    }
    if (zij <= 0.1) {
      t10 = t10 + tj * dzj * 1.0;
    } else if (zij > 0.1 && zi_above < 0.1) {
      const double fracl = (0.1 - zi_above) / dzj;
      t10 = t10 + tj * dzj * fracl;
    }
    zi_above = zij;
  }
  f.tsl[cc] = t1;
  const double fh = f.frac_h2osfc[cc], th = f.t_h2osfc[cc];
  const double t_top = SS(t_soisno, snl + 1);
  if (snl < 0) {
    const double fse = f.frac_sno_eff[cc];
    f.t_grnd[cc] = fse * t_top + (1.0 - fse - fh) * t1 + fh * th;
  } else {
    f.t_grnd[cc] = (1.0 - fh) * t1 + fh * th;
  }
  f.t_soi10cm[cc] = t10 / 0.1;
  f.t_soi17cm[cc] = t17 / 0.17;
  if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP) f.t_grnd_r[cc] = t_top;
  for (int j = 1; j <= CTSM_NLEVGRND; ++j) GG(h2osoi_vol, j) = VOL(j);
}

// the snow / no-snow loops :420-427, :438-447, :463-472, :738-754: threads [0, num_snowc) snow columns, the rest no-snow columns
__global__ void __launch_bounds__(128)
hydrodiag_snow_kernel(HydroDiagDev f, double dtime, int begc0, int ldc_, int num_snowc, const int32_t* __restrict__ filter_snowc,
                      int num_nosnowc, const int32_t* __restrict__ filter_nosnowc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_snowc + num_nosnowc) return;
  const size_t ldc = (size_t)ldc_;
  if (t >= num_snowc) {
    const int cc = filter_nosnowc[t - num_snowc] - begc0;
    f.snow_persistence[cc] = 0.0;
    f.h2osno_top[cc] = 0.0;
    for (int j = SNO_LO; j <= 0; ++j) SS(snw_rds, j) = 0.0;
    f.snot_top[cc] = 1.e36; f.dTdz_top[cc] = 1.e36; f.snw_rds_top[cc] = 1.e36; f.sno_liq_top[cc] = 1.e36;
    return;
  }
  const int cc = filter_snowc[t] - begc0;
  const int snl = f.snl[cc];
  f.snow_persistence[cc] = f.snow_persistence[cc] + dtime;
  double snowice = 0.0, snowliq = 0.0, tm = 0.0, top = 0.0;
  for (int j = snl + 1; j <= 0; ++j) {
    const double ice = SS(h2osoi_ice, j), liq = SS(h2osoi_liq, j);
    snowice = snowice + ice;
    snowliq = snowliq + liq;
    tm = tm + ice * SS(t_soisno, j);
    tm = tm + liq * 273.15;
    if (j == snl + 1) top = ice + liq;
  }
  f.snowice[cc] = snowice; f.snowliq[cc] = snowliq; f.t_sno_mul_mss[cc] = tm; f.h2osno_top[cc] = top;
}

// snowdp over the call bounds :452
__global__ void __launch_bounds__(256) hydrodiag_snowdp_kernel(HydroDiagDev f, int begc0, int begc, int endc) {
  const int c1 = begc + blockIdx.x * blockDim.x + threadIdx.x;
  if (c1 > endc) return;
  f.snowdp[c1 - begc0] = f.snow_depth[c1 - begc0] * f.frac_sno_eff[c1 - begc0];
}

// the hydrology-column loops :598-733: soilpsi, smp_l, wf, wf2.  h2osoi_vol is recomputed from its definition (the same expression
// the non-lake kernel stores), so this kernel does not depend on that one.
__global__ void __launch_bounds__(128)
hydrodiag_soil_kernel(HydroDiagDev f, int begc0, int ldc_, int numf, const int32_t* __restrict__ filterc) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int cc = filterc[fc] - begc0;
  const size_t ldc = (size_t)ldc_;
  const double smpmin = f.smpmin[cc];
  // wf sums the levels above 0.05 m, wf2 CONTINUES the same accumulators over the levels above 0.17 m (the reference does not
  // reset rwat / swat / rz between the two, HydrologyNoDrainageMod.F90:641-733): first sweep = everything per level + wf's sums,
  // second sweep = wf2's additions in level order on top of wf's totals
  double rwat5 = 0.0, swat5 = 0.0, rz5 = 0.0;
  double vol1 = 0.0, watdry1 = 0.0, watsat1 = 0.0;
  for (int j = 1; j <= CTSM_NLEVGRND; ++j) {
    const double liq = SS(h2osoi_liq, j), dzj = SS(dz, j), watsat = GG(watsat, j), sucsat = GG(sucsat, j), bsw = GG(bsw, j);
    const double vol = liq / (dzj * denh2o) + SS(h2osoi_ice, j) / (dzj * denice);
    if (liq > 0.0) {
      const double vwc = liq / (dzj * denh2o);
      const double fsattmp = fmax(vwc / watsat, 0.001);
      const double psi = sucsat * (-9.8e-6) * pow(fsattmp, -bsw);
      GG(soilpsi, j) = fmin(fmax(psi, -15.0), 0.0);
    } else {
      GG(soilpsi, j) = -15.0;
    }
    double s_node = fmax(vol / watsat, 0.01);
    s_node = fmin(1.0, s_node);
    GG(smp_l, j) = fmax(smpmin, -sucsat * pow(s_node, -bsw));
    const double zb = SS(z, j) + 0.5 * dzj;
    if (zb <= 0.05 || j == 1) {
      const double watdry = watsat * pow(316230.0 / sucsat, -1.0 / bsw);
      if (j == 1) { vol1 = vol; watdry1 = watdry; watsat1 = watsat; }
      if (zb <= 0.05) { rwat5 = rwat5 + (vol - watdry) * dzj; swat5 = swat5 + (watsat - watdry) * dzj; rz5 = rz5 + dzj; }
    }
  }
  {
    double tsw, stsw;
    if (rz5 != 0.0) { tsw = rwat5 / rz5; stsw = swat5 / rz5; }
    else { tsw = vol1 - watdry1; stsw = watsat1 - watdry1; }
    f.wf[cc] = tsw / stsw;
  }
  {
    double rw = rwat5, sw = swat5, rzz = rz5;
    for (int j = 1; j <= CTSM_NLEVGRND; ++j) {
      const double dzj = SS(dz, j);
      if (SS(z, j) + 0.5 * dzj <= 0.17) {
        const double watsat = GG(watsat, j);
        const double watdry = watsat * pow(316230.0 / GG(sucsat, j), -1.0 / GG(bsw, j));
        const double vol = SS(h2osoi_liq, j) / (dzj * denh2o) + SS(h2osoi_ice, j) / (dzj * denice);
        rw = rw + (vol - watdry) * dzj;
        sw = sw + (watsat - watdry) * dzj;
        rzz = rzz + dzj;
      }
    }
    double tsw, stsw;
    if (rzz != 0.0) { tsw = rw / rzz; stsw = sw / rzz; }
    else { tsw = vol1 - watdry1; stsw = watsat1 - watdry1; }
    f.wf2[cc] = tsw / stsw;
  }
}
#undef SS
#undef GG
#undef ZI
#undef VOL
}  // namespace

extern "C" int ctsm_b200_water_table(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_hydrologyc, const int32_t* filter_hydrologyc,
                                     int num_urbanc, const int32_t* filter_urbanc, const ctsm_watertable_fields_t* hf, int mem,
                                     ctsm_status_t* st) {
  (void)filter_urbanc;
  if (!ctx || !bounds || !hf || num_hydrologyc < 0 || (num_hydrologyc > 0 && !filter_hydrologyc)) return CTSM_ERR_BAD_ARG;
  if (num_urbanc != 0) return CTSM_ERR_URBAN;
  CUDA_TRY(cudaSetDevice(ctx->device));
  WaterTableDev d;
  const int32_t* dfh = filter_hydrologyc;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_WATERTABLE
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_WATERTABLE
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_hydrologyc, num_hydrologyc, &dfh);
    if (rc) return rc;
  }
  if (num_hydrologyc > 0) {
    water_table_kernel<<<grid_for(num_hydrologyc, 128), 128, 0, ctx->stream>>>(d, ctx->prm.dtime, hf->alloc.begc,
                                                                               hf->alloc.endc - hf->alloc.begc + 1, num_hydrologyc, dfh,
                                                                               ctx->d_status);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

extern "C" int ctsm_b200_hydrology_diagnostics(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec,
                                               const int32_t* filter_nolakec, int num_snowc, const int32_t* filter_snowc,
                                               int num_nosnowc, const int32_t* filter_nosnowc, int num_hydrologyc,
                                               const int32_t* filter_hydrologyc, int num_urbanc, const int32_t* filter_urbanc,
                                               const ctsm_hydrodiag_fields_t* hf, int mem, ctsm_status_t* st) {
  (void)filter_urbanc;
  if (!ctx || !bounds || !hf || num_nolakec < 0 || num_snowc < 0 || num_nosnowc < 0 || num_hydrologyc < 0 ||
      (num_nolakec > 0 && !filter_nolakec) || (num_snowc > 0 && !filter_snowc) || (num_nosnowc > 0 && !filter_nosnowc) ||
      (num_hydrologyc > 0 && !filter_hydrologyc))
    return CTSM_ERR_BAD_ARG;
  if (num_urbanc != 0) return CTSM_ERR_URBAN;
  CUDA_TRY(cudaSetDevice(ctx->device));
  HydroDiagDev d;
  const int32_t *dfn = filter_nolakec, *dfs = filter_snowc, *dfns = filter_nosnowc, *dfh = filter_hydrologyc;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_HYDRODIAG
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_HYDRODIAG
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    // four filters: the two filter arenas plus the two scratch arenas, which nothing else in this call uses
    rc = stage_filter(ctx, ctx->arena_filter0, filter_nolakec, num_nolakec, &dfn);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter1, filter_snowc, num_snowc, &dfs);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_ints, filter_nosnowc, num_nosnowc, &dfns);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_scratch, filter_hydrologyc, num_hydrologyc, &dfh);
    if (rc) return rc;
  }
  const int begc0 = hf->alloc.begc, ldc = hf->alloc.endc - hf->alloc.begc + 1;
  const int nb = bounds->endc - bounds->begc + 1;
  if (num_nolakec > 0) {
    hydrodiag_nolake_kernel<<<grid_for(num_nolakec, 128), 128, 0, ctx->stream>>>(d, begc0, ldc, num_nolakec, dfn, ctx->d_status);
    ctx->launches++;
  }
  if (nb > 0) {
    hydrodiag_snowdp_kernel<<<grid_for(nb, 256), 256, 0, ctx->stream>>>(d, begc0, bounds->begc, bounds->endc);
    ctx->launches++;
  }
  if (num_snowc + num_nosnowc > 0) {
    hydrodiag_snow_kernel<<<grid_for(num_snowc + num_nosnowc, 128), 128, 0, ctx->stream>>>(d, ctx->prm.dtime, begc0, ldc, num_snowc, dfs,
                                                                                           num_nosnowc, dfns);
    ctx->launches++;
  }
  if (num_hydrologyc > 0) {
    hydrodiag_soil_kernel<<<grid_for(num_hydrologyc, 128), 128, 0, ctx->stream>>>(d, begc0, ldc, num_hydrologyc, dfh);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}
