// soilfluxes.cu — SoilFluxes on B200.
//
// Reference: src/biogeophys/SoilFluxesMod.F90:37-521 (call site clm_driver.F90:921, between SoilTemperature and
// HydrologyNoDrainage) and p2c_1d_filter, src/main/subgridAveMod.F90:292-320.  Non-urban landunits.
//
// B200 mapping.  The reference makes eight passes over filter_nolakep with clump-sized scratch (tinc, t_grnd0,
// eflx_lwrad_del) and a level-outer / patch-inner sweep for the soil energy-balance error.  Here:
//   soilfluxes_patch_kernel   one thread per filter patch carries the patch through all of it in registers; the two
//                             column scalars the reference keeps in scratch (t_grnd0, tinc) are recomputed per patch
//                             from the same operands (identical values, 15 patches share a column's cache lines); the
//                             37-level error sum runs level-ascending per patch, i.e. in the reference's order;
//   soilfluxes_p2c_kernel     one thread per filter column: errsoi_col = sum over the column's active patches in
//                             ascending patch order (the reference's order; patches of a column are contiguous).
// Roofline: HBM.  ~45 patch fields of 8 B once each plus the column's level arrays (read once per column when L2
// serves the other 14 patches): DESIGN.md section 4.
#include "common.cuh"

struct Patch2ColDev {    // device-side view of ctsm_patch2col_fields_t
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_PATCH2COL
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_PATCH2COL
#undef CTSM_F
};

struct SoilFluxesDev {   // device-side view of ctsm_soilfluxes_fields_t
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_SOILFLUXES
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SOILFLUXES
#undef CTSM_F
};

namespace {
constexpr double hvap = 2.501e6, tfrz = 273.15, sb = 5.67e-8;
__device__ __forceinline__ double pow4(double t) { const double t2 = t * t; return t2 * t2; }
__device__ __forceinline__ double pow3(double t) { return (t * t) * t; }

__global__ void __launch_bounds__(256)
soilfluxes_patch_kernel(SoilFluxesDev f, double dtime, int begc0, int ldc_, int begp0, int nump,
                        const int32_t* __restrict__ filterp, DevStatus* ds) {
  const int fp = blockIdx.x * blockDim.x + threadIdx.x;
  if (fp >= nump) return;
  const int p1 = filterp[fp];
  const int pp = p1 - begp0;
  const int c1 = f.column[pp];
  const int cc = c1 - begc0;
  const size_t ldc = (size_t)ldc_;
#define C2(name, j) f.name[(size_t)((j) - SNOSOI_LO) * ldc + cc]
  const int lt = f.lun_itype[cc];
  if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
  const bool soilcrop = (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP);
  const int snl = f.snl[cc];
  const int j = snl + 1;
  const double frac_sno_eff = f.frac_sno_eff[cc], frac_h2osfc = f.frac_h2osfc[cc];
  const double tss_top = C2(t_ssbef, snl + 1), tss_1 = C2(t_ssbef, 1), t_h2osfc_bef = f.t_h2osfc_bef[cc];
  const double t_grnd = f.t_grnd[cc], htvp = f.htvp[cc], emg = f.emg[cc], forc_lwrad = f.forc_lwrad[cc];
  // :167-184
  double t_grnd0;
  if (snl < 0) t_grnd0 = frac_sno_eff * tss_top + (1 - frac_sno_eff - frac_h2osfc) * tss_1 + frac_h2osfc * t_h2osfc_bef;
  else t_grnd0 = (1 - frac_h2osfc) * tss_1 + frac_h2osfc * t_h2osfc_bef;
  const double tinc = t_grnd - t_grnd0;
  // :188-206 correct the fluxes to the present soil temperature
  const double cgrnds = f.cgrnds[pp], cgrndl = f.cgrndl[pp];
  double eflx_sh_grnd = f.eflx_sh_grnd[pp] + tinc * cgrnds;
  double qflx_evap_soi = f.qflx_evap_soi[pp] + tinc * cgrndl;
  double qflx_ev_snow = f.qflx_ev_snow[pp] + tinc * cgrndl;
  f.qflx_ev_soil[pp] = f.qflx_ev_soil[pp] + tinc * cgrndl;
  f.qflx_ev_h2osfc[pp] = f.qflx_ev_h2osfc[pp] + tinc * cgrndl;
  // :209-274 partition the evaporation of the top layer
  const double liq = C2(h2osoi_liq, j), ice = C2(h2osoi_ice, j);
  double liqevap = 0.0, solidevap = 0.0, soliddew = 0.0, liqdew = 0.0;
  if (qflx_ev_snow >= 0.0) {
    if ((liq + ice) > 0.0) liqevap = fmax(qflx_ev_snow * (liq / (liq + ice)), 0.0);
    else liqevap = 0.0;
    solidevap = qflx_ev_snow - liqevap;
  } else {
    if (t_grnd < tfrz) soliddew = fabs(qflx_ev_snow);
    else liqdew = fabs(qflx_ev_snow);
  }
  // :277-331 limit evaporation to the available moisture
  if (j < 1) {
    const double evaporation_limit = (ice + liq) / (frac_sno_eff * dtime);
    if (qflx_ev_snow > evaporation_limit) {
      const double evaporation_demand = qflx_ev_snow;
      qflx_ev_snow = evaporation_limit;
      qflx_evap_soi = qflx_evap_soi - frac_sno_eff * (evaporation_demand - evaporation_limit);
      liqevap = fmax(liq / (frac_sno_eff * dtime), 0.0);
      solidevap = fmax(ice / (frac_sno_eff * dtime), 0.0);
      eflx_sh_grnd = eflx_sh_grnd + frac_sno_eff * (evaporation_demand - evaporation_limit) * htvp;
    }
  }
  if (j == 1 && frac_h2osfc < 1.0) {
    const double evaporation_limit = ice / (dtime * (1.0 - frac_h2osfc));
    if (solidevap >= evaporation_limit) {
      const double evaporation_demand = solidevap;
      solidevap = evaporation_limit;
      liqevap = liqevap + (evaporation_demand - evaporation_limit);
    }
  }
  f.eflx_sh_grnd[pp] = eflx_sh_grnd;
  f.qflx_evap_soi[pp] = qflx_evap_soi;
  f.qflx_ev_snow[pp] = qflx_ev_snow;
  f.qflx_liqevap_from_top_layer_patch[pp] = liqevap;
  f.qflx_solidevap_from_top_layer_patch[pp] = solidevap;
  f.qflx_soliddew_to_top_layer_patch[pp] = soliddew;
  f.qflx_liqdew_to_top_layer_patch[pp] = liqdew;
  // :338-400 ground heat flux and totals
  const int fv = f.frac_veg_nosno[pp];
  const double lw_grnd = (frac_sno_eff * pow4(tss_top) + (1.0 - frac_sno_eff - frac_h2osfc) * pow4(tss_1)
                          + frac_h2osfc * pow4(t_h2osfc_bef));
  const double dlrad = f.dlrad[pp];
  const double eflx_soil_grnd = ((1.0 - frac_sno_eff) * f.sabg_soil[pp] + frac_sno_eff * f.sabg_snow[pp]) + dlrad
                                + (1 - fv) * emg * forc_lwrad - emg * sb * lw_grnd - emg * sb * pow3(t_grnd0) * (4.0 * tinc)
                                - (eflx_sh_grnd + qflx_evap_soi * htvp);
  f.eflx_soil_grnd[pp] = eflx_soil_grnd;
  if (soilcrop) f.eflx_soil_grnd_r[pp] = eflx_soil_grnd;
  const double qflx_evap_veg = f.qflx_evap_veg[pp], qflx_tran_veg = f.qflx_tran_veg[pp];
  double eflx_sh_tot = f.eflx_sh_veg[pp] + eflx_sh_grnd;
  eflx_sh_tot = eflx_sh_tot + f.eflx_sh_stem[pp];
  const double eflx_lh_tot = hvap * qflx_evap_veg + htvp * qflx_evap_soi;
  f.eflx_sh_tot[pp] = eflx_sh_tot;
  f.qflx_evap_tot_patch[pp] = qflx_evap_veg + qflx_evap_soi;
  f.eflx_lh_tot[pp] = eflx_lh_tot;
  if (soilcrop) { f.eflx_lh_tot_r[pp] = eflx_lh_tot; f.eflx_sh_tot_r[pp] = eflx_sh_tot; }
  f.qflx_evap_can[pp] = qflx_evap_veg - qflx_tran_veg;
  f.eflx_lh_vege[pp] = (qflx_evap_veg - qflx_tran_veg) * hvap;
  f.eflx_lh_vegt[pp] = qflx_tran_veg * hvap;
  f.eflx_lh_grnd[pp] = qflx_evap_soi * htvp;
  // :406-436 soil energy balance error (level-ascending per patch = the reference's level-outer order)
  double errsoi = eflx_soil_grnd - f.xmf[cc] - f.xmf_h2osfc[cc]
                  - frac_h2osfc * (f.t_h2osfc[cc] - t_h2osfc_bef) * (f.c_h2osfc[cc] / dtime);
  errsoi = errsoi + f.eflx_h2osfc_to_snow[cc];
  for (int jj = snl + 1; jj < 1; ++jj) errsoi = errsoi - frac_sno_eff * (C2(t_soisno, jj) - C2(t_ssbef, jj)) / C2(fact, jj);
  for (int jj = 1; jj <= NLEVGRND; ++jj) errsoi = errsoi - (C2(t_soisno, jj) - C2(t_ssbef, jj)) / C2(fact, jj);
  f.errsoi_patch[pp] = errsoi;
  // :463-500 outgoing longwave and bare-ground skin temperature
  const double eflx_lwrad_out = f.ulrad[pp] + (1 - fv) * (1. - emg) * forc_lwrad + (1 - fv) * emg * sb * lw_grnd
                                + 4.0 * emg * sb * pow3(t_grnd0) * tinc;
  f.eflx_lwrad_out[pp] = eflx_lwrad_out;
  if (fv == 0) f.t_skin[pp] = sqrt(sqrt(lw_grnd));
  f.eflx_lwrad_net[pp] = eflx_lwrad_out - forc_lwrad;
  if (soilcrop) { f.eflx_lwrad_net_r[pp] = eflx_lwrad_out - forc_lwrad; f.eflx_lwrad_out_r[pp] = eflx_lwrad_out; }
#undef C2
}

// p2c_1d_filter (subgridAveMod.F90:312-318)
__global__ void __launch_bounds__(128)
soilfluxes_p2c_kernel(SoilFluxesDev f, int begc0, int begp0, int numc, const int32_t* __restrict__ filterc) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numc) return;
  const int cc = filterc[fc] - begc0;
  double s = 0.0;
  const int pi = f.patchi[cc], pf = f.patchf[cc];
  for (int p1 = pi; p1 <= pf; ++p1) {
    const int pp = p1 - begp0;
    if (f.patch_active[pp]) s = s + f.errsoi_patch[pp] * f.wtcol[pp];
  }
  f.errsoi_col[cc] = s;
}

// clm_drv_patch2col (clm_driver.F90:1655-1739): the reference calls p2c eleven times, i.e. walks every column's patch
// list eleven times; here one thread per column walks it once and forms all averages (each a separate accumulator in
// ascending patch order = the reference's order).  ALLC = false: the ten averages over filter_nolakec; ALLC = true:
// qflx_evap_soi over filter_allc (:1722-1724; it overrides the non-lake average, which is the same number there).
template <bool ALLC>
__global__ void __launch_bounds__(128)
patch2col_kernel(Patch2ColDev f, int begc0, int begp0, int numc, const int32_t* __restrict__ filterc) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numc) return;
  const int cc = filterc[fc] - begc0;
  const int pi = f.patchi[cc], pf = f.patchf[cc];
  double a[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) a[k] = 0.0;
  for (int p1 = pi; p1 <= pf; ++p1) {
    const int pp = p1 - begp0;
    if (!f.patch_active[pp]) continue;
    const double wt = f.wtcol[pp];
    a[3] = a[3] + f.qflx_evap_soi[pp] * wt;
    if (!ALLC) {
      a[0] = a[0] + f.qflx_ev_snow[pp] * wt;
      a[1] = a[1] + f.qflx_ev_soil[pp] * wt;
      a[2] = a[2] + f.qflx_ev_h2osfc[pp] * wt;
      a[4] = a[4] + f.qflx_evap_tot_patch[pp] * wt;
      a[5] = a[5] + f.qflx_tran_veg[pp] * wt;
      a[6] = a[6] + f.qflx_liqevap_from_top_layer_patch[pp] * wt;
      a[7] = a[7] + f.qflx_liqdew_to_top_layer_patch[pp] * wt;
      a[8] = a[8] + f.qflx_solidevap_from_top_layer_patch[pp] * wt;
      a[9] = a[9] + f.qflx_soliddew_to_top_layer_patch[pp] * wt;
    }
  }
  f.qflx_evap_soi_col[cc] = a[3];
  if (!ALLC) {
    f.qflx_ev_snow_col[cc] = a[0]; f.qflx_ev_soil_col[cc] = a[1]; f.qflx_ev_h2osfc_col[cc] = a[2];
    f.qflx_evap_tot[cc] = a[4]; f.qflx_tran_veg_col[cc] = a[5]; f.qflx_liqevap_from_top_layer[cc] = a[6];
    f.qflx_liqdew_to_top_layer[cc] = a[7]; f.qflx_solidevap_from_top_layer[cc] = a[8]; f.qflx_soliddew_to_top_layer[cc] = a[9];
  }
}

// clm_drv_patch2col with one WARP per column (default; CTSM_B200_SINK_WARP=0 selects the thread-per-column kernels above;
// bit-identical, tests/test_gpu_soilfluxes.py).  Lane i owns patch patchi + i, so every patch field is read in contiguous runs
// (a thread per column strides by the column's patch count and fetched 9.4 x the algorithmic bytes from DRAM); the weighted
// terms go through shared memory to lane k, which adds average k's terms in ascending patch order (the reference's order).
constexpr int P2C_WARPS = 8;
constexpr int P2C_LD = 33;
template <bool ALLC>
__global__ void __launch_bounds__(P2C_WARPS * 32)
patch2col_warp_kernel(Patch2ColDev f, int begc0, int begp0, int numc, const int32_t* __restrict__ filterc) {
  __shared__ double s_t[P2C_WARPS][10 * P2C_LD];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int fc = blockIdx.x * P2C_WARPS + w;
  if (fc >= numc) return;                                     // warp-uniform
  const int cc = filterc[fc] - begc0;
  const int pi = f.patchi[cc], np = f.patchf[cc] - pi + 1;
  double acc = 0.0;                                           // lane k accumulates average k
  for (int base = 0; base < np; base += 32) {
    const int pp = pi + base + lane - begp0;
    const bool q = (base + lane < np) && f.patch_active[pp];
    if (q) {
      const double wt = f.wtcol[pp];
      s_t[w][3 * P2C_LD + lane] = f.qflx_evap_soi[pp] * wt;
      if (!ALLC) {
        s_t[w][0 * P2C_LD + lane] = f.qflx_ev_snow[pp] * wt;
        s_t[w][1 * P2C_LD + lane] = f.qflx_ev_soil[pp] * wt;
        s_t[w][2 * P2C_LD + lane] = f.qflx_ev_h2osfc[pp] * wt;
        s_t[w][4 * P2C_LD + lane] = f.qflx_evap_tot_patch[pp] * wt;
        s_t[w][5 * P2C_LD + lane] = f.qflx_tran_veg[pp] * wt;
        s_t[w][6 * P2C_LD + lane] = f.qflx_liqevap_from_top_layer_patch[pp] * wt;
        s_t[w][7 * P2C_LD + lane] = f.qflx_liqdew_to_top_layer_patch[pp] * wt;
        s_t[w][8 * P2C_LD + lane] = f.qflx_solidevap_from_top_layer_patch[pp] * wt;
        s_t[w][9 * P2C_LD + lane] = f.qflx_soliddew_to_top_layer_patch[pp] * wt;
      }
    }
    unsigned m = __ballot_sync(FULL, q);
    __syncwarp();
    if (lane < 10 && (!ALLC || lane == 3)) {
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        acc = acc + s_t[w][lane * P2C_LD + i];
      }
    }
    __syncwarp();
  }
  if (ALLC) {
    if (lane == 3) f.qflx_evap_soi_col[cc] = acc;
    return;
  }
  double* const out[10] = {f.qflx_ev_snow_col, f.qflx_ev_soil_col, f.qflx_ev_h2osfc_col, f.qflx_evap_soi_col, f.qflx_evap_tot,
                           f.qflx_tran_veg_col, f.qflx_liqevap_from_top_layer, f.qflx_liqdew_to_top_layer,
                           f.qflx_solidevap_from_top_layer, f.qflx_soliddew_to_top_layer};
#pragma unroll
  for (int k = 0; k < 10; ++k)
    if (lane == k) out[k][cc] = acc;
}
}  // namespace

extern "C" int ctsm_b200_patch2col(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_allc, const int32_t* filter_allc,
                                   int num_nolakec, const int32_t* filter_nolakec, const ctsm_patch2col_fields_t* hf, int mem,
                                   ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_allc < 0 || num_nolakec < 0 || (num_allc > 0 && !filter_allc) ||
      (num_nolakec > 0 && !filter_nolakec))
    return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  Patch2ColDev d;
  const int32_t *dfa = filter_allc, *dfc = filter_nolakec;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_PATCH2COL
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_PATCH2COL
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_allc, num_allc, &dfa);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter1, filter_nolakec, num_nolakec, &dfc);
    if (rc) return rc;
  }
  const bool warp = ctx->tune.sink_warp != 0;
  if (num_nolakec > 0) {
    if (warp)
      patch2col_warp_kernel<false><<<(num_nolakec + P2C_WARPS - 1) / P2C_WARPS, P2C_WARPS * 32, 0, ctx->stream>>>(
          d, hf->alloc.begc, hf->alloc.begp, num_nolakec, dfc);
    else
      patch2col_kernel<false><<<grid_for(num_nolakec, 128), 128, 0, ctx->stream>>>(d, hf->alloc.begc, hf->alloc.begp, num_nolakec, dfc);
    ctx->launches++;
  }
  if (num_allc > 0) {
    if (warp)
      patch2col_warp_kernel<true><<<(num_allc + P2C_WARPS - 1) / P2C_WARPS, P2C_WARPS * 32, 0, ctx->stream>>>(
          d, hf->alloc.begc, hf->alloc.begp, num_allc, dfa);
    else
      patch2col_kernel<true><<<grid_for(num_allc, 128), 128, 0, ctx->stream>>>(d, hf->alloc.begc, hf->alloc.begp, num_allc, dfa);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

extern "C" int ctsm_b200_soilfluxes(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec,
                                    const int32_t* filter_nolakec, int num_nolakep, const int32_t* filter_nolakep,
                                    const ctsm_soilfluxes_fields_t* hf, int mem, ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_nolakec < 0 || num_nolakep < 0 || (num_nolakec > 0 && !filter_nolakec) ||
      (num_nolakep > 0 && !filter_nolakep))
    return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  SoilFluxesDev d;
  const int32_t *dfc = filter_nolakec, *dfp = filter_nolakep;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_SOILFLUXES
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SOILFLUXES
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
  const int ldc = hf->alloc.endc - hf->alloc.begc + 1;
  if (num_nolakep > 0) {
    soilfluxes_patch_kernel<<<grid_for(num_nolakep, 256), 256, 0, ctx->stream>>>(d, ctx->prm.dtime, hf->alloc.begc, ldc,
                                                                                 hf->alloc.begp, num_nolakep, dfp, ctx->d_status);
    ctx->launches++;
  }
  if (num_nolakec > 0) {
    soilfluxes_p2c_kernel<<<grid_for(num_nolakec, 128), 128, 0, ctx->stream>>>(d, hf->alloc.begc, hf->alloc.begp, num_nolakec, dfc);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}
