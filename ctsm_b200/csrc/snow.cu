// snow.cu — the snow routines of HydrologyNoDrainage on B200; src/biogeophys/SnowHydrologyMod.F90:
//   SnowWater :1015-1165 (UpdateState_TopLayerFluxes, BulkFlux_SnowPercolation, UpdateState_SnowPercolation,
//     CalcAndApplyAerosolFluxes + AerosolFluxes AerosolMod.F90:668-801, PostPercolation_AdjustLayerThicknesses,
//     BulkDiag_SnowWaterAccumulatedSnow, SumFlux_AddSnowPercolation)                       -> ctsm_b200_snow_water
//   SnowCompaction :1870-2080, CombineSnowLayers :2083-2507, DivideSnowLayers :2510-2895 (is_lake = .false.),
//     ZeroEmptySnowLayers :2898-2952  (HydrologyNoDrainageMod.F90:381-399)                  -> ctsm_b200_snow_layers
// (BuildSnowFilter :3975 is ctsm_b200_build_snow_filter, next to the other order-preserving filter split in canopy.cu.)
//
// B200 mapping.  A snow column never reads another column, and its pack has at most 12 layers: one thread per column.
//   snow_water_kernel     the reference's seven passes over the pack are one top-to-bottom sweep: the percolation flux of layer j
//                         needs the pre-update volumes of layers j and j+1 only, so the thread carries layer j+1 as lookahead in
//                         registers and every layer's fields are read once and written once (level-major arrays: the 32 columns
//                         of a warp read 32 consecutive doubles per level).  Threads past the snow filter do the two no-snow
//                         loops.  HBM-bound: 13 fields x active layers.
//   aerosol_dep_kernel    the 19 deposition diagnostics AerosolFluxes writes for every column of the call bounds
//   snow_layers_kernel    compaction, combination, subdivision and zeroing on a thread-private copy of the pack (13 fields x 12
//                         layers + soil layer 1, loaded and stored coalesced); the layer bookkeeping is the reference's, index by
//                         index, because the elements a shift leaves behind in aerosol / grain-radius arrays are part of the
//                         reference's state.  Divergent by nature (layer counts differ), small: a few per cent of SoilTemperature.
// Bulk water only; non-lake, non-urban columns.
#include "common.cuh"
#include <vector>

struct SnowWaterDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_SNOWWATER
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SNOWWATER
#undef CTSM_F
};
struct SnowLayersDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_SNOWLAYERS
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SNOWLAYERS
#undef CTSM_F
};

struct SnowCappingDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_SNOWCAPPING
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SNOWCAPPING
#undef CTSM_F
};

namespace {
using namespace cst;
constexpr int NS = CTSM_NLEVSNO;
constexpr int SLO = -CTSM_NLEVSNO + 1;
constexpr double rpi = 3.14159265358979323846, snw_rds_max = 1500.0;
constexpr double scvng_fct_mlt_ocphi = 0.20, scvng_fct_mlt_ocpho = 0.03;          // SnowHydrologyMod.F90:131-132

struct SnowGeo { int begc0, begg0, ldc, ldg; };
struct SnowWaterPrm { double dtime, wimp, ssi, sf, scv[8]; int use_aerosol; };

__device__ __forceinline__ bool is_urban(int lt) { return lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX; }

// AerosolFluxes, AerosolMod.F90:725-777: every column of the call bounds
__global__ void __launch_bounds__(256)
aerosol_dep_kernel(SnowWaterDev f, SnowGeo geo, int begc, int endc, int on) {
  const int c1 = begc + blockIdx.x * blockDim.x + threadIdx.x;
  if (c1 > endc) return;
  const int cc = c1 - geo.begc0;
  const int g = f.col_gridcell[cc] - geo.begg0;
  double a[14];
#pragma unroll
  for (int k = 0; k < 14; ++k) a[k] = on ? __ldg(&f.forc_aer[(size_t)k * geo.ldg + g]) : 0.0;
  f.flx_bc_dep_dry[cc] = a[0] + a[1];
  f.flx_bc_dep_wet[cc] = a[2];
  f.flx_bc_dep_phi[cc] = a[0] + a[2];
  f.flx_bc_dep_pho[cc] = a[1];
  f.flx_bc_dep[cc] = a[0] + a[1] + a[2];
  f.flx_oc_dep_dry[cc] = a[3] + a[4];
  f.flx_oc_dep_wet[cc] = a[5];
  f.flx_oc_dep_phi[cc] = a[3] + a[5];
  f.flx_oc_dep_pho[cc] = a[4];
  f.flx_oc_dep[cc] = a[3] + a[4] + a[5];
  f.flx_dst_dep_wet1[cc] = a[6];
  f.flx_dst_dep_dry1[cc] = a[7];
  f.flx_dst_dep_wet2[cc] = a[8];
  f.flx_dst_dep_dry2[cc] = a[9];
  f.flx_dst_dep_wet3[cc] = a[10];
  f.flx_dst_dep_dry3[cc] = a[11];
  f.flx_dst_dep_wet4[cc] = a[12];
  f.flx_dst_dep_dry4[cc] = a[13];
  f.flx_dst_dep[cc] = a[6] + a[7] + a[8] + a[9] + a[10] + a[11] + a[12] + a[13];
}

__global__ void __launch_bounds__(128)
snow_water_kernel(SnowWaterDev f, SnowGeo geo, SnowWaterPrm prm, int num_snowc, const int32_t* __restrict__ filter_snowc,
                  int num_nosnowc, const int32_t* __restrict__ filter_nosnowc, DevStatus* ds) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_snowc + num_nosnowc) return;
  const double dtime = prm.dtime;
  if (t >= num_snowc) {                                       // the two no-snow loops (:1805-1811, :1859-1865)
    const int cc = filter_nosnowc[t - num_snowc] - geo.begc0;
    if (f.h2osno_no_layers[cc] <= 0.0) { f.int_snow[cc] = 0.0; f.frac_sno[cc] = 0.0; f.snow_depth[cc] = 0.0; }
    const double melt = f.qflx_snomelt[cc];
    f.qflx_snow_drain[cc] = melt;
    f.qflx_rain_plus_snomelt[cc] = f.qflx_liq_grnd[cc] + melt;
    return;
  }
  const int c1 = filter_snowc[t];
  const int cc = c1 - geo.begc0;
  const size_t ldc = (size_t)geo.ldc;
  const int snl = f.snl[cc];
  const int top = snl + 1;
  const double fse = f.frac_sno_eff[cc];
  const double q_liq_grnd = f.qflx_liq_grnd[cc], q_sdew = f.qflx_soliddew_to_top_layer[cc], q_ldew = f.qflx_liqdew_to_top_layer[cc];
  double* const mss[8] = {f.mss_bcphi, f.mss_bcpho, f.mss_ocphi, f.mss_ocpho, f.mss_dst1, f.mss_dst2, f.mss_dst3, f.mss_dst4};
#define OFF(j) ((size_t)((j) - SLO) * ldc + cc)
  // UpdateState_TopLayerFluxes :1210-1287
  double ice = f.h2osoi_ice[OFF(top)], liq = f.h2osoi_liq[OFF(top)];
  {
    const double ice0 = ice, liq0 = liq;
    ice = ice + fse * (q_sdew - f.qflx_solidevap_from_top_layer[cc]) * dtime;
    liq = liq + fse * (q_liq_grnd + q_ldew - f.qflx_liqevap_from_top_layer[cc]) * dtime;
    if (fabs(ice) < 1.e-12 * fabs(ice0)) ice = 0.0;
    if (fabs(liq) < 1.e-12 * fabs(liq0)) liq = 0.0;
    if (ice < 0.0) { report_failure(ds, c1, CTSM_ERR_SNOW_NEGATIVE, 0); return; }
    if (liq < 0.0) { report_failure(ds, c1, CTSM_ERR_SNOW_NEGATIVE, 1); return; }
  }
  // deposition onto the top layer (AerosolMod.F90:785-798), applied after the inter-layer fluxes as in the reference
  double dep[8];
  {
    const int g = f.col_gridcell[cc] - geo.begg0;
    double a[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) a[k] = prm.use_aerosol ? __ldg(&f.forc_aer[(size_t)k * geo.ldg + g]) : 0.0;
    dep[0] = ((a[0] + a[2]) * dtime); dep[1] = (a[1] * dtime); dep[2] = ((a[3] + a[5]) * dtime); dep[3] = (a[4] * dtime);
    dep[4] = (a[7] + a[6]) * dtime; dep[5] = (a[9] + a[8]) * dtime; dep[6] = (a[11] + a[10]) * dtime; dep[7] = (a[13] + a[12]) * dtime;
  }
  // one sweep down the pack; (ice, liq, dzj) = layer j before percolation, (.._n) = layer j + 1
  double dzj = f.dz[OFF(top)];
  double vol_ice = fmin(1.0, ice / (dzj * fse * denice));
  double eff_por = 1.0 - vol_ice;
  double vol_liq = fmin(eff_por, liq / (dzj * fse * denh2o));
  double q_above = 0.0;                                       // qflx_snow_percolation(c, j-1)
  double qin[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double q = 0.0;
  for (int j = top; j <= 0; ++j) {
    double ice_n = 0.0, liq_n = 0.0, dz_n = 0.0, vol_ice_n = 0.0, eff_por_n = 0.0, vol_liq_n = 0.0;
    if (j <= -1) {
      ice_n = f.h2osoi_ice[OFF(j + 1)]; liq_n = f.h2osoi_liq[OFF(j + 1)]; dz_n = f.dz[OFF(j + 1)];
      vol_ice_n = fmin(1.0, ice_n / (dz_n * fse * denice));
      eff_por_n = 1.0 - vol_ice_n;
      vol_liq_n = fmin(eff_por_n, liq_n / (dz_n * fse * denh2o));
      if (eff_por < prm.wimp || eff_por_n < prm.wimp) {
        q = 0.0;
      } else {
        q = fmax(0.0, (vol_liq - prm.ssi * eff_por) * dzj * fse);
        q = fmin(q, (1.0 - vol_ice_n - vol_liq_n) * dz_n * fse);
      }
    } else {
      q = fmax(0.0, (vol_liq - prm.ssi * eff_por) * dzj * fse);
    }
    q = (q * 1000.0) / dtime;
    f.qflx_snow_percolation[OFF(j)] = q;
    // UpdateState_SnowPercolation :1486-1492
    if (j >= snl + 2) liq = liq + q_above * dtime;
    liq = liq - q * dtime;
    f.h2osoi_liq[OFF(j)] = liq;
    if (j == top) f.h2osoi_ice[OFF(j)] = ice;
    // CalcAndApplyAerosolFluxes :1577-1698
    double mss_liqice = liq + ice;
    if (mss_liqice < 1e-30) mss_liqice = 1e-30;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double m = mss[k][OFF(j)] + qin[k] * dtime;
      double qout = q * prm.sf * prm.scv[k] * (m / mss_liqice);
      if (qout * dtime > m) { qout = m / dtime; m = 0.0; }
      else m = m - qout * dtime;
      qin[k] = qout;
      if (j == top) m = m + dep[k];
      mss[k][OFF(j)] = m;
    }
    // PostPercolation_AdjustLayerThicknesses :1744
    f.dz[OFF(j)] = fmax(dzj, liq / denh2o + ice / denice);
    q_above = q;
    ice = ice_n; liq = liq_n; dzj = dz_n; vol_ice = vol_ice_n; eff_por = eff_por_n; vol_liq = vol_liq_n;
  }
  // BulkDiag_SnowWaterAccumulatedSnow :1796-1800, SumFlux_AddSnowPercolation :1853-1857 (q = the bottom layer's flux)
  f.int_snow[cc] = f.int_snow[cc] + fse * (q_sdew + q_ldew + q_liq_grnd) * dtime;
  f.qflx_snow_drain[cc] = f.qflx_snow_drain[cc] + q;
  f.qflx_rain_plus_snomelt[cc] = q + (1.0 - fse) * q_liq_grnd;
#undef OFF
}

// SnowCapping :3121-3247.  Two launches: the four capping fluxes are zeroed over filter_initc (InitFlux_SnowCapping), then one
// thread per snow column sums its pack (CalculateTotalH2osno), finds the excess (SnowCappingExcess) and, where there is one,
// removes it from the bottom layer (fluxes, state, thickness at constant ice density, aerosol masses).
struct SnowCapPrm { double dtime, h2osno_max, reset_snow_glc_ela; int reset_snow, reset_snow_glc, reset_active; };

__global__ void __launch_bounds__(256)
snow_capping_init_kernel(SnowCappingDev f, int begc0, int numf, const int32_t* __restrict__ filterc) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int cc = filterc[fc] - begc0;
  f.qflx_snwcp_ice[cc] = 0.0; f.qflx_snwcp_liq[cc] = 0.0; f.qflx_snwcp_discarded_ice[cc] = 0.0; f.qflx_snwcp_discarded_liq[cc] = 0.0;
}

__global__ void __launch_bounds__(128)
snow_capping_kernel(SnowCappingDev f, SnowCapPrm prm, int begc0, int ldc_, int numf, const int32_t* __restrict__ filterc, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int c1 = filterc[fc];
  const int cc = c1 - begc0;
  const size_t ldc = (size_t)ldc_;
#define OFF(j) ((size_t)((j) - SLO) * ldc + cc)
  const int snl = f.snl[cc], lt = f.lun_itype[cc];
  double h2osno = f.h2osno_no_layers[cc];
  double ice0 = 0.0, liq0 = 0.0;
  for (int j = snl + 1; j <= 0; ++j) {
    ice0 = f.h2osoi_ice[OFF(j)]; liq0 = f.h2osoi_liq[OFF(j)];
    h2osno = h2osno + ice0 + liq0;
  }
  double excess = 0.0;
  bool apply_runoff = false;
  if (h2osno > prm.h2osno_max) { excess = h2osno - prm.h2osno_max; apply_runoff = true; }
  if (prm.reset_active) {                                                          // :3457-3471, reset_snow_h2osno = 35 mm
    if ((lt != CTSM_ISTICE) && prm.reset_snow && (h2osno > 35.0)) { excess = h2osno - 35.0; apply_runoff = false; }
    else if ((lt == CTSM_ISTICE) && prm.reset_snow_glc && (h2osno > 35.0) && (f.topo[cc] <= prm.reset_snow_glc_ela)) {
      excess = h2osno - 35.0; apply_runoff = false;
    }
  }
  if (!(excess > 0.0)) return;
  // (ice0, liq0) = the bottom layer, the last one the loop read
  const double dz0 = f.dz[OFF(0)];
  const double rho_orig_bottom = ice0 / dz0;
  const double mss_snow_bottom_lyr = ice0 + liq0;
  const double mss_snwcp_tot = fmin(excess, mss_snow_bottom_lyr * (1.0 - 1.e-3));
  const double icefrac = ice0 / mss_snow_bottom_lyr;
  const double snwcp_flux_ice = mss_snwcp_tot / prm.dtime * icefrac;
  const double snwcp_flux_liq = mss_snwcp_tot / prm.dtime * (1.0 - icefrac);
  double q_ice = 0.0, q_liq = 0.0, qd_ice = 0.0, qd_liq = 0.0;
  if (apply_runoff) { q_ice = snwcp_flux_ice; q_liq = snwcp_flux_liq; f.qflx_snwcp_ice[cc] = q_ice; f.qflx_snwcp_liq[cc] = q_liq; }
  else { qd_ice = snwcp_flux_ice; qd_liq = snwcp_flux_liq; f.qflx_snwcp_discarded_ice[cc] = qd_ice; f.qflx_snwcp_discarded_liq[cc] = qd_liq; }
  const double frac_adjust = (mss_snow_bottom_lyr - mss_snwcp_tot) / mss_snow_bottom_lyr;
  const double ice1 = ice0 - (q_ice + qd_ice) * prm.dtime;
  const double liq1 = liq0 - (q_liq + qd_liq) * prm.dtime;
  f.h2osoi_ice[OFF(0)] = ice1; f.h2osoi_liq[OFF(0)] = liq1;
  if (ice1 < 0.0 || liq1 < 0.0) { report_failure(ds, c1, CTSM_ERR_SNOW_NEGATIVE, 3); return; }
  if (rho_orig_bottom > 1.0) f.dz[OFF(0)] = ice1 / rho_orig_bottom;
  double* const mss[8] = {f.mss_bcphi, f.mss_bcpho, f.mss_ocphi, f.mss_ocpho, f.mss_dst1, f.mss_dst2, f.mss_dst3, f.mss_dst4};
#pragma unroll
  for (int k = 0; k < 8; ++k) mss[k][OFF(0)] = mss[k][OFF(0)] * frac_adjust;
#undef OFF
}

struct SnowLayersPrm {
  double dtime, int_snow_max, upplim, Tfactor, eta0_anderson, eta0_vionnet, ceta, drift_gs, tau_ref, rho_max, snw_rds_min;
  double dzmin[NS], dzmax_l[NS], dzmax_u[NS];
  int method, wind, subgrid;
};

// Combo :3902-3946
__device__ __forceinline__ void combo(double& dz, double& wliq, double& wice, double& t, double dz2, double wliq2, double wice2,
                                      double t2) {
  const double dzc = dz + dz2;
  const double wicec = (wice + wice2);
  const double wliqc = (wliq + wliq2);
  const double h = (cpice * wice + cpliq * wliq) * (t - tfrz) + hfus * wliq;
  const double h2 = (cpice * wice2 + cpliq * wliq2) * (t2 - tfrz) + hfus * wliq2;
  const double hc = h + h2;
  const double tc = tfrz + (hc - hfus * wliqc) / (cpice * wicec + cpliq * wliqc);
  dz = dzc; wice = wicec; wliq = wliqc; t = tc;
}

#ifndef SNOW_LAYERS_BLOCKS
#define SNOW_LAYERS_BLOCKS 4      // resident blocks per SM: 128 registers, 16 warps per SM (2 blocks: 252 registers)
#endif
__global__ void __launch_bounds__(128, SNOW_LAYERS_BLOCKS)
snow_layers_kernel(SnowLayersDev f, SnowGeo geo, SnowLayersPrm prm, int num_snowc, const int32_t* __restrict__ filter_snowc,
                   DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= num_snowc) return;
  const int c1 = filter_snowc[fc];
  const int cc = c1 - geo.begc0;
  const size_t ldc = (size_t)geo.ldc;
  const int lt = f.lun_itype[cc];
  if (is_urban(lt)) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
  if (lt == CTSM_ISTDLAK) { report_failure(ds, c1, CTSM_ERR_BAD_ARG, 0); return; }
  const bool soil = (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP);
  const double dtime = prm.dtime;
  // thread-private pack: index i = level + 11 (levels -11 .. 0); ice / liq carry soil layer 1 at index 12
  double dz[NS], tk[NS], rds[NS], ice[NS + 1], liq[NS + 1], ms[8][NS];
  double* const mss[8] = {f.mss_bcphi, f.mss_bcpho, f.mss_ocphi, f.mss_ocpho, f.mss_dst1, f.mss_dst2, f.mss_dst3, f.mss_dst4};
#define OFF(j) ((size_t)((j) - SLO) * ldc + cc)
#define IX(j) ((j) - SLO)
  for (int j = SLO; j <= 0; ++j) {
    dz[IX(j)] = f.dz[OFF(j)]; tk[IX(j)] = f.t_soisno[OFF(j)]; rds[IX(j)] = f.snw_rds[OFF(j)];
    ice[IX(j)] = f.h2osoi_ice[OFF(j)]; liq[IX(j)] = f.h2osoi_liq[OFF(j)];
#pragma unroll
    for (int k = 0; k < 8; ++k) ms[k][IX(j)] = mss[k][OFF(j)];
  }
  ice[NS] = f.h2osoi_ice[OFF(1)]; liq[NS] = f.h2osoi_liq[OFF(1)];
  int snl = f.snl[cc];
  double frac_sno_eff = f.frac_sno_eff[cc];

  // ---- SnowCompaction :1947-2077 ----
  {
    const double c3 = 2.777e-6, c4 = 0.04, c5 = 2.0;
    double burden = 0.0, zpseudo = 0.0;
    bool mobile = true;
    const double frac_sno = frac_sno_eff;
    for (int j = snl + 1; j <= 0; ++j) {
      const double icej = ice[IX(j)], liqj = liq[IX(j)], dzj = dz[IX(j)];
      const double wx = (icej + liqj);
      const double voidf = 1.0 - (icej / denice + liqj / denh2o) / (frac_sno * dzj);
      if (voidf > 0.001 && icej > .1) {
        const double bi = icej / (frac_sno * dzj);
        const double fi = icej / wx;
        const double td = tfrz - tk[IX(j)];
        const double dexpf = exp(-c4 * td);
        double ddz1 = -c3 * dexpf;
        if (bi > prm.upplim) ddz1 = ddz1 * exp(-46.0e-3 * (bi - prm.upplim));
        if (liqj > 0.01 * dzj * frac_sno) ddz1 = ddz1 * c5;
        double ddz2;
        if (prm.method == 1) {                                                    // Anderson1976 :3784-3789
          const double c2 = 23.e-3;
          ddz2 = -(burden + wx / 2.0) * exp(-prm.Tfactor * td - c2 * bi) / prm.eta0_anderson;
        } else {                                                                  // Vionnet2012 :3825-3830
          const double aeta = 0.1, beta = 0.023;
          const double f1 = 1.0 / (1.0 + 60.0 * liqj / (denh2o * dzj));
          const double f2 = 4.0;
          const double eta = f1 * f2 * (bi / prm.ceta) * exp(aeta * td + beta * bi) * prm.eta0_vionnet;
          ddz2 = -(burden + wx / 2.0) / eta;
        }
        double ddz3;
        if (f.imelt[OFF(j)] == 1) {
          if (prm.subgrid) {
            const double swe_old = f.swe_old[OFF(j)];
            ddz3 = fmax(0.0, fmin(1.0, (swe_old - wx) / wx));
            if ((swe_old - wx) > 0.0) {
              double wsum = 0.0;
              for (int jj = snl + 1; jj <= 0; ++jj) wsum += liq[IX(jj)] + ice[IX(jj)];
              // FracSnowDuringMelt, SnowCoverFractionSwensonLawrence2012Mod.F90:263-266
              const double int_snow_limited = fmin(f.int_snow[cc], prm.int_snow_max);
              const double smr = fmin(1.0, wsum / int_snow_limited);
              double fsno_melt = 1. - pow(acos(fmin(1.0, (2. * smr - 1.0))) / rpi, f.n_melt[cc]);
              const double fh = f.frac_h2osfc[cc];
              if ((fsno_melt + fh) > 1.0) fsno_melt = 1.0 - fh;
              ddz3 = ddz3 - fmax(0.0, (fsno_melt - frac_sno) / frac_sno);
            }
            ddz3 = -1.0 / dtime * ddz3;
          } else {
            const double fio = f.frac_iceold[OFF(j)];
            ddz3 = -1.0 / dtime * fmax(0.0, (fio - fi) / fio);
          }
        } else {
          ddz3 = 0.0;
        }
        double ddz4 = 0.0;
        if (prm.wind) {                                                           // WindDriftCompaction :3872-3897
          const double rho_min = 50.0, drift_sph = 1.0;
          if (mobile) {
            const double Frho = 1.25 - 0.0042 * (fmax(rho_min, bi) - rho_min);
            const double MO = 0.34 * (-0.583 * prm.drift_gs - 0.833 * drift_sph + 0.833) + 0.66 * Frho;
            double SI = -2.868 * exp(-0.085 * f.forc_wind[f.col_gridcell[cc] - geo.begg0]) + 1.0 + MO;
            if (SI > 0.0) {
              SI = fmin(SI, 3.25);
              zpseudo = zpseudo + 0.5 * dzj * (3.25 - SI);
              const double gamma_drift = SI * exp(-zpseudo / 0.1);
              const double tau_inverse = gamma_drift / prm.tau_ref;
              ddz4 = -fmax(0.0, prm.rho_max - bi) * tau_inverse;
              zpseudo = zpseudo + 0.5 * dzj * (3.25 - SI);
            } else {
              mobile = false;
              ddz4 = 0.0;
            }
          }
        }
        const double pdzdtc = ddz1 + ddz2 + ddz3 + ddz4;
        dz[IX(j)] = fmax(dzj * (1.0 + pdzdtc * dtime), (icej / denice + liqj / denh2o) / frac_sno);
      } else {
        mobile = false;
      }
      burden = burden + wx;
    }
  }

  // ---- CombineSnowLayers :2186-2503 (dzminloc = dzmin off lakes) ----
  double qflx_sl_top_soil = 0.0;
  {
    const int msn_old = snl;
    for (int j = msn_old + 1; j <= 0; ++j) {                                       // :2203-2283
      if (ice[IX(j)] <= .01) {
        if (j < 0 || soil) {
          liq[IX(j + 1)] = liq[IX(j + 1)] + liq[IX(j)];
          ice[IX(j + 1)] = ice[IX(j + 1)] + ice[IX(j)];
        }
        if (j < 0) {
          dz[IX(j + 1)] = dz[IX(j + 1)] + dz[IX(j)];
#pragma unroll
          for (int k = 0; k < 8; ++k) ms[k][IX(j + 1)] = ms[k][IX(j + 1)] + ms[k][IX(j)];
        }
        if (j == 0) qflx_sl_top_soil = (liq[IX(j)] + ice[IX(j)]) / dtime;
        if (j > snl + 1 && snl < -1) {
          for (int i = j; i >= snl + 2; --i) {
            liq[IX(i)] = liq[IX(i - 1)]; ice[IX(i)] = ice[IX(i - 1)]; tk[IX(i)] = tk[IX(i - 1)];
#pragma unroll
            for (int k = 0; k < 8; ++k) ms[k][IX(i)] = ms[k][IX(i - 1)];
            rds[IX(i)] = rds[IX(i - 1)];
            dz[IX(i)] = dz[IX(i - 1)];
          }
        }
        snl = snl + 1;
      }
    }
  }
  double snow_depth = 0.0, h2osno_total = 0.0;
  bool write_col = false;                                                          // frac_sno / frac_sno_eff / int_snow reset
  double h2osno_no_layers = 0.0;
  bool write_nolayers = false;
  {
    double zwice = 0.0, zwliq = 0.0;
    for (int j = SLO; j <= 0; ++j)
      if (j >= snl + 1) {
        zwice = zwice + ice[IX(j)];
        zwliq = zwliq + liq[IX(j)];
        snow_depth = snow_depth + dz[IX(j)];
        h2osno_total = h2osno_total + ice[IX(j)] + liq[IX(j)];
      }
    if (snow_depth > 0.0) {                                                        // :2329-2369 all snow gone
      if ((frac_sno_eff * snow_depth < prm.dzmin[0]) || (h2osno_total / (frac_sno_eff * snow_depth) < 50.0)) {
        h2osno_no_layers = zwice; write_nolayers = true;
        if (soil) liq[NS] = liq[NS] + zwliq;
        snl = 0;
        h2osno_total = h2osno_no_layers;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          for (int j = 0; j < NS; ++j) ms[k][j] = 0.0;
        if (h2osno_no_layers <= 0.0) snow_depth = 0.0;
      }
    }
    if (h2osno_total <= 0.0) { snow_depth = 0.0; frac_sno_eff = 0.0; write_col = true; }
  }
  if (snl < -1) {                                                                  // :2386-2487
    const int msn_old = snl;
    int mssi = 1;
    for (int i = msn_old + 1; i <= 0; ++i) {
      if ((frac_sno_eff * dz[IX(i)] < prm.dzmin[mssi - 1]) || ((ice[IX(i)] + liq[IX(i)]) / (frac_sno_eff * dz[IX(i)]) < 50.0)) {
        int neibor;
        if (i == snl + 1) neibor = i + 1;
        else if (i == 0) neibor = i - 1;
        else {
          neibor = i + 1;
          if ((dz[IX(i - 1)] + dz[IX(i)]) < (dz[IX(i + 1)] + dz[IX(i)])) neibor = i - 1;
        }
        int j, l;
        if (neibor > i) { j = neibor; l = i; } else { j = i; l = neibor; }
#pragma unroll
        for (int k = 0; k < 8; ++k) ms[k][IX(j)] = ms[k][IX(j)] + ms[k][IX(l)];
        rds[IX(j)] = (rds[IX(j)] * (liq[IX(j)] + ice[IX(j)]) + rds[IX(l)] * (liq[IX(l)] + ice[IX(l)])) /
                     (liq[IX(j)] + ice[IX(j)] + liq[IX(l)] + ice[IX(l)]);
        combo(dz[IX(j)], liq[IX(j)], ice[IX(j)], tk[IX(j)], dz[IX(l)], liq[IX(l)], ice[IX(l)], tk[IX(l)]);
        if (j - 1 > snl + 1) {
          for (int k = j - 1; k >= snl + 2; --k) {
            ice[IX(k)] = ice[IX(k - 1)]; liq[IX(k)] = liq[IX(k - 1)]; tk[IX(k)] = tk[IX(k - 1)];
#pragma unroll
            for (int a = 0; a < 8; ++a) ms[a][IX(k)] = ms[a][IX(k - 1)];
            rds[IX(k)] = rds[IX(k - 1)];
            dz[IX(k)] = dz[IX(k - 1)];
          }
        }
        snl = snl + 1;
        if (snl >= -1) break;
      } else {
        mssi = mssi + 1;
      }
    }
  }

  // ---- DivideSnowLayers :2620-2879 (is_lake = .false.: thicknesses weighted by frac_sno_eff) ----
  if (snl < 0) {
    double dzsno[NS + 1], swice[NS + 1], swliq[NS + 1], tsno[NS + 1], rd[NS + 1], ma[8][NS + 1];
    const int snl0 = snl;
    int msno = -snl0;
    for (int j = 1; j <= NS; ++j) {
      const bool on = j <= msno;
      const int s = on ? IX(j + snl0) : 0;
      dzsno[j] = on ? frac_sno_eff * dz[s] : 0.0;
      swice[j] = on ? ice[s] : 0.0; swliq[j] = on ? liq[s] : 0.0; tsno[j] = on ? tk[s] : 0.0; rd[j] = on ? rds[s] : 0.0;
#pragma unroll
      for (int a = 0; a < 8; ++a) ma[a][j] = on ? ms[a][s] : 0.0;
    }
    int k = 1;
    while (k <= msno && k < NS) {
      if (k == msno) {
        if (dzsno[k] > prm.dzmax_l[k - 1]) {
          msno = msno + 1;
          dzsno[k] = dzsno[k] / 2.0; dzsno[k + 1] = dzsno[k];
          swice[k] = swice[k] / 2.0; swice[k + 1] = swice[k];
          swliq[k] = swliq[k] / 2.0; swliq[k + 1] = swliq[k];
          if (k == 1) {
            tsno[k + 1] = tsno[k];
          } else {
            const double dtdz = (tsno[k - 1] - tsno[k]) / ((dzsno[k - 1] + 2 * dzsno[k]) / 2.0);
            tsno[k + 1] = tsno[k] - dtdz * dzsno[k] / 2.0;
            if (tsno[k + 1] >= tfrz) tsno[k + 1] = tsno[k];
            else tsno[k] = tsno[k] + dtdz * dzsno[k] / 2.0;
          }
#pragma unroll
          for (int a = 0; a < 8; ++a) { ma[a][k] = ma[a][k] / 2.0; ma[a][k + 1] = ma[a][k]; }
          rd[k + 1] = rd[k];
        }
      }
      if (k < msno) {
        if (dzsno[k] > prm.dzmax_u[k - 1]) {
          const double drr = dzsno[k] - prm.dzmax_u[k - 1] - 0.0;
          double propor = drr / dzsno[k];
          const double zwice = propor * swice[k], zwliq = propor * swliq[k];
          double zm[8];
#pragma unroll
          for (int a = 0; a < 8; ++a) zm[a] = propor * ma[a][k];
          propor = (prm.dzmax_u[k - 1] + 0.0) / dzsno[k];
          swice[k] = propor * swice[k];
          swliq[k] = propor * swliq[k];
#pragma unroll
          for (int a = 0; a < 8; ++a) ma[a][k] = propor * ma[a][k];
          dzsno[k] = prm.dzmax_u[k - 1] + 0.0;
#pragma unroll
          for (int a = 0; a < 8; ++a) ma[a][k + 1] = ma[a][k + 1] + zm[a];
          {                                                                        // MassWeightedSnowRadius :3966-3971
            const double swtot = (swliq[k + 1] + swice[k + 1]), zwtot = (zwliq + zwice);
            double r = (rd[k + 1] * swtot + rd[k] * zwtot) / (swtot + zwtot);
            if (r > snw_rds_max) r = snw_rds_max;
            else if (r < prm.snw_rds_min) r = prm.snw_rds_min;
            rd[k + 1] = r;
          }
          combo(dzsno[k + 1], swliq[k + 1], swice[k + 1], tsno[k + 1], drr, zwliq, zwice, tsno[k]);
        }
      }
      k = k + 1;
    }
    snl = -msno;
    for (int j = snl + 1; j <= 0; ++j) {
      const int jj = j - snl;
      dz[IX(j)] = dzsno[jj] / frac_sno_eff;
      ice[IX(j)] = swice[jj]; liq[IX(j)] = swliq[jj]; tk[IX(j)] = tsno[jj]; rds[IX(j)] = rd[jj];
#pragma unroll
      for (int a = 0; a < 8; ++a) ms[a][IX(j)] = ma[a][jj];
    }
  }

  // ---- write back; z / zi from the final thicknesses (:2493-2501 = :2883-2891), ZeroEmptySnowLayers :2935-2947 ----
  const bool zero_empty = snl > -NS;
  double zi_below = f.zi[(size_t)(0 + NS) * ldc + cc];                             // zi(c, 0)
  for (int j = 0; j >= SLO; --j) {
    if (j >= snl + 1) {
      f.z[OFF(j)] = zi_below - 0.5 * dz[IX(j)];
      zi_below = zi_below - dz[IX(j)];
      f.zi[(size_t)(j - 1 + NS) * ldc + cc] = zi_below;
      f.dz[OFF(j)] = dz[IX(j)]; f.t_soisno[OFF(j)] = tk[IX(j)];
      f.h2osoi_ice[OFF(j)] = ice[IX(j)]; f.h2osoi_liq[OFF(j)] = liq[IX(j)];
    } else if (zero_empty) {
      f.z[OFF(j)] = 0.0; f.zi[(size_t)(j - 1 + NS) * ldc + cc] = 0.0; f.dz[OFF(j)] = 0.0; f.t_soisno[OFF(j)] = 0.0;
      f.h2osoi_ice[OFF(j)] = 0.0; f.h2osoi_liq[OFF(j)] = 0.0;
    }
    f.snw_rds[OFF(j)] = rds[IX(j)];
#pragma unroll
    for (int k = 0; k < 8; ++k) mss[k][OFF(j)] = ms[k][IX(j)];
  }
  f.h2osoi_ice[OFF(1)] = ice[NS]; f.h2osoi_liq[OFF(1)] = liq[NS];
  f.snl[cc] = snl;
  f.snow_depth[cc] = snow_depth;
  f.qflx_sl_top_soil[cc] = qflx_sl_top_soil;
  if (write_nolayers) f.h2osno_no_layers[cc] = h2osno_no_layers;
  if (write_col) { f.frac_sno[cc] = 0.0; f.frac_sno_eff[cc] = 0.0; f.int_snow[cc] = 0.0; }
#undef OFF
#undef IX
}
}  // namespace

// InitSnowLayers :2985-3002
static void snow_dz_limits(const ctsm_params_t& p, double* dzmin, double* dzmax_l, double* dzmax_u) {
  dzmin[0] = p.snow_dzmin_1; dzmax_l[0] = p.snow_dzmax_l_1; dzmax_u[0] = p.snow_dzmax_u_1;
  dzmin[1] = p.snow_dzmin_2; dzmax_l[1] = p.snow_dzmax_l_2; dzmax_u[1] = p.snow_dzmax_u_2;
  for (int j = 2; j < NS; ++j) {
    dzmin[j] = dzmax_u[j - 1] * 0.5;
    dzmax_u[j] = 2.0 * dzmax_u[j - 1] + 0.01;
    dzmax_l[j] = dzmax_u[j] + dzmax_l[j - 1];
    if (j == NS - 1) { dzmax_u[j] = 1.79769313486231571e308; dzmax_l[j] = 1.79769313486231571e308; }
  }
}

extern "C" int ctsm_b200_snow_water(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_snowc, const int32_t* filter_snowc,
                                    int num_nosnowc, const int32_t* filter_nosnowc, const ctsm_snowwater_fields_t* hf, int mem,
                                    ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_snowc < 0 || num_nosnowc < 0 || (num_snowc > 0 && !filter_snowc) ||
      (num_nosnowc > 0 && !filter_nosnowc))
    return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  SnowWaterDev d;
  const int32_t *dfs = filter_snowc, *dfn = filter_nosnowc;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_SNOWWATER
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SNOWWATER
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_snowc, num_snowc, &dfs);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter1, filter_nosnowc, num_nosnowc, &dfn);
    if (rc) return rc;
  }
  const ctsm_params_t& p = ctx->prm;
  const SnowGeo geo{hf->alloc.begc, hf->alloc.begg, hf->alloc.endc - hf->alloc.begc + 1, hf->alloc.endg - hf->alloc.begg + 1};
  SnowWaterPrm prm{p.dtime, p.wimp, p.ssi, p.scvng_fct_mlt_sf,
                   {p.scvng_fct_mlt_bcphi, p.scvng_fct_mlt_bcpho, scvng_fct_mlt_ocphi, scvng_fct_mlt_ocpho, p.scvng_fct_mlt_dst1,
                    p.scvng_fct_mlt_dst2, p.scvng_fct_mlt_dst3, p.scvng_fct_mlt_dst4},
                   p.snicar_use_aerosol};
  const int nb = bounds->endc - bounds->begc + 1;
  if (nb > 0) {
    aerosol_dep_kernel<<<grid_for(nb, 256), 256, 0, ctx->stream>>>(d, geo, bounds->begc, bounds->endc, p.snicar_use_aerosol != 0);
    ctx->launches++;
  }
  if (num_snowc + num_nosnowc > 0) {
    snow_water_kernel<<<grid_for(num_snowc + num_nosnowc, 128), 128, 0, ctx->stream>>>(d, geo, prm, num_snowc, dfs, num_nosnowc, dfn,
                                                                                       ctx->d_status);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

extern "C" int ctsm_b200_snow_layers(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_snowc, const int32_t* filter_snowc,
                                     const ctsm_snowlayers_fields_t* hf, int mem, ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_snowc < 0 || (num_snowc > 0 && !filter_snowc)) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  SnowLayersDev d;
  const int32_t* dfs = filter_snowc;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_SNOWLAYERS
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SNOWLAYERS
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_snowc, num_snowc, &dfs);
    if (rc) return rc;
  }
  const ctsm_params_t& p = ctx->prm;
  const SnowGeo geo{hf->alloc.begc, hf->alloc.begg, hf->alloc.endc - hf->alloc.begc + 1, hf->alloc.endg - hf->alloc.begg + 1};
  SnowLayersPrm prm;
  prm.dtime = p.dtime; prm.int_snow_max = p.int_snow_max; prm.upplim = p.upplim_destruct_metamorph;
  prm.Tfactor = p.overburden_compress_Tfactor; prm.eta0_anderson = p.eta0_anderson; prm.eta0_vionnet = p.eta0_vionnet;
  prm.ceta = p.ceta; prm.drift_gs = p.drift_gs; prm.tau_ref = p.tau_ref; prm.rho_max = p.rho_max; prm.snw_rds_min = p.snw_rds_min;
  snow_dz_limits(p, prm.dzmin, prm.dzmax_l, prm.dzmax_u);
  prm.method = p.snow_overburden_compaction_method; prm.wind = p.wind_dependent_snow_density; prm.subgrid = p.use_subgrid_fluxes;
  if (num_snowc > 0) {
    snow_layers_kernel<<<grid_for(num_snowc, 128), 128, 0, ctx->stream>>>(d, geo, prm, num_snowc, dfs, ctx->d_status);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

extern "C" int ctsm_b200_snow_capping(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_initc, const int32_t* filter_initc,
                                      int num_snowc, const int32_t* filter_snowc, const ctsm_snowcapping_fields_t* hf, int nstep,
                                      int mem, ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_initc < 0 || num_snowc < 0 || (num_initc > 0 && !filter_initc) || (num_snowc > 0 && !filter_snowc))
    return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  SnowCappingDev d;
  const int32_t *dfi = filter_initc, *dfs = filter_snowc;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_SNOWCAPPING
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SNOWCAPPING
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_initc, num_initc, &dfi);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter1, filter_snowc, num_snowc, &dfs);
    if (rc) return rc;
  }
  const ctsm_params_t& p = ctx->prm;
  // SnowCappingExcess :3447-3454: reset_snow_timesteps_per_layer = 4
  const int reset_active = ((p.reset_snow || p.reset_snow_glc) && nstep <= 4 * NS) ? 1 : 0;
  const SnowCapPrm prm{p.dtime, p.h2osno_max, p.reset_snow_glc_ela, p.reset_snow, p.reset_snow_glc, reset_active};
  const int begc0 = hf->alloc.begc, ldc = hf->alloc.endc - hf->alloc.begc + 1;
  if (num_initc > 0) {
    snow_capping_init_kernel<<<grid_for(num_initc, 256), 256, 0, ctx->stream>>>(d, begc0, num_initc, dfi);
    ctx->launches++;
  }
  if (num_snowc > 0) {
    snow_capping_kernel<<<grid_for(num_snowc, 128), 128, 0, ctx->stream>>>(d, prm, begc0, ldc, num_snowc, dfs, ctx->d_status);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}
