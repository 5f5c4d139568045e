// balance.cu — BalanceCheck / EnergyBalanceCheck and the PHS root-water sink.
//
// Reference:
//   BalanceCheck            src/biogeophys/BalanceCheckMod.F90:445-857
//   EnergyBalanceCheck      :859-1119        BalanceCheckInit :74-95
//   c2g_1d                  src/main/subgridAveMod.F90:761-818 ('urbanf','unity'; non-urban)
//   CalculateTotalH2osno    src/biogeophys/WaterStateType.F90:887-896
//   Compute_EffecRootFrac_And_VertTranSink_HydStress   src/biogeophys/SoilWaterPlantSinkMod.F90:236-328
//
// B200 mapping.  The residuals are embarrassingly parallel maps (one thread per column / gridcell /
// patch, all fields 1-D and coalesced, HBM-bound at ~0.4 KB per column); the reference's
// maxval/maxloc over the clump become a two-pass device reduction: pass 1 folds |err| into a
// 64-bit atomicMax (the IEEE bit pattern of a non-negative double orders like an unsigned
// integer), pass 2 finds the lowest index attaining it (maxloc returns the first).  The warning /
// abort decision is taken on the host from 7 (value, index) pairs, as the reference does after
// its reductions.  The gridcell aggregation c2g loops over the gridcell's own contiguous columns
// in ascending index order (the reference's summation order; no atomics).
#include "common.cuh"

struct BalanceDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_BALANCECHECK
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_BALANCECHECK
#undef CTSM_F
};
struct PlantSinkDefaultDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_PLANTSINKDEFAULT
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_PLANTSINKDEFAULT
#undef CTSM_F
};
struct PlantSinkDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_PLANTSINK
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_PLANTSINK
#undef CTSM_F
};

namespace {
constexpr double spval = 1.e36;
struct BGeo { int begc0, begp0, begg0, ldc, ldg, begc, endc, begp, endp, begg, endg; };
struct Red { unsigned long long mx[CTSM_BAL_NKIND]; int idx[CTSM_BAL_NKIND]; int pad; };

__device__ __forceinline__ void fold_max(unsigned long long* slot, double v) {
  // warp-level max first, one atomic per warp
  unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v));
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
    b = t > b ? t : b;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(slot, b);
}

// columns: errh2o_col :592-608, snow balance :754-803 (+ CalculateTotalH2osno), errsoi max :1101
__global__ void __launch_bounds__(256)
balance_col_kernel(BalanceDev f, BGeo g, double dtime, const int* __restrict__ in_allc, Red* red) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = g.endc - g.begc + 1;
  double e1 = 0.0, e2 = 0.0, e3 = 0.0;
  if (i < n) {
    const int cc = g.begc - g.begc0 + i;
    const bool active = f.col_active[cc] != 0;
    if (active) {
      e1 = f.endwb[cc] - f.begwb[cc]
           - (f.forc_rain[cc] + f.forc_snow[cc] + f.qflx_flood[cc] + f.qflx_sfc_irrig[cc] + f.qflx_glcice_dyn_water_flux[cc]
              - f.qflx_evap_tot[cc] - f.qflx_surf[cc] - f.qflx_qrgwl[cc] - f.qflx_drain[cc] - f.qflx_drain_perched[cc]
              - f.qflx_ice_runoff[cc] - f.qflx_snwcp_discarded_liq[cc] - f.qflx_snwcp_discarded_ice[cc]) * dtime;
      const int snl = f.snl[cc];
      if (snl < 0) {
        // h2osno_total is only defined over filter_allc in the reference (:752); outside it the local array is unset
        double tot = f.h2osno_no_layers[cc];
        for (int j = snl + 1; j <= 0; ++j)
          tot = tot + f.h2osoi_ice[(size_t)(j - SNOSOI_LO) * g.ldc + cc] + f.h2osoi_liq[(size_t)(j - SNOSOI_LO) * g.ldc + cc];
        const int lt = f.lun_itype[cc];
        const double fs = f.frac_sno_eff[cc];
        const double sdew = f.qflx_soliddew_to_top_layer[cc], ldew = f.qflx_liqdew_to_top_layer[cc];
        const double sev = f.qflx_solidevap_from_top_layer[cc], lev = f.qflx_liqevap_from_top_layer[cc];
        const double drain = f.qflx_snow_drain[cc], cpi = f.qflx_snwcp_ice[cc], cpl = f.qflx_snwcp_liq[cc];
        const double di = f.qflx_snwcp_discarded_ice[cc], dl = f.qflx_snwcp_discarded_liq[cc], sl = f.qflx_sl_top_soil[cc];
        double src = f.qflx_prec_grnd[cc] + sdew + ldew;
        double snk = sev + lev + drain + cpi + cpl + di + dl + sl;
        if (lt == CTSM_ISTDLAK) {
          src = f.qflx_snow_grnd[cc] + fs * (f.qflx_liq_grnd[cc] + sdew + ldew);
          snk = fs * (sev + lev) + cpi + cpl + di + dl + drain + sl;
        }
        if (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP || lt == CTSM_ISTWET || lt == CTSM_ISTICE) {
          src = (f.qflx_snow_grnd[cc] - f.qflx_snow_h2osfc[cc]) + fs * (f.qflx_liq_grnd[cc] + sdew + ldew) + f.qflx_h2osfc_to_ice[cc];
          snk = fs * (sev + lev) + cpi + cpl + di + dl + drain + sl;
        }
        f.snow_sources[cc] = src;
        f.snow_sinks[cc] = snk;
        e2 = (tot - f.h2osno_old[cc]) - (src - snk) * dtime;
        (void)in_allc;
      } else {
        f.snow_sources[cc] = 0.0;
        f.snow_sinks[cc] = 0.0;
      }
      e3 = f.errsoi_col[cc];
    }
    f.errh2o[cc] = e1;
    f.errh2osno[cc] = e2;
  }
  fold_max(&red->mx[CTSM_BAL_H2O_COL], e1);
  fold_max(&red->mx[CTSM_BAL_H2OSNO], e2);
  fold_max(&red->mx[CTSM_BAL_SOI], e3);
}

// gridcells: c2g of three column fluxes + errh2o_grc :666-694
__global__ void __launch_bounds__(256)
balance_grc_kernel(BalanceDev f, BGeo g, double dtime, Red* red) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = g.endg - g.begg + 1;
  double e = 0.0;
  if (i < n) {
    const int gg = g.begg - g.begg0 + i;
    double acc[3] = {spval, spval, spval}, sw[3] = {0.0, 0.0, 0.0};
    for (int c1 = f.grc_coli[gg]; c1 <= f.grc_colf[gg]; ++c1) {
      const int cc = c1 - g.begc0;
      const double wt = f.wtgcell[cc];
      if (f.col_active[cc] && wt != 0.0) {
        const double v[3] = {f.qflx_glcice_dyn_water_flux[cc], f.qflx_snwcp_discarded_liq[cc], f.qflx_snwcp_discarded_ice[cc]};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (v[k] != spval) {
            if (sw[k] == 0.0) acc[k] = 0.0;
            acc[k] = acc[k] + v[k] * 1.0 * 1.0 * wt;
            sw[k] = sw[k] + wt;
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (!(sw[k] > 1.0 + 1.e-6) && sw[k] != 0.0) acc[k] = acc[k] / sw[k];
    e = f.endwb_grc[gg] - f.begwb_grc[gg]
        - (f.forc_rain_grc[gg] + f.forc_snow_grc[gg] + f.forc_flood_grc[gg] + f.qflx_sfc_irrig_grc[gg] + acc[0]
           - f.qflx_evap_tot_grc[gg] - f.qflx_surf_grc[gg] - f.qflx_qrgwl_grc[gg] - f.qflx_drain_grc[gg]
           - f.qflx_drain_perched_grc[gg] - f.qflx_ice_runoff_grc[gg] - acc[1] - acc[2]) * dtime;
    f.errh2o_grc[gg] = e;
  }
  fold_max(&red->mx[CTSM_BAL_H2O_GRC], e);
}

// patches: errsol / errlon / errseb / netrad :962-1005 (non-urban)
__global__ void __launch_bounds__(256)
balance_patch_kernel(BalanceDev f, BGeo g, Red* red) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = g.endp - g.begp + 1;
  double e1 = 0.0, e2 = 0.0, e3 = 0.0;
  if (i < n) {
    const int pp = g.begp - g.begp0 + i;
    if (f.patch_active[pp]) {
      const int cc = f.column[pp] - g.begc0, gg = f.gridcell[pp] - g.begg0;
      const double lw = f.forc_lwrad[cc], out = f.eflx_lwrad_out[pp], net = f.eflx_lwrad_net[pp], fsa = f.fsa[pp];
      e1 = fsa + f.fsr[pp] - (f.forc_solad[cc] + f.forc_solad[(size_t)g.ldc + cc] + f.forc_solai[gg] + f.forc_solai[(size_t)g.ldg + gg]);
      e2 = out - net - lw;
      e3 = f.sabv[pp] + f.sabg_chk[pp] + lw - out - f.eflx_sh_tot[pp] - f.eflx_lh_tot[pp] - f.eflx_soil_grnd[pp] - f.dhsdt_canopy[pp];
      f.netrad[pp] = fsa - net;
    }
    f.errsol[pp] = e1; f.errlon[pp] = e2; f.errseb[pp] = e3;
  }
  fold_max(&red->mx[CTSM_BAL_SOL], e1 == spval ? 0.0 : e1);
  fold_max(&red->mx[CTSM_BAL_LON], e2 == spval ? 0.0 : e2);
  fold_max(&red->mx[CTSM_BAL_SEB], e3);
}

// maxloc: lowest index whose |err| equals the maximum
__global__ void __launch_bounds__(256)
balance_loc_kernel(BalanceDev f, BGeo g, Red* red) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nc = g.endc - g.begc + 1, ng = g.endg - g.begg + 1, np = g.endp - g.begp + 1;
  auto hit = [&](int kind, double v, int index1) {
    if ((unsigned long long)__double_as_longlong(fabs(v)) == red->mx[kind]) atomicMin(&red->idx[kind], index1);
  };
  if (i < nc) {
    const int cc = g.begc - g.begc0 + i;
    hit(CTSM_BAL_H2O_COL, f.errh2o[cc], g.begc + i);
    hit(CTSM_BAL_H2OSNO, f.errh2osno[cc], g.begc + i);
    if (f.col_active[cc]) hit(CTSM_BAL_SOI, f.errsoi_col[cc], g.begc + i);
  }
  if (i < ng) hit(CTSM_BAL_H2O_GRC, f.errh2o_grc[g.begg - g.begg0 + i], g.begg + i);
  if (i < np) {
    const int pp = g.begp - g.begp0 + i;
    if (f.errsol[pp] != spval) hit(CTSM_BAL_SOL, f.errsol[pp], g.begp + i);
    if (f.errlon[pp] != spval) hit(CTSM_BAL_LON, f.errlon[pp], g.begp + i);
    hit(CTSM_BAL_SEB, f.errseb[pp], g.begp + i);
  }
}

// Compute_EffecRootFrac_And_VertTranSink_HydStress: column-outer, contiguous patches inner (ascending => the
// reference's summation order)
__global__ void __launch_bounds__(128)
plantsink_kernel(PlantSinkDev f, int begc0, int ldc, int begp0, int ldp, int numf, const int32_t* __restrict__ filterc) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int cc = filterc[fc] - begc0;
  const int pi = f.patchi[cc], pf = pi + f.npatches[cc] - 1;
  for (int p1 = pi; p1 <= pf; ++p1) f.qflx_hydr_redist[p1 - begp0] = 0.0;
  double neg = 0.0;
  for (int j = 1; j <= NLEVSOI; ++j) {
    const double grav2 = f.z[(size_t)(j - SNOSOI_LO) * ldc + cc] * 1000.0;
    const double smp = f.smp_l[(size_t)(j - 1) * ldc + cc];
    double temp = 0.0;
    for (int p1 = pi; p1 <= pf; ++p1) {
      const int pp = p1 - begp0;
      if (f.patch_active[pp] && f.frac_veg_nosno[pp] > 0) {
        const double wt = f.wtcol[pp];
        if (wt > 0.0) {
          const double flux = f.k_soil_root[(size_t)(j - 1) * ldp + pp] * (smp - f.vegwp[(size_t)3 * ldp + pp] - grav2);
          if (flux < 0) f.qflx_hydr_redist[pp] = f.qflx_hydr_redist[pp] + flux;
          temp = temp + flux * wt;
        }
      }
    }
    f.qflx_rootsoi[(size_t)(j - 1) * ldc + cc] = temp;
    if (temp < 0.0) neg = neg + temp;
  }
  f.qflx_phs_neg[cc] = neg;
}

// The same routine with one WARP per column (default; CTSM_B200_SINK_WARP=0 selects the thread-per-column kernel above, and
// tests/test_gpu_balance.py proves the two bit-identical).  The routine is a stream over k_soil_root(p, 1:nlevsoi) - 71 % of its
// bytes - and a thread per column reads it with a stride of the column's patch count (15 doubles) between lanes.  Here lane i owns
// patch patchi + i: the 20 level loads of a warp are contiguous runs of the column's patches (consecutive warps continue them),
// all 20 are in flight before the first is used, and the reference's ascending-patch summation order is kept by handing the
// contributions through shared memory to lane j, which adds level j's terms one by one.
constexpr int SINK_WARPS = 8;                 // warps per block
constexpr int SINK_LD = 33;                   // padded lane stride of a level's row (bank-conflict-free transposition)
constexpr int SINK_HALF = NLEVSOI / 2;        // levels handed over per phase (halves the shared memory: 5 blocks per SM)
static_assert(NLEVSOI % 2 == 0 && NLEVSOI <= 32, "plantsink_warp_kernel: level halves across the lanes");
__global__ void __launch_bounds__(SINK_WARPS * 32, 4)
plantsink_warp_kernel(PlantSinkDev f, int begc0, int ldc, int begp0, int ldp, int numf, const int32_t* __restrict__ filterc) {
  __shared__ double s_c[SINK_WARPS][SINK_HALF * SINK_LD];     // [level of the half][lane] contributions flux * wtcol
  __shared__ double s_col[SINK_WARPS][2 * NLEVSOI];           // smp_l(c, j), grav2(j); later temp(j)
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int fc = blockIdx.x * SINK_WARPS + w;
  if (fc >= numf) return;                                     // warp-uniform
  const int cc = filterc[fc] - begc0;
  const int pi = f.patchi[cc], np = f.npatches[cc];
  if (lane < NLEVSOI) {
    s_col[w][lane] = f.smp_l[(size_t)lane * ldc + cc];
    s_col[w][NLEVSOI + lane] = f.z[(size_t)(lane + 1 - SNOSOI_LO) * ldc + cc] * 1000.0;
  }
  __syncwarp();
  double temp = 0.0;                                          // lane j - 1 accumulates level j
  for (int base = 0; base < np; base += 32) {
    const bool mine = base + lane < np;
    const int pp = pi + base + lane - begp0;
    bool q = false;
    double wt = 0.0, vw = 0.0, redist = 0.0;
    double k[NLEVSOI];
    if (mine) {                                               // the level loads do not wait for the predicate's loads
#pragma unroll
      for (int j = 0; j < NLEVSOI; ++j) k[j] = __ldg(&f.k_soil_root[(size_t)j * ldp + pp]);
      vw = f.vegwp[(size_t)3 * ldp + pp];
      wt = f.wtcol[pp];
      q = f.patch_active[pp] && f.frac_veg_nosno[pp] > 0 && wt > 0.0;
    }
    unsigned qm = __ballot_sync(FULL, q);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (q) {
#pragma unroll
        for (int jj = 0; jj < SINK_HALF; ++jj) {
          const int j = h * SINK_HALF + jj;
          const double flux = k[j] * (s_col[w][j] - vw - s_col[w][NLEVSOI + j]);
          if (flux < 0) redist = redist + flux;
          s_c[w][jj * SINK_LD + lane] = flux * wt;
        }
      }
      __syncwarp();
      if (lane >= h * SINK_HALF && lane < (h + 1) * SINK_HALF) {
        unsigned m = qm;
        while (m) {                                           // ascending patches: the reference's order
          const int i = __ffs(m) - 1;
          m &= m - 1;
          temp = temp + s_c[w][(lane - h * SINK_HALF) * SINK_LD + i];
        }
      }
      __syncwarp();
    }
    if (mine) f.qflx_hydr_redist[pp] = redist;
  }
  if (lane < NLEVSOI) {
    f.qflx_rootsoi[(size_t)lane * ldc + cc] = temp;
    s_col[w][lane] = temp;
  }
  __syncwarp();
  if (lane == 0) {
    double neg = 0.0;
#pragma unroll
    for (int j = 0; j < NLEVSOI; ++j) { const double t = s_col[w][j]; if (t < 0.0) neg = neg + t; }
    f.qflx_phs_neg[cc] = neg;
  }
}

// Compute_EffecRootFrac_And_VertTranSink_Default (SoilWaterPlantSinkMod.F90:332-424): one thread per column, its contiguous
// patches inner and ascending (the reference's summation order).  Levels nlevsoi+1..nlevgrnd of rootr_col are not touched.
__global__ void __launch_bounds__(128)
plantsink_default_kernel(PlantSinkDefaultDev f, int begc0, int ldc, int begp0, int ldp, int numf, const int32_t* __restrict__ filterc) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int cc = filterc[fc] - begc0;
  const int pi = f.patchi[cc], pf = pi + f.npatches[cc] - 1;
  double temp = 0.0;
  for (int p1 = pi; p1 <= pf; ++p1) {
    const int pp = p1 - begp0;
    if (f.patch_active[pp]) temp = temp + f.qflx_tran_veg[pp] * f.wtcol[pp];
  }
  const double tran_col = f.qflx_tran_veg_col[cc];
  for (int j = 1; j <= NLEVSOI; ++j) {
    double r = 0.0;
    for (int p1 = pi; p1 <= pf; ++p1) {
      const int pp = p1 - begp0;
      if (f.patch_active[pp]) r = r + f.rootr[(size_t)(j - 1) * ldp + pp] * f.qflx_tran_veg[pp] * f.wtcol[pp];
    }
    if (temp != 0.0) r = r / temp;
    f.rootr_col[(size_t)(j - 1) * ldc + cc] = r;
    f.qflx_rootsoi[(size_t)(j - 1) * ldc + cc] = r * tran_col;
  }
}
}  // namespace

struct WaterBalanceDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_WATERBALANCE
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_WATERBALANCE
#undef CTSM_F
};
struct WaterGridDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_WATERGRIDBALANCE
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_WATERGRIDBALANCE
#undef CTSM_F
};

namespace {
// ComputeLiqIceMassNonLake (TotalWaterAndHeatMod.F90:200-326) + AccumulateSoilLiqIceMassNonLake (:329-393) for one non-urban
// column: level sums ascending (the reference's level-outer / column-inner sweep visits a column's levels in that order),
// canopy water by p2c over the column's contiguous patches.  HBM bound: 2 x 37 level reads + 25 excess-ice reads per
// column, coalesced.  F = WaterBalanceDev or WaterGridDev (same member names).
template <class F>
__device__ __forceinline__ void water_mass_nonlake(const F& f, int cc, size_t ldc, int begp0, double aquifer_water_baseline,
                                                   double& liquid_mass, double& ice_mass) {
  double liqcan_col = 0.0, snocan_col = 0.0;
  const int pi = f.patchi[cc], pf = f.patchf[cc];
  for (int p1 = pi; p1 <= pf; ++p1) {
    const int pp = p1 - begp0;
    if (f.patch_active[pp]) {
      const double wt = f.wtcol[pp];
      liqcan_col = liqcan_col + f.liqcan[pp] * wt;
      snocan_col = snocan_col + f.snocan[pp] * wt;
    }
  }
  liquid_mass = 0.0; ice_mass = 0.0;
  liquid_mass = liquid_mass + liqcan_col + f.total_plant_stored_h2o[cc];
  ice_mass = ice_mass + snocan_col;
  ice_mass = ice_mass + f.h2osno_no_layers[cc];
  const int snl = f.snl[cc];
  for (int j = snl + 1; j <= 0; ++j) {
    liquid_mass = liquid_mass + f.h2osoi_liq[(size_t)(j - SNOSOI_LO) * ldc + cc];
    ice_mass = ice_mass + f.h2osoi_ice[(size_t)(j - SNOSOI_LO) * ldc + cc];
  }
  if (f.col_hydrologically_active[cc]) liquid_mass = liquid_mass + (f.wa[cc] - aquifer_water_baseline);
  liquid_mass = liquid_mass + f.h2osfc[cc];
  for (int j = 1; j <= NLEVGRND; ++j) {
    liquid_mass = liquid_mass + f.h2osoi_liq[(size_t)(j - SNOSOI_LO) * ldc + cc];
    ice_mass = ice_mass + f.h2osoi_ice[(size_t)(j - SNOSOI_LO) * ldc + cc] + f.excess_ice[(size_t)(j - 1) * ldc + cc];
  }
}
// snow and soil layers of a lake column: ComputeLiqIceMassLake :449-462
template <class F>
__device__ __forceinline__ void water_mass_lake_layers(const F& f, int cc, size_t ldc, double& liquid_mass, double& ice_mass) {
  ice_mass = ice_mass + f.h2osno_no_layers[cc];
  const int snl = f.snl[cc];
  for (int j = snl + 1; j <= 0; ++j) {
    liquid_mass = liquid_mass + f.h2osoi_liq[(size_t)(j - SNOSOI_LO) * ldc + cc];
    ice_mass = ice_mass + f.h2osoi_ice[(size_t)(j - SNOSOI_LO) * ldc + cc];
  }
  for (int j = 1; j <= NLEVGRND; ++j) {
    liquid_mass = liquid_mass + f.h2osoi_liq[(size_t)(j - SNOSOI_LO) * ldc + cc];
    ice_mass = ice_mass + f.h2osoi_ice[(size_t)(j - SNOSOI_LO) * ldc + cc];
  }
}
template <class F>
__device__ __forceinline__ double total_h2osno(const F& f, int cc, size_t ldc) {      // CalculateTotalH2osno, WaterStateType.F90:887-896
  double t = f.h2osno_no_layers[cc];
  const int snl = f.snl[cc];
  for (int j = snl + 1; j <= 0; ++j)
    t = t + f.h2osoi_ice[(size_t)(j - SNOSOI_LO) * ldc + cc] + f.h2osoi_liq[(size_t)(j - SNOSOI_LO) * ldc + cc];
  return t;
}

// BeginWaterColumnBalanceSingle (BalanceCheckMod.F90:353-442): one thread per column of the non-lake filter, then of the lake filter
__global__ void __launch_bounds__(128)
water_mass_kernel(WaterBalanceDev f, double aquifer_water_baseline, int begc0, int ldc_, int begp0, int numc,
                  const int32_t* __restrict__ filterc, int lake, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numc) return;
  const int c1 = filterc[fc], cc = c1 - begc0;
  const size_t ldc = (size_t)ldc_;
  double liquid_mass = 0.0, ice_mass = 0.0;
  if (lake) {
    water_mass_lake_layers(f, cc, ldc, liquid_mass, ice_mass);
  } else {
    const int lt = f.lun_itype[cc];
    if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
    water_mass_nonlake(f, cc, ldc, begp0, aquifer_water_baseline, liquid_mass, ice_mass);
  }
  f.begwb[cc] = liquid_mass + ice_mass;
  f.h2osno_old[cc] = total_h2osno(f, cc, ldc);
}

// WaterGridcellBalanceSingle (BalanceCheckMod.F90:212-350), step 1: column water mass with the dynbal baselines subtracted
// and, for lake columns, the lake water added (AccumulateLiqIceMassLake, TotalWaterAndHeatMod.F90:536-544)
__global__ void __launch_bounds__(128)
watergrid_col_kernel(WaterGridDev f, double aquifer_water_baseline, int begc0, int ldc_, int begp0, int begc_call, int numc,
                     const int32_t* __restrict__ filterc, int lake, double* __restrict__ wb_col, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numc) return;
  const int c1 = filterc[fc], cc = c1 - begc0;
  const size_t ldc = (size_t)ldc_;
  double liquid_mass = 0.0, ice_mass = 0.0;
  if (lake) {
    liquid_mass = liquid_mass - f.dynbal_baseline_liq[cc];
    ice_mass = ice_mass - f.dynbal_baseline_ice[cc];
    for (int j = 1; j <= CTSM_NLEVLAK; ++j) {
      const double dzl = f.dz_lake[(size_t)(j - 1) * ldc + cc], fr = f.lake_icefrac[(size_t)(j - 1) * ldc + cc];
      const double h2olak_liq = dzl * cst::denh2o * (1 - fr) * 1.0;
      const double h2olak_ice = dzl * cst::denh2o * fr * 1.0;
      liquid_mass = liquid_mass + h2olak_liq;
      ice_mass = ice_mass + h2olak_ice;
    }
    water_mass_lake_layers(f, cc, ldc, liquid_mass, ice_mass);
  } else {
    const int lt = f.lun_itype[cc];
    if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
    water_mass_nonlake(f, cc, ldc, begp0, aquifer_water_baseline, liquid_mass, ice_mass);
    liquid_mass = liquid_mass - f.dynbal_baseline_liq[cc];
    ice_mass = ice_mass - f.dynbal_baseline_ice[cc];
  }
  wb_col[c1 - begc_call] = liquid_mass + ice_mass;
}
// step 2: c2g (subgridAveMod.F90:791-816, scale factors 1 off the urban landunits), minus the dribblers' remainders
__global__ void __launch_bounds__(128)
watergrid_grc_kernel(WaterGridDev f, int begg0, int begc0, int begg, int endg, int begc, int endc, int flag_endwb,
                     const double* __restrict__ wb_col, DevStatus* ds) {
  const int g1 = begg + blockIdx.x * blockDim.x + threadIdx.x;
  if (g1 > endg) return;
  const int gg = g1 - begg0;
  double garr = 1.0e36, sumwt = 0.0;
  for (int c1 = f.grc_coli[gg]; c1 <= f.grc_colf[gg]; ++c1) {
    if (c1 < begc || c1 > endc) continue;
    const int cc = c1 - begc0;
    const double wt = f.wtgcell[cc];
    if (f.col_active[cc] && wt != 0.0) {
      const double v = wb_col[c1 - begc];
      if (v != 1.0e36) {
        if (sumwt == 0.0) garr = 0.0;
        garr = garr + v * 1.0 * 1.0 * wt;
        sumwt = sumwt + wt;
      }
    }
  }
  if (sumwt > 1.0 + 1.e-6) report_failure(ds, g1, CTSM_ERR_BALANCE, 0);
  else if (sumwt != 0.0) garr = garr / sumwt;
  const double wb = garr - f.qflx_liq_dynbal_left_to_dribble[gg] - f.qflx_ice_dynbal_left_to_dribble[gg];
  if (flag_endwb) f.endwb_grc[gg] = wb - 0.0;
  else f.begwb_grc[gg] = wb;
}
}  // namespace

extern "C" int ctsm_b200_begin_water_column_balance(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec,
                                                    const int32_t* filter_nolakec, int num_lakec, const int32_t* filter_lakec,
                                                    const ctsm_waterbalance_fields_t* hf, double aquifer_water_baseline, int mem,
                                                    ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_nolakec < 0 || (num_nolakec > 0 && !filter_nolakec) || num_lakec < 0 ||
      (num_lakec > 0 && !filter_lakec))
    return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  WaterBalanceDev d;
  const int32_t *dfilter = filter_nolakec, *dlake = filter_lakec;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_WATERBALANCE
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_WATERBALANCE
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_nolakec, num_nolakec, &dfilter);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter1, filter_lakec, num_lakec, &dlake);
    if (rc) return rc;
  }
  const int begc0 = hf->alloc.begc, ldc = hf->alloc.endc - hf->alloc.begc + 1, begp0 = hf->alloc.begp;
  if (num_nolakec > 0) {
    water_mass_kernel<<<grid_for(num_nolakec, 128), 128, 0, ctx->stream>>>(d, aquifer_water_baseline, begc0, ldc, begp0,
                                                                            num_nolakec, dfilter, 0, ctx->d_status);
    ctx->launches++;
  }
  if (num_lakec > 0) {
    water_mass_kernel<<<grid_for(num_lakec, 128), 128, 0, ctx->stream>>>(d, aquifer_water_baseline, begc0, ldc, begp0,
                                                                          num_lakec, dlake, 1, ctx->d_status);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

extern "C" int ctsm_b200_water_gridcell_balance(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec,
                                                const int32_t* filter_nolakec, int num_lakec, const int32_t* filter_lakec,
                                                const ctsm_watergridbalance_fields_t* hf, double aquifer_water_baseline,
                                                int flag_endwb, int mem, ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_nolakec < 0 || (num_nolakec > 0 && !filter_nolakec) || num_lakec < 0 ||
      (num_lakec > 0 && !filter_lakec))
    return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  WaterGridDev d;
  const int32_t *dfilter = filter_nolakec, *dlake = filter_lakec;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_WATERGRIDBALANCE
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_WATERGRIDBALANCE
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_nolakec, num_nolakec, &dfilter);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter1, filter_lakec, num_lakec, &dlake);
    if (rc) return rc;
  }
  const int ncb = bounds->endc - bounds->begc + 1, ngb = bounds->endg - bounds->begg + 1;
  if (ncb > 0 && ngb > 0) {
    int rc = arena_reserve(ctx, ctx->arena_scratch, sizeof(double) * (size_t)ncb);
    if (rc) return rc;
    double* wb_col = (double*)ctx->arena_scratch.p;
    cudaStream_t s = ctx->stream;
    CUDA_TRY(cudaMemsetAsync(wb_col, 0, sizeof(double) * (size_t)ncb, s));      // columns in neither filter
    const int begc0 = hf->alloc.begc, ldc = hf->alloc.endc - hf->alloc.begc + 1, begp0 = hf->alloc.begp;
    if (num_nolakec > 0)
      watergrid_col_kernel<<<grid_for(num_nolakec, 128), 128, 0, s>>>(d, aquifer_water_baseline, begc0, ldc, begp0, bounds->begc,
                                                                      num_nolakec, dfilter, 0, wb_col, ctx->d_status);
    if (num_lakec > 0)
      watergrid_col_kernel<<<grid_for(num_lakec, 128), 128, 0, s>>>(d, aquifer_water_baseline, begc0, ldc, begp0, bounds->begc,
                                                                    num_lakec, dlake, 1, wb_col, ctx->d_status);
    watergrid_grc_kernel<<<grid_for(ngb, 128), 128, 0, s>>>(d, hf->alloc.begg, begc0, bounds->begg, bounds->endg, bounds->begc,
                                                            bounds->endc, flag_endwb, wb_col, ctx->d_status);
    ctx->launches += (num_nolakec > 0) + (num_lakec > 0) + 1;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

extern "C" int ctsm_b200_balancecheck_init(ctsm_b200_ctx* ctx) {
  if (!ctx) return -1;
  // BalanceCheckMod.F90:91: skip_steps = max(2, nint(skip_size/dtime)) + 1, skip_size = 3600 s (:56)
  long n = lround(3600.0 / ctx->prm.dtime);
  if (n < 2) n = 2;
  ctx->prm.balance_skip_steps = (int)n + 1;
  return ctx->prm.balance_skip_steps;
}

__global__ void balance_init_kernel(Red* red) {
  const int k = threadIdx.x;
  if (k < CTSM_BAL_NKIND) { red->mx[k] = 0ULL; red->idx[k] = 0x7fffffff; }
  if (k == 0) red->pad = 0;
}

// thresholds and the abort decision of BalanceCheck / EnergyBalanceCheck on the maxima the kernels left (host side)
static int balance_decide(ctsm_b200_ctx* ctx, const Red& got, ctsm_balance_report_t* rep, int DAnstep, ctsm_status_t* st) {
  for (int k = 0; k < CTSM_BAL_NKIND; ++k) {
    long long b = (long long)got.mx[k];
    memcpy(&rep->max_abs[k], &b, sizeof(double));
    rep->index[k] = got.idx[k] == 0x7fffffff ? 0 : got.idx[k];
  }
  // thresholds: BalanceCheckMod.F90:65,506,901; order of the checks as in the reference
  const double error_thresh = 1.e-5, h2o_warn = 1.e-9, en_warn = 1.e-7;
  const int skip = ctx->prm.balance_skip_steps;
  int abort_kind = -1;
  const int worder[3] = {CTSM_BAL_H2O_COL, CTSM_BAL_H2O_GRC, CTSM_BAL_H2OSNO};
  for (int q = 0; q < 3; ++q) {
    const int k = worder[q];
    if (rep->max_abs[k] > h2o_warn) {
      rep->warn[k] = 1;
      if (rep->max_abs[k] > error_thresh && DAnstep > skip && abort_kind < 0) abort_kind = k;
    }
  }
  const int eorder[3] = {CTSM_BAL_SOL, CTSM_BAL_LON, CTSM_BAL_SEB};
  for (int q = 0; q < 3; ++q) {
    const int k = eorder[q];
    if (rep->max_abs[k] > en_warn && DAnstep > skip) {
      rep->warn[k] = 1;
      if (rep->max_abs[k] > error_thresh && abort_kind < 0) abort_kind = k;
    }
  }
  if (rep->max_abs[CTSM_BAL_SOI] > 1.0e-5) {
    rep->warn[CTSM_BAL_SOI] = 1;
    if (rep->max_abs[CTSM_BAL_SOI] > 1.e-4 && DAnstep > skip && abort_kind < 0) abort_kind = CTSM_BAL_SOI;
  }
  rep->abort_kind = abort_kind;
  if (abort_kind >= 0) {
    if (st) {
      st->code = CTSM_ERR_BALANCE;
      st->subgrid_index = rep->index[abort_kind];
      st->subgrid_level = (abort_kind == CTSM_BAL_H2O_GRC) ? CTSM_SUBGRID_GRIDCELL
                          : (abort_kind == CTSM_BAL_SOL || abort_kind == CTSM_BAL_LON || abort_kind == CTSM_BAL_SEB) ? CTSM_SUBGRID_PATCH
                                                                                                                       : CTSM_SUBGRID_COLUMN;
      st->info = abort_kind;
      st->value = rep->max_abs[abort_kind];
      static const char* what[CTSM_BAL_NKIND] = {"errh2o", "errh2o_grc", "errh2osno", "errsol", "errlon", "errseb", "errsoi_col"};
      snprintf(st->msg, sizeof st->msg, "BalanceCheck: CTSM is stopping because %s > threshold", what[abort_kind]);
    }
    return CTSM_ERR_BALANCE;
  }
  return CTSM_OK;
}

extern "C" int ctsm_b200_balancecheck(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_allc, const int32_t* filter_allc,
                                      const ctsm_balancecheck_fields_t* hf, int DAnstep, int mem, ctsm_balance_report_t* rep,
                                      ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || !rep || num_allc < 0) return CTSM_ERR_BAD_ARG;
  memset(rep, 0, sizeof *rep);
  rep->abort_kind = -1;
  rep->skip_steps = ctx->prm.balance_skip_steps;
  if (st) memset(st, 0, sizeof *st);
  if (ctx->prm.balance_skip_steps <= 0) {
    fprintf(stderr, "ctsm_b200_balancecheck called before ctsm_b200_balancecheck_init\n");   // GetBalanceCheckSkipSteps :117-128
    return CTSM_ERR_BAD_ARG;
  }
  (void)filter_allc;   // allc = every column in bounds (filterMod.F90 allc); the kernels loop over the bounds
  CUDA_TRY(cudaSetDevice(ctx->device));
  BalanceDev d;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_BALANCECHECK
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_BALANCECHECK
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
  }
  BGeo g;
  g.begc0 = hf->alloc.begc; g.begp0 = hf->alloc.begp; g.begg0 = hf->alloc.begg;
  g.ldc = hf->alloc.endc - hf->alloc.begc + 1; g.ldg = hf->alloc.endg - hf->alloc.begg + 1;
  g.begc = bounds->begc; g.endc = bounds->endc; g.begp = bounds->begp; g.endp = bounds->endp;
  g.begg = bounds->begg; g.endg = bounds->endg;
  const int nc = g.endc - g.begc + 1, ng = g.endg - g.begg + 1, np = g.endp - g.begp + 1;
  int rc = arena_reserve(ctx, ctx->arena_ints, sizeof(Red) + 64);
  if (rc) return rc;
  Red* red = (Red*)ctx->arena_ints.p;
  void* pinned_slot = nullptr;
  // Deferred decision: with device-resident state, and for host arrays inside a resident window, the call is
  // asynchronous; the warn / abort decision is taken when the host next synchronises (ctsm_b200_sync /
  // ctsm_b200_host_window_end), which is where the reference's endrun would surface to the caller anyway.
  const bool deferred = ctx->window_open || mem == CTSM_MEM_DEVICE;
  if (deferred) {
    // arena_ints is reused by the next call, so the maxima of this call get their own device + pinned slot (pooled)
    if (ctx->bal_pinned_used == ctx->bal_pinned.size()) {
      void *q = nullptr, *dq = nullptr;
      CUDA_TRY(cudaMallocHost(&q, sizeof(Red)));
      ctx->bal_pinned.push_back(q);
      CUDA_TRY(cudaMalloc(&dq, sizeof(Red)));
      ctx->bal_dev.push_back(dq);
    }
    red = (Red*)ctx->bal_dev[ctx->bal_pinned_used];
    pinned_slot = ctx->bal_pinned[ctx->bal_pinned_used++];
  }
  cudaStream_t s = ctx->stream;
  balance_init_kernel<<<1, 32, 0, s>>>(red);               // (no host->device copy: nothing on this stream may queue behind bulk DMA)
  const double dtime = ctx->prm.dtime;
  if (nc > 0) balance_col_kernel<<<grid_for(nc, 256), 256, 0, s>>>(d, g, dtime, nullptr, red);
  if (ng > 0) balance_grc_kernel<<<grid_for(ng, 256), 256, 0, s>>>(d, g, dtime, red);
  if (np > 0) balance_patch_kernel<<<grid_for(np, 256), 256, 0, s>>>(d, g, red);
  const int nmax = nc > np ? (nc > ng ? nc : ng) : (np > ng ? np : ng);
  if (nmax > 0) balance_loc_kernel<<<grid_for(nmax, 256), 256, 0, s>>>(d, g, red);
  ctx->launches += (nc > 0) + (ng > 0) + (np > 0) + (nmax > 0);
  if (deferred) {
    ctx->bal_last_dev = red;
    if (ctx->window_open) {
      // the maxima travel to the pinned slot on the download stream (a copy on the compute stream would queue behind
      // the bulk downloads in the copy engine and stall the next kernels)
      cudaEvent_t ev;
      rc = window_event(ctx, &ev);
      if (rc) return rc;
      CUDA_TRY(cudaEventRecord(ev, s));
      CUDA_TRY(cudaStreamWaitEvent(ctx->s_d2h, ev, 0));
      CUDA_TRY(cudaMemcpyAsync(pinned_slot, red, sizeof(Red), cudaMemcpyDeviceToHost, ctx->s_d2h));
    } else {
      CUDA_TRY(cudaMemcpyAsync(pinned_slot, red, sizeof(Red), cudaMemcpyDeviceToHost, s));
    }
    ctx->bal_pending.push_back(ctsm_b200_ctx::BalPending{pinned_slot, rep, DAnstep});
    if (mem != CTSM_MEM_DEVICE) {
      rc = stage_end(ctx, fl, hf->alloc, *bounds);
      if (rc) return rc;
    }
    return finish_call(ctx, mem, st);
  }
  Red got;
  CUDA_TRY(cudaMemcpyAsync(&got, red, sizeof got, cudaMemcpyDeviceToHost, s));
  if (mem != CTSM_MEM_DEVICE) {
    rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return balance_decide(ctx, got, rep, DAnstep, st);
}

int balance_finish_pending(ctsm_b200_ctx* ctx, ctsm_status_t* st) {
  int rc_all = CTSM_OK;
  for (const auto& pnd : ctx->bal_pending) {
    Red got;
    memcpy(&got, pnd.pinned, sizeof got);
    ctsm_status_t local;
    memset(&local, 0, sizeof local);
    const int rc = balance_decide(ctx, got, pnd.rep, pnd.DAnstep, &local);
    if (rc != CTSM_OK && rc_all == CTSM_OK) { rc_all = rc; if (st) *st = local; }
  }
  ctx->bal_pending.clear();
  ctx->bal_pinned_used = 0;
  return rc_all;
}

// device address of the 7 maxima (|residual| as doubles, order CTSM_BAL_*) the last deferred BalanceCheck call left
extern "C" void* ctsm_b200_balance_device_maxima(ctsm_b200_ctx* ctx) { return ctx ? ctx->bal_last_dev : nullptr; }

extern "C" int ctsm_b200_vert_tran_sink_hydstress(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_filterc,
                                                  const int32_t* filterc, const ctsm_plantsink_fields_t* hf, int mem,
                                                  ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_filterc < 0 || (num_filterc > 0 && !filterc)) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  PlantSinkDev d;
  const int32_t* dfilter = filterc;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_PLANTSINK
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_PLANTSINK
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filterc, num_filterc, &dfilter);
    if (rc) return rc;
  }
  if (num_filterc > 0) {
    if (ctx->tune.sink_warp)
      plantsink_warp_kernel<<<(num_filterc + SINK_WARPS - 1) / SINK_WARPS, SINK_WARPS * 32, 0, ctx->stream>>>(
          d, hf->alloc.begc, hf->alloc.endc - hf->alloc.begc + 1, hf->alloc.begp, hf->alloc.endp - hf->alloc.begp + 1,
          num_filterc, dfilter);
    else
      plantsink_kernel<<<grid_for(num_filterc, 128), 128, 0, ctx->stream>>>(
          d, hf->alloc.begc, hf->alloc.endc - hf->alloc.begc + 1, hf->alloc.begp, hf->alloc.endp - hf->alloc.begp + 1,
          num_filterc, dfilter);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

extern "C" int ctsm_b200_vert_tran_sink_default(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_filterc,
                                                  const int32_t* filterc, const ctsm_plantsinkdefault_fields_t* hf, int mem,
                                                  ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_filterc < 0 || (num_filterc > 0 && !filterc)) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  PlantSinkDefaultDev d;
  const int32_t* dfilter = filterc;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_PLANTSINKDEFAULT
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_PLANTSINKDEFAULT
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filterc, num_filterc, &dfilter);
    if (rc) return rc;
  }
  if (num_filterc > 0) {
    plantsink_default_kernel<<<grid_for(num_filterc, 128), 128, 0, ctx->stream>>>(
        d, hf->alloc.begc, hf->alloc.endc - hf->alloc.begc + 1, hf->alloc.begp, hf->alloc.endp - hf->alloc.begp + 1,
        num_filterc, dfilter);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}
