// soilwater.cu — SoilWater / soilwater_moisture_form as ONE fused kernel.
//
// Reference: src/biogeophys/SoilWaterMovementMod.F90:240 -> :976-1440 with
// compute_hydraulic_properties :1444, compute_moisture_fluxes_and_derivs :1560,
// compute_RHS/LHS_moisture_form :1871/:1941, IceImpedance :2139, the CH78
// retention curve (SoilWaterRetentionCurveClappHornberg1978Mod.F90:75-79,115-120)
// and the LAPACK dgtsv call at :1287.
//
// Mapping: the reference already loops column-outer (:1180), so one thread owns
// one hydrology column and runs the whole adaptive sub-step loop privately:
// per-layer coefficients that do not change between sub-steps (1000*dz,
// watsat, bsw, sucsat, sink, imped*hksat, 1000*(z(j+1)-z(j))) are loaded once
// (coalesced: column index is the fastest array index) and kept in per-thread
// arrays; the state h2osoi_liq is read once and written once.
// Roofline: HBM, ~2.4 KB/column-step; the sub-step loop costs
// 2 pow per layer per sub-step on the FP64 pipe.
// Two launches: 91 % of the columns accept the full time step at the first attempt while the rest need 3-20
// attempts, and a warp runs as long as its slowest column.  PASS 1 lets every column try the full step; a column whose
// first attempt is rejected (:1349) writes nothing and queues itself.  PASS 2 runs the complete adaptive loop for the
// queued columns only, compacted into full warps.  Same arithmetic per column, hence bit-identical results.
#include "solvers.cuh"

struct SoilWaterDev {   // device-side view of ctsm_soilwater_fields_t
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_SOILWATER
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SOILWATER
#undef CTSM_F
};

struct SoilWaterPrm {
  double dtime, dtmin, verySmall, xTolerUpper, xTolerLower, e_ice;
  int lower_bc, flux_calculation;
  // perturbed-parameter ensembles: e_ice per member, member of every column (NULL: the scalar above)
  const double* m_e_ice; const int32_t* col_member; int member_begc;
};

template <int PASS>
__global__ void __launch_bounds__(128)
soilwater_kernel(SoilWaterDev f, SoilWaterPrm prm, int begc0, int ldc, int numf_in, const int32_t* __restrict__ filter,
                 int32_t* __restrict__ retry, int* __restrict__ nretry, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  const int numf = (PASS == 1) ? numf_in : *nretry;
  if (fc >= numf) return;
  const int c1 = (PASS == 1) ? filter[fc] : retry[fc];          // 1-based proc-local column index
  const int ci = c1 - begc0;
  const size_t ld = (size_t)ldc;
  const int n = f.nbedrock[ci];       // nlayers, :1184
  const double e_ice = prm.m_e_ice ? prm.m_e_ice[prm.col_member[c1 - prm.member_begc]] : prm.e_ice;

  constexpr int N = NLEVSOI;
  double dz1000[N], watsat[N], bsw[N], sucsat[N], sink[N], ih[N], zden[N], liq[N];
  double s2[N], smp[N], dsmpdw[N], hk[N], qout[N];
  double amx[N], bmx[N], cmx[N], rmx[N];

  // ---- one-time loads (levels 1..n) -----------------------------------------
  {
    double icef_j = f.icefrac[ci];
    double z_j = f.z[(size_t)(1 - SNOSOI_LO) * ld + ci];
    for (int j = 0; j < n; ++j) {
      const size_t o1 = (size_t)j * ld + ci;                       // (c, j+1) of a (1:nlev) array
      const size_t os = (size_t)(j + 1 - SNOSOI_LO) * ld + ci;     // (c, j+1) of a (-nlevsno+1:) array
      dz1000[j] = cst::m_to_mm * f.dz[os];                         // == dz*denh2o bit for bit
      watsat[j] = f.watsat[o1];
      bsw[j] = f.bsw[o1];
      sucsat[j] = f.sucsat[o1];
      sink[j] = f.qflx_rootsoi[o1];
      liq[j] = f.h2osoi_liq[os];
      // IceImpedance :2158 with the interface ice fraction of :1521-1526
      double icef_n = 0.0, z_n = 0.0, imped;
      if (j < n - 1) {
        icef_n = f.icefrac[o1 + ld];
        z_n = f.z[os + ld];
        imped = pow(10.0, -e_ice * (0.5 * (icef_j + icef_n)));
        zden[j] = cst::m_to_mm * (z_n - z_j);                      // den of :1710
      } else {
        imped = pow(10.0, -e_ice * icef_j);
        zden[j] = 0.0;
      }
      ih[j] = imped * f.hksat[o1];                                 // soil_hk: imped*hksat*s**(2b+3)
      icef_j = icef_n; z_j = z_n;
    }
  }
  const double qflx_infl = f.qflx_infl[ci];

  int nsubstep = 0;                   // :1187
  double dtsub = prm.dtime;           // :1190
  double dtdone = 0.0;
  double qcharge = 0.0;               // :1194
  double dhkdw_bot = 0.0;

  for (;;) {                          // :1197
    nsubstep = nsubstep + 1;
    // :1203-1206, :1511-1516
    for (int j = 0; j < n; ++j) {
      const double vwc = fmax(liq[j], 1.0e-6) / dz1000[j];
      double s = vwc / watsat[j];
      s = fmin(s, 1.0);
      s = fmax(0.01, s);
      s2[j] = s;
      // soil_suction :115-120 and dsmpdw :1546
      const double sm = -sucsat[j] * pow(s, -bsw[j]);
      smp[j] = sm;
      dsmpdw[j] = (-bsw[j] * sm / s) / watsat[j];
    }
    // interface conductivities, fluxes, derivatives and matrix rows in one sweep
    double qin_j = qflx_infl;         // bc_flux :1677
    double dqidw0 = 0.0, dqidw1 = 0.0;
    for (int j = 0; j < n; ++j) {
      double s1 = (j == n - 1) ? s2[j] : 0.5 * (s2[j] + s2[j + 1]);   // :1520-1526
      s1 = fmin(s1, 1.0);
      s1 = fmax(0.01, s1);
      const double ex = 2.0 * bsw[j] + 3.0;
      const double hkj = ih[j] * pow(s1, ex);                          // soil_hk :75
      const double dhkds = ex * hkj / s1;                              // :79
      hk[j] = hkj;
      const double dt_dz = dtsub / dz1000[j];                          // :1205
      double qo, dqodw1, dqodw2;
      if (j < n - 1) {
        const double dhkds1 = 0.5 * dhkds / watsat[j];                 // :1697-1698
        const double dhkds2 = 0.5 * dhkds / watsat[j + 1];
        const double num = (smp[j + 1] - smp[j]);                      // :1709-1715
        const double den = zden[j];
        qo = -hkj * num / den + hkj;
        dqodw1 = (hkj * dsmpdw[j] - dhkds1 * num) / den + dhkds1;
        dqodw2 = (-hkj * dsmpdw[j + 1] - dhkds2 * num) / den + dhkds2;
      } else {
        dhkdw_bot = dhkds;
        if (prm.lower_bc == 1) {      // bc_flux :1840-1846
          qo = hkj;
          dqodw1 = dhkds / watsat[j];
        } else {                      // bc_zero_flux :1848-1854
          qo = 0.0;
          dqodw1 = 0.0;
        }
        dqodw2 = 0.0;
      }
      qout[j] = qo;
      // compute_RHS_moisture_form :1910-1926, compute_LHS_moisture_form :1984-2003
      const double fluxNet = qin_j - qo - sink[j];
      rmx[j] = -fluxNet * dt_dz;
      amx[j] = (j == 0) ? 0.0 : dqidw0 * dt_dz;
      bmx[j] = -1.0 - (-dqidw1 + dqodw1) * dt_dz;
      cmx[j] = (j == n - 1) ? 0.0 : -dqodw2 * dt_dz;
      // hand the interface to the layer below (:1722-1726)
      qin_j = qo; dqidw0 = dqodw1; dqidw1 = dqodw2;
    }
    // dgtsv :1279-1299 — dl(1:n-1) = amx(2:n): shift the sub-diagonal in place
    for (int j = 0; j < n - 1; ++j) amx[j] = amx[j + 1];
    const int info = dgtsv_solve(n, amx, bmx, cmx, rmx);
    if (info != 0) { report_failure(ds, c1, CTSM_ERR_DGTSV, info); return; }
    // rmx now holds dwat.  Error estimate :1307-1348 (flux_calculation = inexpensive)
    double errorMax = -INFINITY;
    {
      double qin_e = qflx_infl;
      for (int j = 0; j < n; ++j) {
        const double dt_dz = dtsub / dz1000[j];
        const double fluxNet0 = rmx[j] / dt_dz;
        const double fluxNet1 = qin_e - qout[j] - sink[j];
        const double e = fabs(fluxNet1 - fluxNet0) * dtsub * 0.5;
        errorMax = fmax(errorMax, e);
        qin_e = qout[j];
      }
    }
    if (errorMax > prm.xTolerUpper && dtsub > prm.dtmin) {   // :1349-1353
      if (PASS == 1) {                                       // needs sub-stepping: leave it to pass 2
        retry[atomicAdd(nretry, 1)] = c1;
        return;
      }
      dtsub = fmax(dtsub / 2.0, prm.dtmin);
      continue;
    }
    for (int j = 0; j < n; ++j) liq[j] = liq[j] + rmx[j] * dz1000[j];   // :1360-1362
    double qcTemp = 0.0;                                                // :1365-1383
    if (prm.lower_bc == 1) qcTemp = hk[n - 1] + dhkdw_bot * rmx[n - 1];
    qcharge = qcharge + qcTemp * (dtsub / prm.dtime);                   // :1386
    dtdone = dtdone + dtsub;                                            // :1389-1390
    if (fabs(prm.dtime - dtdone) < prm.verySmall) break;
    if (errorMax < prm.xTolerLower) dtsub = dtsub * 2.0;                // :1393-1395
    dtsub = fmin(dtsub, prm.dtime - dtdone);                            // :1398
  }

  // :1406-1410 over-saturation moves upward
  for (int j = n - 1; j >= 1; --j) {
    const double cap = f.eff_porosity[(size_t)j * ld + ci] * dz1000[j];   // eff_porosity*m_to_mm*dz
    const double over = fmax(liq[j] - cap, 0.0);
    liq[j] = fmin(cap, liq[j]);
    liq[j - 1] = liq[j - 1] + over;
  }
  // stores: state, diagnostics of the last sub-step (:1549-1550, :1419-1420)
  double qin_s = qflx_infl;
  for (int j = 0; j < n; ++j) {
    const size_t o1 = (size_t)j * ld + ci;
    f.h2osoi_liq[(size_t)(j + 1 - SNOSOI_LO) * ld + ci] = liq[j];
    f.smp_l[o1] = smp[j];
    f.hk_l[o1] = hk[j];
    f.qin[o1] = qin_s;
    f.qout[o1] = qout[j];
    qin_s = qout[j];
  }
  f.qcharge[ci] = qcharge;
  f.num_substeps[ci] = (double)nsubstep;
}

// PASS 2 with one WARP per queued column, lane j = soil level j + 1.  In soilwater_kernel<2> a thread walks its column's
// sub-steps alone (3-40 attempts of 20 levels x 3 pow each): the launch lasts as long as the slowest column's serial
// chain (2.4 ms at f02 for 9 % of the columns, 8 of 32 lanes busy).  Per sub-step everything except the pivoted
// tridiagonal solve is independent per level, so the levels go across the lanes: retention curve, conductivities, fluxes,
// derivatives and matrix rows per lane with the level below / above fetched by shuffle; the rows go to shared memory,
// lane 0 runs the same dgtsv_solve, the error estimate is a warp maximum (exact).  Same expressions per level, same
// operation order inside the solve and the over-saturation sweep: bit-identical to the one-thread form
// (tests/test_gpu_soil.py::test_soilwater_retry_kernels_agree_bit_for_bit).
#define SW_WARPS 4
__global__ void __launch_bounds__(32 * SW_WARPS)
soilwater_retry_warp_kernel(SoilWaterDev f, SoilWaterPrm prm, int begc0, int ldc, const int32_t* __restrict__ retry,
                            const int* __restrict__ nretry, DevStatus* ds) {
  constexpr int N = NLEVSOI;
  __shared__ double sh[SW_WARPS][4][N];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* dl = sh[wib][0];
  double* dd = sh[wib][1];
  double* du = sh[wib][2];
  double* bb = sh[wib][3];
  const unsigned FULLM = 0xffffffffu;
  const int numf = *nretry;
  const int nwarps = gridDim.x * SW_WARPS;
  const size_t ld = (size_t)ldc;
  for (int w = blockIdx.x * SW_WARPS + wib; w < numf; w += nwarps) {
    const int c1 = retry[w];
    const int ci = c1 - begc0;
    const int n = f.nbedrock[ci];
    const double e_ice = prm.m_e_ice ? prm.m_e_ice[prm.col_member[c1 - prm.member_begc]] : prm.e_ice;
    const int j = lane;
    const bool on = j < n;
    // ---- one-time loads of this lane's level (inactive lanes carry harmless values) ----
    double dz1000 = 1.0, watsat = 1.0, bsw = 1.0, sucsat = 1.0, sink = 0.0, liq = 1.0, ih = 0.0, zden = 1.0;
    if (on) {
      const size_t o1 = (size_t)j * ld + ci;
      const size_t os = (size_t)(j + 1 - SNOSOI_LO) * ld + ci;
      dz1000 = cst::m_to_mm * f.dz[os];
      watsat = f.watsat[o1];
      bsw = f.bsw[o1];
      sucsat = f.sucsat[o1];
      sink = f.qflx_rootsoi[o1];
      liq = f.h2osoi_liq[os];
      const double icef_j = f.icefrac[o1], z_j = f.z[os];
      double imped;
      if (j < n - 1) {
        const double icef_n = f.icefrac[o1 + ld], z_n = f.z[os + ld];
        imped = pow(10.0, -e_ice * (0.5 * (icef_j + icef_n)));
        zden = cst::m_to_mm * (z_n - z_j);
      } else {
        imped = pow(10.0, -e_ice * icef_j);
        zden = 0.0;
      }
      ih = imped * f.hksat[o1];
    }
    const double watsat_n = __shfl_down_sync(FULLM, watsat, 1);
    const double qflx_infl = f.qflx_infl[ci];
    int nsubstep = 0;
    double dtsub = prm.dtime, dtdone = 0.0, qcharge = 0.0;
    double smp = 0.0, hk = 0.0, qout = 0.0, qin = 0.0;
    bool failed = false;
    for (;;) {
      nsubstep = nsubstep + 1;
      const double vwc = fmax(liq, 1.0e-6) / dz1000;
      double s2 = vwc / watsat;
      s2 = fmin(s2, 1.0);
      s2 = fmax(0.01, s2);
      const double sm = -sucsat * pow(s2, -bsw);
      smp = sm;
      const double dsmpdw = (-bsw * sm / s2) / watsat;
      const double s2_n = __shfl_down_sync(FULLM, s2, 1);
      const double smp_n = __shfl_down_sync(FULLM, sm, 1);
      const double dsmpdw_n = __shfl_down_sync(FULLM, dsmpdw, 1);
      double s1 = (j == n - 1) ? s2 : 0.5 * (s2 + s2_n);
      s1 = fmin(s1, 1.0);
      s1 = fmax(0.01, s1);
      const double ex = 2.0 * bsw + 3.0;
      const double hkj = ih * pow(s1, ex);
      const double dhkds = ex * hkj / s1;
      hk = hkj;
      const double dt_dz = dtsub / dz1000;
      double qo, dqodw1, dqodw2;
      if (j < n - 1) {
        const double dhkds1 = 0.5 * dhkds / watsat;
        const double dhkds2 = 0.5 * dhkds / watsat_n;
        const double num = (smp_n - sm);
        const double den = zden;
        qo = -hkj * num / den + hkj;
        dqodw1 = (hkj * dsmpdw - dhkds1 * num) / den + dhkds1;
        dqodw2 = (-hkj * dsmpdw_n - dhkds2 * num) / den + dhkds2;
      } else {
        if (prm.lower_bc == 1) { qo = hkj; dqodw1 = dhkds / watsat; }
        else { qo = 0.0; dqodw1 = 0.0; }
        dqodw2 = 0.0;
      }
      qout = qo;
      double qin_j = __shfl_up_sync(FULLM, qo, 1);
      double dqidw0 = __shfl_up_sync(FULLM, dqodw1, 1);
      double dqidw1 = __shfl_up_sync(FULLM, dqodw2, 1);
      if (j == 0) { qin_j = qflx_infl; dqidw0 = 0.0; dqidw1 = 0.0; }
      qin = qin_j;
      const double fluxNet = qin_j - qo - sink;
      const double rmx = -fluxNet * dt_dz;
      const double amx = (j == 0) ? 0.0 : dqidw0 * dt_dz;
      const double bmx = -1.0 - (-dqidw1 + dqodw1) * dt_dz;
      const double cmx = (j == n - 1) ? 0.0 : -dqodw2 * dt_dz;
      __syncwarp();
      if (on) {
        if (j >= 1) dl[j - 1] = amx;          // dgtsv: dl(1:n-1) = amx(2:n)
        dd[j] = bmx; du[j] = cmx; bb[j] = rmx;
      }
      __syncwarp();
      int info = 0;
      if (lane == 0) info = dgtsv_solve(n, dl, dd, du, bb);
      info = __shfl_sync(FULLM, info, 0);
      __syncwarp();
      if (info != 0) {
        if (lane == 0) report_failure(ds, c1, CTSM_ERR_DGTSV, info);
        failed = true;
        break;
      }
      const double dwat = on ? bb[j] : 0.0;
      // error estimate :1307-1348
      double e = -INFINITY;
      if (on) {
        const double fluxNet0 = dwat / dt_dz;
        const double fluxNet1 = qin_j - qo - sink;
        e = fabs(fluxNet1 - fluxNet0) * dtsub * 0.5;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(FULLM, e, o));
      const double errorMax = e;
      if (errorMax > prm.xTolerUpper && dtsub > prm.dtmin) {   // :1349-1353
        dtsub = fmax(dtsub / 2.0, prm.dtmin);
        continue;
      }
      if (on) liq = liq + dwat * dz1000;                         // :1360-1362
      double qcTemp = 0.0;                                       // :1365-1383
      {
        const double hk_b = __shfl_sync(FULLM, hkj, n - 1);
        const double dh_b = __shfl_sync(FULLM, dhkds, n - 1);
        const double dw_b = __shfl_sync(FULLM, dwat, n - 1);
        if (prm.lower_bc == 1) qcTemp = hk_b + dh_b * dw_b;
      }
      qcharge = qcharge + qcTemp * (dtsub / prm.dtime);
      dtdone = dtdone + dtsub;
      if (fabs(prm.dtime - dtdone) < prm.verySmall) break;
      if (errorMax < prm.xTolerLower) dtsub = dtsub * 2.0;
      dtsub = fmin(dtsub, prm.dtime - dtdone);
    }
    if (failed) continue;
    // :1406-1410 over-saturation moves upward (sequential recurrence: lane 0)
    __syncwarp();
    if (on) { dd[j] = liq; du[j] = (j >= 1) ? f.eff_porosity[(size_t)j * ld + ci] * dz1000 : 0.0; }
    __syncwarp();
    if (lane == 0) {
      for (int q = n - 1; q >= 1; --q) {
        const double cap = du[q];
        const double over = fmax(dd[q] - cap, 0.0);
        dd[q] = fmin(cap, dd[q]);
        dd[q - 1] = dd[q - 1] + over;
      }
    }
    __syncwarp();
    if (on) {
      const size_t o1 = (size_t)j * ld + ci;
      f.h2osoi_liq[(size_t)(j + 1 - SNOSOI_LO) * ld + ci] = dd[j];
      f.smp_l[o1] = smp;
      f.hk_l[o1] = hk;
      f.qin[o1] = qin;
      f.qout[o1] = qout;
    }
    if (lane == 0) {
      f.qcharge[ci] = qcharge;
      f.num_substeps[ci] = (double)nsubstep;
    }
    __syncwarp();
  }
}

extern "C" int ctsm_b200_soilwater(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_hydrologyc,
                                   const int32_t* filter_hydrologyc, const ctsm_soilwater_fields_t* hf, int mem,
                                   ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || (num_hydrologyc > 0 && !filter_hydrologyc) || num_hydrologyc < 0)
    return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  SoilWaterDev d;
  const int32_t* dfilter = filter_hydrologyc;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_SOILWATER
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SOILWATER
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_hydrologyc, num_hydrologyc, &dfilter);
    if (rc) return rc;
  }
  SoilWaterPrm p{ctx->prm.dtime, ctx->prm.dtmin, ctx->prm.verySmall, ctx->prm.xTolerUpper, ctx->prm.xTolerLower,
                 ctx->prm.e_ice, ctx->prm.lower_boundary_condition, ctx->prm.flux_calculation,
                 ctx->member.e_ice, ctx->member.col_member, ctx->member.begc};
  if (p.m_e_ice && (bounds->begc < ctx->member.begc || bounds->endc > ctx->member.endc)) return CTSM_ERR_BAD_ARG;
  if (p.flux_calculation != 1) return CTSM_ERR_BAD_ARG;
  if (num_hydrologyc > 0) {
    int rc = arena_reserve(ctx, ctx->arena_ints, sizeof(int32_t) * ((size_t)num_hydrologyc + 64));
    if (rc) return rc;
    int* nretry = (int*)ctx->arena_ints.p;
    int32_t* retry = (int32_t*)ctx->arena_ints.p + 64;
    CUDA_TRY(cudaMemsetAsync(nretry, 0, sizeof(int), ctx->stream));
    const int ldc = hf->alloc.endc - hf->alloc.begc + 1;
    soilwater_kernel<1><<<grid_for(num_hydrologyc, 128), 128, 0, ctx->stream>>>(d, p, hf->alloc.begc, ldc, num_hydrologyc,
                                                                                dfilter, retry, nretry, ctx->d_status);
    // the queue length is only known on the device: pass 2 is sized for the whole filter and exits early
    if (ctx->tune.sw_warp) {
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
      const int need = grid_for(num_hydrologyc, SW_WARPS);
      soilwater_retry_warp_kernel<<<need < sms * 8 ? need : sms * 8, 32 * SW_WARPS, 0, ctx->stream>>>(d, p, hf->alloc.begc, ldc, retry,
                                                                                                   nretry, ctx->d_status);
    } else {
      soilwater_kernel<2><<<grid_for(num_hydrologyc, 128), 128, 0, ctx->stream>>>(d, p, hf->alloc.begc, ldc, num_hydrologyc,
                                                                                  dfilter, retry, nretry, ctx->d_status);
    }
    ctx->launches += 2;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}
