// soiltemp.cu — SoilTemperature as one fused column kernel (+ a patch-mask helper).
//
// Reference: src/biogeophys/SoilTemperatureMod.F90
//   SoilTemperature :92-599        SoilThermProp :602-901
//   PhaseChangeH2osfc :904-1130    Phasechange :1133-1540
//   ComputeGroundHeatFluxAndDeriv :1543-1796
//   ComputeHeatDiffFluxAndFactor :1799-1910
//   SetRHSVec* :1913-2353          SetMatrix* :2356-2926
//   BandDiagonal -> dgbsv (BandDiagonalMod.F90:167-219)
// Non-urban columns only (istsoil, istcrop, istice, istwet); use_excess_ice=.false.
//
// Mapping: one thread owns one column.  The ~15 level x column loop nests of the
// reference collapse into sweeps over the column's own levels; the 5-band
// matrix (7 zero-filled (c,5,lev) temporaries on the CPU) is never materialised:
// rows are produced on demand for the streaming pivoted LU (solvers.cuh).
// The patch -> column sums of ComputeGroundHeatFluxAndDeriv run serially over
// the column's contiguous patches in ascending index order, which is the
// filter order of the reference (summation order matters for <=1e-10 parity;
// SURVEY.md H7) — no atomics.
// Roofline: HBM, ~9.1 KB per column-step with 15 patches.
#include "solvers.cuh"

struct SoilTempDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_SOILTEMPERATURE
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SOILTEMPERATURE
#undef CTSM_F
};

struct SoilTempPrm {
  double dtime;
  int snow_method, snow_glc_method;
};

__global__ void patchmask_kernel(int32_t* __restrict__ mask, int begp0, int nump, const int32_t* __restrict__ filterp) {
  const int fp = blockIdx.x * blockDim.x + threadIdx.x;
  if (fp < nump) mask[filterp[fp] - begp0] = 1;
}

__device__ __forceinline__ double snow_thk(int method, double bw) {
  // SoilTemperatureMod.F90:750-781
  if (method == 1) return cst::tkair + (7.75e-5 * bw + 1.105e-6 * bw * bw) * (cst::tkice - cst::tkair);
  if (bw <= 156.0) return R4(0.023) + R4(0.234) * (bw / 1000.0);
  return R4(0.138) - R4(1.01) * (bw / 1000.0) + (R4(3.233) * ((bw / 1000.0) * (bw / 1000.0)));
}

__device__ __forceinline__ double pow4(double t) { const double t2 = t * t; return t2 * t2; }
__device__ __forceinline__ double pow3(double t) { return t * t * t; }

#define NL (NLEVSNO + NLEVGRND)        // 37 layers, k = j + NLEVSNO - 1
#define KOF(j) ((j) + NLEVSNO - 1)

__global__ void __launch_bounds__(128)
soiltemp_kernel(SoilTempDev f, SoilTempPrm prm, int begc0, int ldc_, int begp0, int ldp_, int numc,
                const int32_t* __restrict__ filterc, const int32_t* __restrict__ patchmask, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numc) return;
  const int c1 = filterc[fc];
  const int ci = c1 - begc0;
  const size_t ldc = (size_t)ldc_, ldp = (size_t)ldp_;
  const double dtime = prm.dtime;
  using namespace cst;

  const int lt = f.lun_itype[ci];
  if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
  const bool soilcrop = (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP);
  const int snl = f.snl[ci];
  const int jtopl = snl + 1;                      // top active layer
  const int nbed = f.nbedrock[ci];
  const double frac_sno_eff = f.frac_sno_eff[ci];
  const double frac_h2osfc = f.frac_h2osfc[ci];
  double h2osfc = f.h2osfc[ci];
  double h2osno_no_layers = f.h2osno_no_layers[ci];
  double t_h2osfc = f.t_h2osfc[ci];
  const double t_grnd_old = f.t_grnd[ci];

  // (c, j) offsets
#define OS(j) ((size_t)((j) - SNOSOI_LO) * ldc + ci)     /* (-nlevsno+1:nlevgrnd) arrays */
#define OZ(j) ((size_t)((j) - SNOSOI0_LO) * ldc + ci)    /* zi (-nlevsno:nlevgrnd)       */
#define OG(j) ((size_t)((j) - 1) * ldc + ci)             /* (1:nlevgrnd) arrays          */

  double t[NL], tk[NL], fn[NL], dzp[NL];   // dzp(j) = z(j+1)-z(j)

  // ---- SoilThermProp :602-901 + heat capacity; fact/fn of :1799-1910 -------------
  // cv is consumed immediately (fact = dtime/cv); thk goes to its output array.
  double thk_prev = 0.0, z_prev = 0.0, zi_prev = 0.0, thk1 = 0.0, z1 = 0.0;
  for (int j = jtopl; j <= NLEVGRND; ++j) {
    const int k = KOF(j);
    const double tj = f.t_soisno[OS(j)];
    const double liq = f.h2osoi_liq[OS(j)];
    const double ice = f.h2osoi_ice[OS(j)];
    const double dzj = f.dz[OS(j)];
    const double zj = f.z[OS(j)];
    const double zij = f.zi[OZ(j)];
    t[k] = tj;
    double thk, cv;
    if (j >= 1) {
      if (lt != CTSM_ISTWET && lt != CTSM_ISTICE) {                          // :707-726
        const double watsat = f.watsat[OG(j)];
        double satw = (liq / denh2o + ice / denice) / (dzj * watsat);
        satw = fmin(1.0, satw);
        if (satw > .1e-6) {
          double dke;
          if (tj >= tfrz) dke = fmax(0.0, log10(satw) + 1.0);
          else dke = satw;
          const double fl = (liq / (denh2o * dzj)) / (liq / (denh2o * dzj) + ice / (denice * dzj));
          const double dksat = f.tkmg[OG(j)] * pow(tkwat, fl * watsat) * pow(tkice, (1.0 - fl) * watsat);
          thk = dke * dksat + (1.0 - dke) * f.tkdry[OG(j)];
        } else {
          thk = f.tkdry[OG(j)];
        }
        if (j > nbed) thk = thk_bedrock;
        cv = f.csol[OG(j)] * (1.0 - watsat) * dzj + (ice * cpice + liq * cpliq);   // :847-849
        if (j > nbed) cv = csol_bedrock * dzj;
      } else if (lt == CTSM_ISTICE) {                                        // :727-729, :853-854
        thk = tkwat;
        if (tj < tfrz) thk = tkice;
        cv = (ice * cpice + liq * cpliq);
      } else {                                                               // istwet :730-737, :850-852
        if (j > NLEVSOI) thk = thk_bedrock;
        else { thk = tkwat; if (tj < tfrz) thk = tkice; }
        cv = (ice * cpice + liq * cpliq);
        if (j > nbed) cv = csol_bedrock * dzj;
      }
      if (j == 1 && h2osno_no_layers > 0.0) cv = cv + cpice * h2osno_no_layers;    // :875-880
    } else {                                                                 // snow :741-781, :884-895
      const double bw = (ice + liq) / (frac_sno_eff * dzj);
      f.bw[OS(j)] = bw;
      thk = snow_thk(lt == CTSM_ISTICE ? prm.snow_glc_method : prm.snow_method, bw);
      if (frac_sno_eff > 0.0) cv = fmax(thin_sfclayer, (cpliq * liq + cpice * ice) / frac_sno_eff);
      else cv = thin_sfclayer;
    }
    f.thk[OS(j)] = thk;
    if (j == 1) { thk1 = thk; z1 = zj; }
    // interface conductivity of the layer above, :803-826
    if (j > jtopl) {
      const int km = k - 1;
      dzp[km] = zj - z_prev;
      tk[km] = thk_prev * thk * (zj - z_prev) / (thk_prev * (zj - zi_prev) + thk * (zi_prev - z_prev));
    }
    // fact, :1887-1902 (the top-layer form needs z(j+1); patched after the loop)
    fn[k] = cv;   // park cv here until fact is formed
    thk_prev = thk; z_prev = zj; zi_prev = zij;
  }
  tk[KOF(NLEVGRND)] = 0.0;
  dzp[KOF(NLEVGRND)] = 0.0;
  const double eflx_bot = f.eflx_bot[ci];
  double factv[NL];
  for (int j = jtopl; j <= NLEVGRND; ++j) {
    const int k = KOF(j);
    const double cv = fn[k];
    double fa;
    if (j == jtopl) {
      const double zj = f.z[OS(j)], zim = f.zi[OZ(j - 1)], zjp = f.z[OS(j + 1)], dzj = f.dz[OS(j)];
      fa = dtime / cv * dzj / (0.5 * (zj - zim + capr * (zjp - zim)));
    } else {
      fa = dtime / cv;
    }
    factv[k] = fa;
    f.fact[OS(j)] = fa;
  }
  for (int j = jtopl; j <= NLEVGRND; ++j) {
    const int k = KOF(j);
    if (j <= NLEVGRND - 1) fn[k] = tk[k] * (t[k + 1] - t[k]) / dzp[k];
    else fn[k] = eflx_bot;
  }
  // thermal conductivity of h2osfc :829-835
  const double zh2osfc = R4(1.0e-3) * (0.5 * h2osfc);
  const double tk_h2osfc = tkwat * thk1 * (z1 + zh2osfc) / (tkwat * z1 + thk1 * zh2osfc);

  // ---- ComputeGroundHeatFluxAndDeriv :1543-1796 -----------------------------------
  const double emg = f.emg[ci], forc_lwrad = f.forc_lwrad[ci], htvp = f.htvp[ci];
  const double lwrad_emit = emg * sb * pow4(t_grnd_old);
  const double dlwrad_emit = 4.0 * emg * sb * pow3(t_grnd_old);
  const double lwrad_emit_snow = emg * sb * pow4(t[KOF(jtopl)]);
  const double lwrad_emit_soil = emg * sb * pow4(t[KOF(1)]);
  const double lwrad_emit_h2osfc = emg * sb * pow4(t_h2osfc);
  double hs_soil = 0.0, hs_h2osfc = 0.0, dhsdT = 0.0, hs_top = 0.0, hs_top_snow = 0.0;
  double sabg_lyr_col[NLEVSNO + 1];   // j = -nlevsno+1 .. 1
#pragma unroll
  for (int q = 0; q < NLEVSNO + 1; ++q) sabg_lyr_col[q] = 0.0;
  {
    const int pi = f.patchi[ci], pf = f.patchf[ci];
    for (int p1 = pi; p1 <= pf; ++p1) {
      const int pp = p1 - begp0;
      if (!patchmask[pp]) continue;
      const double wt = f.wtcol[pp];
      const double fv = (double)f.frac_veg_nosno[pp];
      const double dlrad = f.dlrad[pp];
      const double lwdn = (1.0 - fv) * emg * forc_lwrad;
      const double sabg_soil = f.sabg_soil[pp];
      const double eflx_gnet = f.sabg[pp] + dlrad + lwdn - lwrad_emit
                               - (f.eflx_sh_grnd[pp] + f.qflx_evap_soi[pp] * htvp);
      f.eflx_gnet[pp] = eflx_gnet;
      f.sabg_chk[pp] = frac_sno_eff * f.sabg_snow[pp] + (1.0 - frac_sno_eff) * sabg_soil;
      const double sh_soil = f.eflx_sh_soil[pp], ev_soil = f.qflx_ev_soil[pp];
      const double sh_snow = f.eflx_sh_snow[pp], ev_snow = f.qflx_ev_snow[pp];
      const double eflx_gnet_soil = sabg_soil + dlrad + lwdn - lwrad_emit_soil - (sh_soil + ev_soil * htvp);
      const double eflx_gnet_h2osfc = sabg_soil + dlrad + lwdn - lwrad_emit_h2osfc
                                      - (f.eflx_sh_h2osfc[pp] + f.qflx_ev_h2osfc[pp] * htvp);
      const double dgnetdT = -f.cgrnd[pp] - dlwrad_emit;
      f.dgnetdT[pp] = dgnetdT;
      dhsdT = dhsdT + dgnetdT * wt;
      hs_soil = hs_soil + eflx_gnet_soil * wt;
      hs_h2osfc = hs_h2osfc + eflx_gnet_h2osfc * wt;
      // second patch loop of the reference (:1760-1792): separate accumulators, same patch order
      const double sabg_top = f.sabg_lyr[(size_t)(jtopl - SNOSOI_LO) * ldp + pp];
      const double eflx_gnet_top = sabg_top + dlrad + lwdn - lwrad_emit - (f.eflx_sh_grnd[pp] + f.qflx_evap_soi[pp] * htvp);
      hs_top = hs_top + eflx_gnet_top * wt;
      const double eflx_gnet_snow = sabg_top + dlrad + lwdn - lwrad_emit_snow - (sh_snow + ev_snow * htvp);
      hs_top_snow = hs_top_snow + eflx_gnet_snow * wt;
      for (int j = jtopl; j <= 1; ++j) {
        const int q = j - SNOSOI_LO;
        sabg_lyr_col[q] = sabg_lyr_col[q] + f.sabg_lyr[(size_t)q * ldp + pp] * wt;
      }
    }
  }
  (void)hs_top;   // only urban non-road columns consume hs_top (:2106-2112)

  // ---- h2osfc thermal properties :348-357 --------------------------------------------
  double c_h2osfc, dz_h2osfc;
  if ((h2osfc > thin_sfclayer) && (frac_h2osfc > thin_sfclayer)) {
    c_h2osfc = fmax(thin_sfclayer, cpliq * h2osfc / frac_h2osfc);
    dz_h2osfc = fmax(thin_sfclayer, R4(1.0e-3) * h2osfc / frac_h2osfc);
  } else {
    c_h2osfc = thin_sfclayer;
    dz_h2osfc = thin_sfclayer;
  }
  f.c_h2osfc[ci] = c_h2osfc;
  const double dzm_ssw = (0.5 * dz_h2osfc + z1);                                  // :2199, :2799, :2905
  const double fn_h2osfc = tk_h2osfc * (t[KOF(1)] - t_h2osfc) / dzm_ssw;          // :2201

  // ---- SetRHSVec / SetMatrix rows on demand, BandDiagonal -> dgbsv ---------------------
  // matrix row r in [snl, nlevgrnd]: r <= -1 snow layer j = r+1; r = 0 standing water; r >= 1 soil
  double U[NL + 1][5], y[NL + 1];
  const int n = NLEVGRND - snl + 1;
  auto row = [&](int i, double* e) -> double {
    const int r = snl + i;
    e[0] = 0.0; e[1] = 0.0; e[2] = 0.0; e[3] = 0.0; e[4] = 0.0;
    double rhs;
    if (r <= -1) {                                   // SetRHSVec_Snow :2127-2144, SetMatrix_Snow :2608-2634
      const int j = r + 1, k = KOF(j);
      const double fa = factv[k], tj = t[k];
      if (j == jtopl) {
        rhs = tj + fa * (hs_top_snow - dhsdT * tj + cnfac * fn[k]);
        e[1] = 0.0;
        e[2] = 1.0 + (1.0 - cnfac) * fa * tk[k] / dzp[k] - fa * dhsdT;
      } else {
        rhs = tj + cnfac * fa * (fn[k] - fn[k - 1]);
        rhs = rhs + fa * sabg_lyr_col[j - SNOSOI_LO];
        e[1] = -(1.0 - cnfac) * fa * tk[k - 1] / dzp[k - 1];
        e[2] = 1.0 + (1.0 - cnfac) * fa * (tk[k] / dzp[k] + tk[k - 1] / dzp[k - 1]);
      }
      const double sup = -(1.0 - cnfac) * fa * tk[k] / dzp[k];
      if (j != 0) e[3] = sup;                        // band 2
      else e[4] = sup;                               // band 1: snow layer 0 couples to soil row 1
    } else if (r == 0) {                             // :2199-2204, :2905-2911
      rhs = t_h2osfc + (dtime / c_h2osfc) * (hs_h2osfc - dhsdT * t_h2osfc + cnfac * fn_h2osfc);
      e[2] = 1.0 + (1.0 - cnfac) * (dtime / c_h2osfc) * tk_h2osfc / dzm_ssw - (dtime / c_h2osfc) * dhsdT;
      e[3] = -(1.0 - cnfac) * (dtime / c_h2osfc) * tk_h2osfc / dzm_ssw;
    } else {                                         // SetRHSVec_Soil :2307-2349, SetMatrix_Soil :2746-2807
      const int j = r, k = KOF(j);
      const double fa = factv[k], tj = t[k];
      if (j == jtopl) {                              // j == 1, no snow layers
        rhs = tj + fa * (hs_top_snow - dhsdT * tj + cnfac * fn[k]);
        e[0] = 0.0;
        e[2] = 1.0 + (1.0 - cnfac) * fa * tk[k] / dzp[k] - fa * dhsdT;
        e[3] = -(1.0 - cnfac) * fa * tk[k] / dzp[k];
      } else if (j == 1) {                           // snow/soil interface layer
        rhs = tj + fa * ((1.0 - frac_sno_eff) * (hs_soil - dhsdT * tj) + cnfac * (fn[k] - frac_sno_eff * fn[k - 1]));
        rhs = rhs + frac_sno_eff * fa * sabg_lyr_col[1 - SNOSOI_LO];
        e[3] = -(1.0 - cnfac) * fa * tk[k] / dzp[k];
        e[2] = 1.0 + (1.0 - cnfac) * fa * (tk[k] / dzp[k] + frac_sno_eff * tk[k - 1] / dzp[k - 1])
               - (1.0 - frac_sno_eff) * fa * dhsdT;
        e[0] = -frac_sno_eff * (1.0 - cnfac) * fa * tk[k - 1] / dzp[k - 1];   // band 5: column -1 (snow layer 0)
      } else if (j <= NLEVGRND - 1) {
        rhs = tj + cnfac * fa * (fn[k] - fn[k - 1]);
        e[3] = -(1.0 - cnfac) * fa * tk[k] / dzp[k];
        e[2] = 1.0 + (1.0 - cnfac) * fa * (tk[k] / dzp[k] + tk[k - 1] / dzp[k - 1]);
        e[1] = -(1.0 - cnfac) * fa * tk[k - 1] / dzp[k - 1];
      } else {
        rhs = tj - cnfac * fa * fn[k - 1] + fa * fn[k];
        e[3] = 0.0;
        e[2] = 1.0 + (1.0 - cnfac) * fa * tk[k - 1] / dzp[k - 1];
        e[1] = -(1.0 - cnfac) * fa * tk[k - 1] / dzp[k - 1];
      }
      if (j == 1 && frac_h2osfc != 0.0) {            // :2342-2349, :2797-2807, :2918-2921
        rhs = rhs - frac_h2osfc * fa * ((hs_soil - dhsdT * tj) + cnfac * fn_h2osfc);
        e[2] = e[2] + frac_h2osfc * ((1.0 - cnfac) * fa * tk_h2osfc / dzm_ssw + fa * dhsdT);
        e[1] = -frac_h2osfc * (1.0 - cnfac) * fa * tk_h2osfc / dzm_ssw;       // band 4: column 0 (standing water)
      }
    }
    return rhs;
  };
  const int info = band5_solve<NL + 1>(n, row, U, y);
  if (info != 0) { report_failure(ds, c1, CTSM_ERR_DGBSV, info); return; }

  // ---- unpack :422-434, fn1 :438-483, eflx_fgr :580-595 ---------------------------------
  for (int j = jtopl; j <= 0; ++j) t[KOF(j)] = y[j - 1 - snl];
  for (int j = 1; j <= NLEVGRND; ++j) t[KOF(j)] = y[j - snl];
  if (frac_h2osfc == 0.0) t_h2osfc = t[KOF(1)];
  else t_h2osfc = y[0 - snl];
  double fn1_1 = 0.0;
  for (int j = 1; j <= NLEVGRND; ++j) {
    const int k = KOF(j);
    double fn1;
    if (j <= NLEVGRND - 1) fn1 = tk[k] * (t[k + 1] - t[k]) / dzp[k];
    else fn1 = 0.0;
    if (j == 1) fn1_1 = fn1;
    if (soilcrop) {
      if (j < NLEVGRND) f.eflx_fgr[OG(j)] = -cnfac * fn[k] - (1.0 - cnfac) * fn1;
      else f.eflx_fgr[OG(j)] = 0.0;
    }
  }
  const double eflx_fgr12 = -cnfac * fn[KOF(1)] - (1.0 - cnfac) * fn1_1;
  f.eflx_fgr12[ci] = eflx_fgr12;

  // ---- PhaseChangeH2osfc :904-1130 --------------------------------------------------------
  double xmf_h2osfc = 0.0, qflx_h2osfc_to_ice = 0.0, eflx_h2osfc_to_snow = 0.0;
  double snow_depth = f.snow_depth[ci];
  double ice0_delta = 0.0;      // change applied to h2osoi_ice(c,0); folded in during the Phasechange sweep
  bool t0_written = false;      // t_soisno(c,0) written while snl == 0 (":1039 initialize for next time step")
  double t0_value = 0.0;
  if (frac_h2osfc > 0.0 && t_h2osfc <= tfrz) {
    double h2osno_total = h2osno_no_layers;                                   // CalculateTotalH2osno
    for (int j = jtopl; j <= 0; ++j) h2osno_total = h2osno_total + f.h2osoi_ice[OS(j)] + f.h2osoi_liq[OS(j)];
    double int_snow = f.int_snow[ci];
    const double tinc = tfrz - t_h2osfc;
    t_h2osfc = tfrz;
    const double hm = frac_h2osfc * (dhsdT * tinc - tinc * c_h2osfc / dtime);
    const double xm = hm * dtime / hfus;
    const double temp1 = h2osfc + xm;
    const double z_avg = frac_sno_eff * snow_depth;
    double rho_avg;
    if (z_avg > 0.0) rho_avg = fmin(800.0, h2osno_total / z_avg);
    else rho_avg = 200.0;
    const double fact0 = (snl < 0) ? factv[KOF(0)] : 0.0;
    if (temp1 >= 0.0) {
      int_snow = int_snow - xm;
      if (snl == 0) h2osno_no_layers = h2osno_no_layers - xm;
      else ice0_delta = -xm;
      h2osno_total = h2osno_total - xm;
      h2osfc = h2osfc + xm;
      xmf_h2osfc = hm;
      qflx_h2osfc_to_ice = -xm / dtime;
      if (frac_sno_eff > 0 && snl < 0) snow_depth = h2osno_total / (rho_avg * frac_sno_eff);
      else snow_depth = h2osno_total / denice;
      if (snl == 0) {
        t0_written = true; t0_value = t_h2osfc;
        eflx_h2osfc_to_snow = 0.;
      } else {
        double c1v, c2v;
        if (snl == -1) c1v = frac_sno_eff * (dtime / fact0 - dhsdT * dtime);
        else c1v = frac_sno_eff / fact0 * dtime;
        if (frac_h2osfc != 0.0) c2v = (-cpliq * xm - frac_h2osfc * dhsdT * dtime);
        else c2v = 0.0;
        t[KOF(0)] = (c1v * t[KOF(0)] + c2v * t_h2osfc) / (c1v + c2v);
        eflx_h2osfc_to_snow = (t_h2osfc - t[KOF(0)]) * c2v / dtime;
      }
    } else {
      rho_avg = (h2osno_total * rho_avg + h2osfc * denice) / (h2osno_total + h2osfc);
      int_snow = int_snow + h2osfc;
      if (snl == 0) h2osno_no_layers = h2osno_no_layers + h2osfc;
      else ice0_delta = h2osfc;
      h2osno_total = h2osno_total + h2osfc;
      qflx_h2osfc_to_ice = h2osfc / dtime;
      t_h2osfc = t_h2osfc - temp1 * hfus / (dtime * dhsdT - c_h2osfc);
      xmf_h2osfc = (hm - frac_h2osfc * temp1 * hfus / dtime);
      if (snl == 0) {
        t0_written = true; t0_value = t_h2osfc;
      } else {
        double c1v, c2v;
        if (snl == -1) c1v = frac_sno_eff * (dtime / fact0 - dhsdT * dtime);
        else c1v = frac_sno_eff / fact0 * dtime;
        if (frac_h2osfc != 0.0) c2v = frac_h2osfc * (c_h2osfc - dtime * dhsdT);
        else c2v = 0.0;
        t[KOF(0)] = (c1v * t[KOF(0)] + c2v * t_h2osfc) / (c1v + c2v);
        t_h2osfc = t[KOF(0)];
      }
      h2osfc = 0.0;
      if (frac_sno_eff > 0 && snl < 0) snow_depth = h2osno_total / (rho_avg * frac_sno_eff);
      else snow_depth = h2osno_total / denice;
    }
    f.int_snow[ci] = int_snow;
    f.h2osfc[ci] = h2osfc;
  }
  f.xmf_h2osfc[ci] = xmf_h2osfc;
  f.qflx_h2osfc_to_ice[ci] = qflx_h2osfc_to_ice;
  f.eflx_h2osfc_to_snow[ci] = eflx_h2osfc_to_snow;
  f.t_h2osfc[ci] = t_h2osfc;
  if (t0_written) f.t_soisno[OS(0)] = t0_value;

  // ---- Phasechange :1133-1540 (one ascending sweep; per level: identify, then melt/freeze) --
  double xmf = 0.0, qflx_snomelt = 0.0, qflx_snofrz = 0.0, qflx_snow_drain = 0.0;
  double snomelt_accum = f.snomelt_accum[ci];
  for (int j = -NLEVSNO + 1; j <= 0; ++j) {          // :1268-1272 zeroed for all possible snow layers
    f.qflx_snomelt_lyr[OS(j)] = 0.0;
    f.qflx_snofrz_lyr[OS(j)] = 0.0;
  }
  for (int j = jtopl; j <= NLEVGRND; ++j) {
    const int k = KOF(j);
    double tj = t[k];
    double ice = f.h2osoi_ice[OS(j)];
    double liq = f.h2osoi_liq[OS(j)];
    if (j == 0) ice = ice + ice0_delta;              // PhaseChangeH2osfc's update of h2osoi_ice(c,0)
    const double wice0 = ice, wliq0 = liq;
    const double wmass0 = ice + liq;
    (void)wliq0;
    const double fa = factv[k];
    int imelt = 0;
    double tinc = 0.0, supercool = 0.0;
    if (j <= 0) {                                    // :1276-1299
      if (ice > 0.0 && tj > tfrz) { imelt = 1; tinc = tfrz - tj; tj = tfrz; }
      if (liq > 0.0 && tj < tfrz) { imelt = 2; tinc = tfrz - tj; tj = tfrz; }
    } else {                                         // :1302-1357
      if (ice > 0. && tj > tfrz) { imelt = 1; tinc = tfrz - tj; tj = tfrz; }
      if (soilcrop) {
        if (tj < tfrz) {
          const double smp = hfus * (tfrz - tj) / (grav * tj) * 1000.0;
          supercool = f.watsat[OG(j)] * pow(smp / f.sucsat[OG(j)], -1.0 / f.bsw[OG(j)]);
          supercool = supercool * f.dz[OS(j)] * 1000.0;
        }
      }
      if (liq > supercool && tj < tfrz) { imelt = 2; tinc = tfrz - tj; tj = tfrz; }
      if (h2osno_no_layers > 0.0 && j == 1) {
        if (tj > tfrz) { imelt = 1; tinc = tfrz - tj; tj = tfrz; }
      }
    }
    double hm = 0.0, xm = 0.0;
    if (imelt > 0) {                                 // :1373-1399
      if (j == jtopl) {
        if (j > 0) hm = dhsdT * tinc - tinc / fa;
        else hm = frac_sno_eff * (dhsdT * tinc - tinc / fa);
        if (j == 1 && frac_h2osfc != 0.0) hm = hm - frac_h2osfc * (dhsdT * tinc);
      } else if (j == 1) {
        hm = (1.0 - frac_sno_eff - frac_h2osfc) * dhsdT * tinc - tinc / fa;
      } else {
        if (j < 1) hm = -frac_sno_eff * (tinc / fa);
        else hm = -tinc / fa;
      }
    }
    if (imelt == 1 && hm < 0.0) { hm = 0.0; imelt = 0; }      // :1403-1410
    if (imelt == 2 && hm > 0.0) { hm = 0.0; imelt = 0; }
    if (imelt > 0 && fabs(hm) > 0.0) {                        // :1414
      xm = hm * dtime / hfus;
      if (j == 1) {                                           // :1420-1440
        if (h2osno_no_layers > 0.0 && xm > 0.0) {
          const double temp1 = h2osno_no_layers;
          h2osno_no_layers = fmax(0.0, temp1 - xm);
          const double propor = h2osno_no_layers / temp1;
          snow_depth = propor * snow_depth;
          const double heatr0 = hm - hfus * (temp1 - h2osno_no_layers) / dtime;
          if (heatr0 > 0.0) { xm = heatr0 * dtime / hfus; hm = heatr0; }
          else { xm = 0.0; hm = 0.0; }
          qflx_snomelt = fmax(0.0, (temp1 - h2osno_no_layers)) / dtime;
          xmf = hfus * qflx_snomelt;
          qflx_snow_drain = qflx_snomelt;
        }
      }
      double heatr = 0.0;
      if (xm > 0.0) {                                         // :1443-1457
        ice = fmax(0.0, wice0 - xm);
        heatr = hm - hfus * (wice0 - ice) / dtime;
      } else if (xm < 0.0) {                                  // :1458-1469
        if (j <= 0) {
          ice = fmin(wmass0, wice0 - xm);
        } else {
          if (wmass0 < supercool) ice = 0.0;
          else ice = fmin(wmass0 - supercool, wice0 - xm);
        }
        heatr = hm - hfus * (wice0 - ice) / dtime;
      }
      liq = fmax(0.0, wmass0 - ice);                          // :1471
      if (fabs(heatr) > 0.0) {                                // :1474-1501
        if (j == jtopl) {
          if (j == 1) tj = tj + fa * heatr / (1.0 - (1.0 - frac_h2osfc) * fa * dhsdT);
          else tj = tj + (fa / frac_sno_eff) * heatr / (1.0 - fa * dhsdT);
        } else if (j == 1) {
          tj = tj + fa * heatr / (1.0 - (1.0 - frac_sno_eff - frac_h2osfc) * fa * dhsdT);
        } else {
          if (j > 0) tj = tj + fa * heatr;
          else if (frac_sno_eff > 0.0) tj = tj + (fa / frac_sno_eff) * heatr;
        }
        if (j <= 0) {
          if (liq * ice > 0.0) tj = tfrz;
        }
      }
      if (j >= 1) xmf = xmf + hfus * (wice0 - ice) / dtime + 0.0;   // + hfus*(wexice0-excess_ice)/dtime == 0
      else xmf = xmf + hfus * (wice0 - ice) / dtime;
      if (imelt == 1 && j < 1) {                              // :1513-1517
        const double q = fmax(0.0, (wice0 - ice)) / dtime;
        f.qflx_snomelt_lyr[OS(j)] = q;
        qflx_snomelt = qflx_snomelt + q;
        snomelt_accum = snomelt_accum + q * dtime * 1.e-3;
      }
      if (imelt == 2 && j < 1) {                              // :1520-1523
        const double q = fmax(0.0, (ice - wice0)) / dtime;
        f.qflx_snofrz_lyr[OS(j)] = q;
        qflx_snofrz = qflx_snofrz + q;
      }
    }
    t[k] = tj;
    f.t_soisno[OS(j)] = tj;
    f.h2osoi_ice[OS(j)] = ice;
    f.h2osoi_liq[OS(j)] = liq;
    f.imelt[OS(j)] = imelt;
  }
  f.xmf[ci] = xmf;
  f.qflx_snomelt[ci] = qflx_snomelt;
  f.qflx_snofrz[ci] = qflx_snofrz;
  f.qflx_snow_drain[ci] = qflx_snow_drain;
  f.snomelt_accum[ci] = snomelt_accum;
  f.h2osno_no_layers[ci] = h2osno_no_layers;
  f.snow_depth[ci] = snow_depth;
  const double eflx_snomelt = qflx_snomelt * hfus;            // :1523-1534
  f.eflx_snomelt[ci] = eflx_snomelt;
  if (soilcrop) f.eflx_snomelt_r[ci] = eflx_snomelt;

  // ---- t_grnd :546-568 ------------------------------------------------------------------------
  double t_grnd;
  if (snl < 0) {
    if (frac_h2osfc != 0.0)
      t_grnd = frac_sno_eff * t[KOF(jtopl)] + (1.0 - frac_sno_eff - frac_h2osfc) * t[KOF(1)] + frac_h2osfc * t_h2osfc;
    else
      t_grnd = frac_sno_eff * t[KOF(jtopl)] + (1.0 - frac_sno_eff) * t[KOF(1)];
  } else {
    if (frac_h2osfc != 0.0) t_grnd = (1.0 - frac_h2osfc) * t[KOF(1)] + frac_h2osfc * t_h2osfc;
    else t_grnd = t[KOF(1)];
  }
  f.t_grnd[ci] = t_grnd;
#undef OS
#undef OZ
#undef OG
}

// The same routine with NO per-thread level arrays.  soiltemp_kernel keeps t, tk, fn, dzp, fact (5 x 37)
// and the LU factors U[38][5], y[38] in local memory: 3.3 KB per thread that spill through L2 to DRAM (17.6 GB per f02 call
// against 2.3 GB of fields).  Here the levels are streamed:
//   * thermal properties, interface conductivities, fluxes and fact are produced level by level, with one level of lookahead,
//     exactly when the row generator of the streaming LU asks for the next matrix row (rows are requested in ascending order);
//   * the finished LU rows, the eliminated right-hand side and the soil interface conductivities go to a coalesced scratch
//     [slot][filter position], written once and read once by the back substitution;
//   * the back substitution hands every new temperature straight to t_soisno and forms fn1 / eflx_fgr on the way up;
//   * PhaseChangeH2osfc / Phasechange read the new temperatures and fact back from their field arrays.
// Same expressions, same operation order per element: results are bit-identical to soiltemp_kernel
// (tests/test_gpu_soil.py::test_soiltemperature_kernels_agree_bit_for_bit).
#define ST_SCR_ROWS (NL + 1)
#define ST_SCR_SLOTS (6 * ST_SCR_ROWS + 2 * NLEVGRND)      // U[5] + y per matrix row, tk + dzp per soil level
struct SoilTempScratch {
  double* p;
  size_t stride;          // filter positions
  int fc;
  __device__ __forceinline__ void put_u(int j, int k, double v) { p[((size_t)j * 6 + k) * stride + fc] = v; }
  __device__ __forceinline__ void put_y(int j, double v) { p[((size_t)j * 6 + 5) * stride + fc] = v; }
  __device__ __forceinline__ double u(int i, int k) const { return p[((size_t)i * 6 + k) * stride + fc]; }
  __device__ __forceinline__ double y(int i) const { return p[((size_t)i * 6 + 5) * stride + fc]; }
  __device__ __forceinline__ double& tk(int j) { return p[((size_t)6 * ST_SCR_ROWS + (j - 1)) * stride + fc]; }            // j = 1..nlevgrnd
  __device__ __forceinline__ double& dzp(int j) { return p[((size_t)6 * ST_SCR_ROWS + NLEVGRND + (j - 1)) * stride + fc]; }
};

#ifndef ST_MINBLOCKS
#define ST_MINBLOCKS 4          /* 128 registers, 16 warps per SM: measured 3.02 ms at f02 against 3.87 (1), 3.28 (3), 3.40 (5) */
#endif
#define LDG(x) __ldg(&(x))          /* IN fields only: nothing in this kernel writes them */
__global__ void __launch_bounds__(128, ST_MINBLOCKS)
soiltemp_stream_kernel(SoilTempDev f, SoilTempPrm prm, int begc0, int ldc_, int begp0, int ldp_, int numc,
                       const int32_t* __restrict__ filterc, const int32_t* __restrict__ patchmask, double* __restrict__ scratch,
                       DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numc) return;
  const int c1 = filterc[fc];
  const int ci = c1 - begc0;
  const size_t ldc = (size_t)ldc_, ldp = (size_t)ldp_;
  const double dtime = prm.dtime;
  using namespace cst;

  const int lt = f.lun_itype[ci];
  if (lt >= CTSM_ISTURB_MIN && lt <= CTSM_ISTURB_MAX) { report_failure(ds, c1, CTSM_ERR_URBAN, 0); return; }
  const bool soilcrop = (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP);
  const int snl = f.snl[ci];
  const int jtopl = snl + 1;                      // top active layer
  const int nbed = f.nbedrock[ci];
  const double frac_sno_eff = f.frac_sno_eff[ci];
  const double frac_h2osfc = f.frac_h2osfc[ci];
  double h2osfc = f.h2osfc[ci];
  double h2osno_no_layers = f.h2osno_no_layers[ci];
  double t_h2osfc = f.t_h2osfc[ci];
  const double t_grnd_old = f.t_grnd[ci];

  // (c, j) offsets
#define OS(j) ((size_t)((j) - SNOSOI_LO) * ldc + ci)     /* (-nlevsno+1:nlevgrnd) arrays */
#define OZ(j) ((size_t)((j) - SNOSOI0_LO) * ldc + ci)    /* zi (-nlevsno:nlevgrnd)       */
#define OG(j) ((size_t)((j) - 1) * ldc + ci)             /* (1:nlevgrnd) arrays          */

  SoilTempScratch scr{scratch, (size_t)numc, fc};
  const double eflx_bot = f.eflx_bot[ci];

  // ---- SoilThermProp :602-901 + heat capacity for ONE level (the body of the reference's level loops) --------------
  struct Lvl { double t, thk, cv, z, zi, dz; };
  auto load_level = [&](int j) -> Lvl {
    Lvl L;
    const double tj = f.t_soisno[OS(j)];
    const double liq = f.h2osoi_liq[OS(j)];
    const double ice = f.h2osoi_ice[OS(j)];
    const double dzj = LDG(f.dz[OS(j)]);
    L.t = tj; L.dz = dzj; L.z = LDG(f.z[OS(j)]); L.zi = LDG(f.zi[OZ(j)]);
    double thk, cv;
    if (j >= 1) {
      if (lt != CTSM_ISTWET && lt != CTSM_ISTICE) {                          // :707-726
        const double watsat = LDG(f.watsat[OG(j)]);
        double satw = (liq / denh2o + ice / denice) / (dzj * watsat);
        satw = fmin(1.0, satw);
        if (satw > .1e-6) {
          double dke;
          if (tj >= tfrz) dke = fmax(0.0, log10(satw) + 1.0);
          else dke = satw;
          const double fl = (liq / (denh2o * dzj)) / (liq / (denh2o * dzj) + ice / (denice * dzj));
          const double dksat = LDG(f.tkmg[OG(j)]) * pow(tkwat, fl * watsat) * pow(tkice, (1.0 - fl) * watsat);
          thk = dke * dksat + (1.0 - dke) * LDG(f.tkdry[OG(j)]);
        } else {
          thk = LDG(f.tkdry[OG(j)]);
        }
        if (j > nbed) thk = thk_bedrock;
        cv = LDG(f.csol[OG(j)]) * (1.0 - watsat) * dzj + (ice * cpice + liq * cpliq);   // :847-849
        if (j > nbed) cv = csol_bedrock * dzj;
      } else if (lt == CTSM_ISTICE) {                                        // :727-729, :853-854
        thk = tkwat;
        if (tj < tfrz) thk = tkice;
        cv = (ice * cpice + liq * cpliq);
      } else {                                                               // istwet :730-737, :850-852
        if (j > NLEVSOI) thk = thk_bedrock;
        else { thk = tkwat; if (tj < tfrz) thk = tkice; }
        cv = (ice * cpice + liq * cpliq);
        if (j > nbed) cv = csol_bedrock * dzj;
      }
      if (j == 1 && h2osno_no_layers > 0.0) cv = cv + cpice * h2osno_no_layers;    // :875-880
    } else {                                                                 // snow :741-781, :884-895
      const double bw = (ice + liq) / (frac_sno_eff * dzj);
      f.bw[OS(j)] = bw;
      thk = snow_thk(lt == CTSM_ISTICE ? prm.snow_glc_method : prm.snow_method, bw);
      if (frac_sno_eff > 0.0) cv = fmax(thin_sfclayer, (cpliq * liq + cpice * ice) / frac_sno_eff);
      else cv = thin_sfclayer;
    }
    f.thk[OS(j)] = thk;
    L.thk = thk; L.cv = cv;
    return L;
  };
  const double t_top_old = f.t_soisno[OS(jtopl)], t1_old = f.t_soisno[OS(1)];      // t(c,snl+1), t(c,1) before the solve

  // ---- ComputeGroundHeatFluxAndDeriv :1543-1796 -----------------------------------
  const double emg = f.emg[ci], forc_lwrad = f.forc_lwrad[ci], htvp = f.htvp[ci];
  const double lwrad_emit = emg * sb * pow4(t_grnd_old);
  const double dlwrad_emit = 4.0 * emg * sb * pow3(t_grnd_old);
  const double lwrad_emit_snow = emg * sb * pow4(t_top_old);
  const double lwrad_emit_soil = emg * sb * pow4(t1_old);
  const double lwrad_emit_h2osfc = emg * sb * pow4(t_h2osfc);
  double hs_soil = 0.0, hs_h2osfc = 0.0, dhsdT = 0.0, hs_top = 0.0, hs_top_snow = 0.0;
  double sabg_lyr_col[NLEVSNO + 1];   // j = -nlevsno+1 .. 1
#pragma unroll
  for (int q = 0; q < NLEVSNO + 1; ++q) sabg_lyr_col[q] = 0.0;
  {
    const int pi = f.patchi[ci], pf = f.patchf[ci];
    for (int p1 = pi; p1 <= pf; ++p1) {
      const int pp = p1 - begp0;
      if (!LDG(patchmask[pp])) continue;
      const double wt = LDG(f.wtcol[pp]);
      const double fv = (double)LDG(f.frac_veg_nosno[pp]);
      const double dlrad = LDG(f.dlrad[pp]);
      const double lwdn = (1.0 - fv) * emg * forc_lwrad;
      const double sabg_soil = LDG(f.sabg_soil[pp]);
      const double eflx_gnet = LDG(f.sabg[pp]) + dlrad + lwdn - lwrad_emit
                               - (LDG(f.eflx_sh_grnd[pp]) + LDG(f.qflx_evap_soi[pp]) * htvp);
      f.eflx_gnet[pp] = eflx_gnet;
      f.sabg_chk[pp] = frac_sno_eff * LDG(f.sabg_snow[pp]) + (1.0 - frac_sno_eff) * sabg_soil;
      const double sh_soil = LDG(f.eflx_sh_soil[pp]), ev_soil = LDG(f.qflx_ev_soil[pp]);
      const double sh_snow = LDG(f.eflx_sh_snow[pp]), ev_snow = LDG(f.qflx_ev_snow[pp]);
      const double eflx_gnet_soil = sabg_soil + dlrad + lwdn - lwrad_emit_soil - (sh_soil + ev_soil * htvp);
      const double eflx_gnet_h2osfc = sabg_soil + dlrad + lwdn - lwrad_emit_h2osfc
                                      - (LDG(f.eflx_sh_h2osfc[pp]) + LDG(f.qflx_ev_h2osfc[pp]) * htvp);
      const double dgnetdT = -LDG(f.cgrnd[pp]) - dlwrad_emit;
      f.dgnetdT[pp] = dgnetdT;
      dhsdT = dhsdT + dgnetdT * wt;
      hs_soil = hs_soil + eflx_gnet_soil * wt;
      hs_h2osfc = hs_h2osfc + eflx_gnet_h2osfc * wt;
      // second patch loop of the reference (:1760-1792): separate accumulators, same patch order
      const double sabg_top = LDG(f.sabg_lyr[(size_t)(jtopl - SNOSOI_LO) * ldp + pp]);
      const double eflx_gnet_top = sabg_top + dlrad + lwdn - lwrad_emit - (LDG(f.eflx_sh_grnd[pp]) + LDG(f.qflx_evap_soi[pp]) * htvp);
      hs_top = hs_top + eflx_gnet_top * wt;
      const double eflx_gnet_snow = sabg_top + dlrad + lwdn - lwrad_emit_snow - (sh_snow + ev_snow * htvp);
      hs_top_snow = hs_top_snow + eflx_gnet_snow * wt;
#pragma unroll
      for (int q = 0; q < NLEVSNO + 1; ++q)            // j = q + SNOSOI_LO; absent snow layers are skipped (level-aligned across the warp)
        if (q >= jtopl - SNOSOI_LO) sabg_lyr_col[q] = sabg_lyr_col[q] + LDG(f.sabg_lyr[(size_t)q * ldp + pp]) * wt;
    }
  }
  (void)hs_top;   // only urban non-road columns consume hs_top (:2106-2112)

  // ---- h2osfc thermal properties :348-357 --------------------------------------------
  double c_h2osfc, dz_h2osfc;
  if ((h2osfc > thin_sfclayer) && (frac_h2osfc > thin_sfclayer)) {
    c_h2osfc = fmax(thin_sfclayer, cpliq * h2osfc / frac_h2osfc);
    dz_h2osfc = fmax(thin_sfclayer, R4(1.0e-3) * h2osfc / frac_h2osfc);
  } else {
    c_h2osfc = thin_sfclayer;
    dz_h2osfc = thin_sfclayer;
  }
  f.c_h2osfc[ci] = c_h2osfc;
  double dzm_ssw = 0.0, fn_h2osfc = 0.0, tk_h2osfc = 0.0;                         // set when the standing-water row is generated

  // ---- SetRHSVec / SetMatrix rows on demand, BandDiagonal -> dgbsv ---------------------
  // matrix row r in [snl, nlevgrnd]: r <= -1 snow layer j = r+1; r = 0 standing water; r >= 1 soil
  // generator state: `nx` = the next level (loaded, not yet finished), then per finished level its fact, interface
  // conductivity to the level below, spacing and old heat flux, and the same three of the level above it
  Lvl nx = load_level(jtopl);
  int jn = jtopl;
  double g_fa = 0.0, g_t = 0.0, g_tk = 0.0, g_dzp = 0.0, g_fn = 0.0, p_tk = 0.0, p_dzp = 0.0, p_fn = 0.0;
  double fn_1 = 0.0;                                     // fn of soil level 1 (eflx_fgr12)
  auto advance = [&](int j) {                            // finish level j (= jn), look ahead to j + 1
    p_tk = g_tk; p_dzp = g_dzp; p_fn = g_fn;
    const Lvl cur = nx;
    if (j < NLEVGRND) {
      nx = load_level(j + 1);
      jn = j + 1;
      g_dzp = nx.z - cur.z;                                                                           // :803-826
      g_tk = cur.thk * nx.thk * (nx.z - cur.z) / (cur.thk * (nx.z - cur.zi) + nx.thk * (cur.zi - cur.z));
      g_fn = g_tk * (nx.t - cur.t) / g_dzp;                                                           // :1904-1908
    } else {
      g_tk = 0.0; g_dzp = 0.0; g_fn = eflx_bot;
    }
    if (j == jtopl) {                                                                                 // :1887-1902
      const double zim = f.zi[OZ(j - 1)];
      g_fa = dtime / cur.cv * cur.dz / (0.5 * (cur.z - zim + capr * (nx.z - zim)));
    } else {
      g_fa = dtime / cur.cv;
    }
    f.fact[OS(j)] = g_fa;
    g_t = cur.t;
    if (j >= 1) {
      scr.tk(j) = g_tk; scr.dzp(j) = g_dzp;
      if (j == 1) fn_1 = g_fn;
      if (soilcrop && j < NLEVGRND) f.eflx_fgr[OG(j)] = -cnfac * g_fn;       // first term of :585; completed on the way up
    }
  };
  const int n = NLEVGRND - snl + 1;
  auto row = [&](int i, double* e) -> double {
    const int r = snl + i;
    e[0] = 0.0; e[1] = 0.0; e[2] = 0.0; e[3] = 0.0; e[4] = 0.0;
    double rhs;
    if (r <= -1) {                                   // SetRHSVec_Snow :2127-2144, SetMatrix_Snow :2608-2634
      const int j = r + 1;
      advance(j);
      const double fa = g_fa, tj = g_t;
      if (j == jtopl) {
        rhs = tj + fa * (hs_top_snow - dhsdT * tj + cnfac * g_fn);
        e[1] = 0.0;
        e[2] = 1.0 + (1.0 - cnfac) * fa * g_tk / g_dzp - fa * dhsdT;
      } else {
        rhs = tj + cnfac * fa * (g_fn - p_fn);
        rhs = rhs + fa * sabg_lyr_col[j - SNOSOI_LO];
        e[1] = -(1.0 - cnfac) * fa * p_tk / p_dzp;
        e[2] = 1.0 + (1.0 - cnfac) * fa * (g_tk / g_dzp + p_tk / p_dzp);
      }
      const double sup = -(1.0 - cnfac) * fa * g_tk / g_dzp;
      if (j != 0) e[3] = sup;                        // band 2
      else e[4] = sup;                               // band 1: snow layer 0 couples to soil row 1
    } else if (r == 0) {                             // :2199-2204, :2905-2911; nx is soil level 1 here
      const double zh2osfc = R4(1.0e-3) * (0.5 * h2osfc);                                             // :829-835
      tk_h2osfc = tkwat * nx.thk * (nx.z + zh2osfc) / (tkwat * nx.z + nx.thk * zh2osfc);
      dzm_ssw = (0.5 * dz_h2osfc + nx.z);                                                             // :2199, :2799, :2905
      fn_h2osfc = tk_h2osfc * (nx.t - t_h2osfc) / dzm_ssw;                                            // :2201
      rhs = t_h2osfc + (dtime / c_h2osfc) * (hs_h2osfc - dhsdT * t_h2osfc + cnfac * fn_h2osfc);
      e[2] = 1.0 + (1.0 - cnfac) * (dtime / c_h2osfc) * tk_h2osfc / dzm_ssw - (dtime / c_h2osfc) * dhsdT;
      e[3] = -(1.0 - cnfac) * (dtime / c_h2osfc) * tk_h2osfc / dzm_ssw;
    } else {                                         // SetRHSVec_Soil :2307-2349, SetMatrix_Soil :2746-2807
      const int j = r;
      advance(j);
      const double fa = g_fa, tj = g_t;
      if (j == jtopl) {                              // j == 1, no snow layers
        rhs = tj + fa * (hs_top_snow - dhsdT * tj + cnfac * g_fn);
        e[0] = 0.0;
        e[2] = 1.0 + (1.0 - cnfac) * fa * g_tk / g_dzp - fa * dhsdT;
        e[3] = -(1.0 - cnfac) * fa * g_tk / g_dzp;
      } else if (j == 1) {                           // snow/soil interface layer
        rhs = tj + fa * ((1.0 - frac_sno_eff) * (hs_soil - dhsdT * tj) + cnfac * (g_fn - frac_sno_eff * p_fn));
        rhs = rhs + frac_sno_eff * fa * sabg_lyr_col[1 - SNOSOI_LO];
        e[3] = -(1.0 - cnfac) * fa * g_tk / g_dzp;
        e[2] = 1.0 + (1.0 - cnfac) * fa * (g_tk / g_dzp + frac_sno_eff * p_tk / p_dzp)
               - (1.0 - frac_sno_eff) * fa * dhsdT;
        e[0] = -frac_sno_eff * (1.0 - cnfac) * fa * p_tk / p_dzp;   // band 5: column -1 (snow layer 0)
      } else if (j <= NLEVGRND - 1) {
        rhs = tj + cnfac * fa * (g_fn - p_fn);
        e[3] = -(1.0 - cnfac) * fa * g_tk / g_dzp;
        e[2] = 1.0 + (1.0 - cnfac) * fa * (g_tk / g_dzp + p_tk / p_dzp);
        e[1] = -(1.0 - cnfac) * fa * p_tk / p_dzp;
      } else {
        rhs = tj - cnfac * fa * p_fn + fa * g_fn;
        e[3] = 0.0;
        e[2] = 1.0 + (1.0 - cnfac) * fa * p_tk / p_dzp;
        e[1] = -(1.0 - cnfac) * fa * p_tk / p_dzp;
      }
      if (j == 1 && frac_h2osfc != 0.0) {            // :2342-2349, :2797-2807, :2918-2921
        rhs = rhs - frac_h2osfc * fa * ((hs_soil - dhsdT * tj) + cnfac * fn_h2osfc);
        e[2] = e[2] + frac_h2osfc * ((1.0 - cnfac) * fa * tk_h2osfc / dzm_ssw + fa * dhsdT);
        e[1] = -frac_h2osfc * (1.0 - cnfac) * fa * tk_h2osfc / dzm_ssw;       // band 4: column 0 (standing water)
      }
    }
    return rhs;
  };
  // back substitution: x(i) is the new temperature of matrix row r = snl + i (:422-434); fn1 / eflx_fgr (:438-483, :580-595)
  double y_h2osfc = 0.0, fn1_1 = 0.0, t_new_top = 0.0, t_new_1 = 0.0;
  const int info = band5_solve_stream(n, row, scr, [&](int i, double x, double xnext) {
    const int r = snl + i;
    if (r == 0) { y_h2osfc = x; return; }
    const int j = (r <= -1) ? r + 1 : r;
    f.t_soisno[OS(j)] = x;
    if (j == jtopl) t_new_top = x;
    if (j >= 1) {
      double fn1;
      if (j <= NLEVGRND - 1) fn1 = scr.tk(j) * (xnext - x) / scr.dzp(j);
      else fn1 = 0.0;
      if (j == 1) { fn1_1 = fn1; t_new_1 = x; }
      if (soilcrop) {
        if (j < NLEVGRND) f.eflx_fgr[OG(j)] = f.eflx_fgr[OG(j)] - (1.0 - cnfac) * fn1;
        else f.eflx_fgr[OG(j)] = 0.0;
      }
    }
  }, snl + NLEVSNO);     // aligned at soil level 1: every lane of a warp fetches / eliminates the same level in the same iteration
  if (info != 0) { report_failure(ds, c1, CTSM_ERR_DGBSV, info); return; }
  if (frac_h2osfc == 0.0) t_h2osfc = t_new_1;
  else t_h2osfc = y_h2osfc;
  const double eflx_fgr12 = -cnfac * fn_1 - (1.0 - cnfac) * fn1_1;
  f.eflx_fgr12[ci] = eflx_fgr12;

  // ---- PhaseChangeH2osfc :904-1130 --------------------------------------------------------
  double xmf_h2osfc = 0.0, qflx_h2osfc_to_ice = 0.0, eflx_h2osfc_to_snow = 0.0;
  double snow_depth = f.snow_depth[ci];
  double ice0_delta = 0.0;      // change applied to h2osoi_ice(c,0); folded in during the Phasechange sweep
  bool t0_written = false;      // t_soisno(c,0) written while snl == 0 (":1039 initialize for next time step")
  double t0_value = 0.0;
  if (frac_h2osfc > 0.0 && t_h2osfc <= tfrz) {
    double h2osno_total = h2osno_no_layers;                                   // CalculateTotalH2osno
    for (int j = jtopl; j <= 0; ++j) h2osno_total = h2osno_total + f.h2osoi_ice[OS(j)] + f.h2osoi_liq[OS(j)];
    double int_snow = f.int_snow[ci];
    const double tinc = tfrz - t_h2osfc;
    t_h2osfc = tfrz;
    const double hm = frac_h2osfc * (dhsdT * tinc - tinc * c_h2osfc / dtime);
    const double xm = hm * dtime / hfus;
    const double temp1 = h2osfc + xm;
    const double z_avg = frac_sno_eff * snow_depth;
    double rho_avg;
    if (z_avg > 0.0) rho_avg = fmin(800.0, h2osno_total / z_avg);
    else rho_avg = 200.0;
    const double fact0 = (snl < 0) ? f.fact[OS(0)] : 0.0;
    double t0 = (snl < 0) ? f.t_soisno[OS(0)] : 0.0;                   // t_soisno(c,0) after the solve
    if (temp1 >= 0.0) {
      int_snow = int_snow - xm;
      if (snl == 0) h2osno_no_layers = h2osno_no_layers - xm;
      else ice0_delta = -xm;
      h2osno_total = h2osno_total - xm;
      h2osfc = h2osfc + xm;
      xmf_h2osfc = hm;
      qflx_h2osfc_to_ice = -xm / dtime;
      if (frac_sno_eff > 0 && snl < 0) snow_depth = h2osno_total / (rho_avg * frac_sno_eff);
      else snow_depth = h2osno_total / denice;
      if (snl == 0) {
        t0_written = true; t0_value = t_h2osfc;
        eflx_h2osfc_to_snow = 0.;
      } else {
        double c1v, c2v;
        if (snl == -1) c1v = frac_sno_eff * (dtime / fact0 - dhsdT * dtime);
        else c1v = frac_sno_eff / fact0 * dtime;
        if (frac_h2osfc != 0.0) c2v = (-cpliq * xm - frac_h2osfc * dhsdT * dtime);
        else c2v = 0.0;
        t0 = (c1v * t0 + c2v * t_h2osfc) / (c1v + c2v);
        f.t_soisno[OS(0)] = t0;
        eflx_h2osfc_to_snow = (t_h2osfc - t0) * c2v / dtime;
      }
    } else {
      rho_avg = (h2osno_total * rho_avg + h2osfc * denice) / (h2osno_total + h2osfc);
      int_snow = int_snow + h2osfc;
      if (snl == 0) h2osno_no_layers = h2osno_no_layers + h2osfc;
      else ice0_delta = h2osfc;
      h2osno_total = h2osno_total + h2osfc;
      qflx_h2osfc_to_ice = h2osfc / dtime;
      t_h2osfc = t_h2osfc - temp1 * hfus / (dtime * dhsdT - c_h2osfc);
      xmf_h2osfc = (hm - frac_h2osfc * temp1 * hfus / dtime);
      if (snl == 0) {
        t0_written = true; t0_value = t_h2osfc;
      } else {
        double c1v, c2v;
        if (snl == -1) c1v = frac_sno_eff * (dtime / fact0 - dhsdT * dtime);
        else c1v = frac_sno_eff / fact0 * dtime;
        if (frac_h2osfc != 0.0) c2v = frac_h2osfc * (c_h2osfc - dtime * dhsdT);
        else c2v = 0.0;
        t0 = (c1v * t0 + c2v * t_h2osfc) / (c1v + c2v);
        f.t_soisno[OS(0)] = t0;
        t_h2osfc = t0;
      }
      h2osfc = 0.0;
      if (frac_sno_eff > 0 && snl < 0) snow_depth = h2osno_total / (rho_avg * frac_sno_eff);
      else snow_depth = h2osno_total / denice;
    }
    f.int_snow[ci] = int_snow;
    f.h2osfc[ci] = h2osfc;
  }
  f.xmf_h2osfc[ci] = xmf_h2osfc;
  f.qflx_h2osfc_to_ice[ci] = qflx_h2osfc_to_ice;
  f.eflx_h2osfc_to_snow[ci] = eflx_h2osfc_to_snow;
  f.t_h2osfc[ci] = t_h2osfc;
  if (t0_written) f.t_soisno[OS(0)] = t0_value;

  // ---- Phasechange :1133-1540 (one ascending sweep; per level: identify, then melt/freeze) --
  double xmf = 0.0, qflx_snomelt = 0.0, qflx_snofrz = 0.0, qflx_snow_drain = 0.0;
  double snomelt_accum = f.snomelt_accum[ci];
  for (int j = -NLEVSNO + 1; j <= 0; ++j) {          // :1268-1272 zeroed for all possible snow layers
    f.qflx_snomelt_lyr[OS(j)] = 0.0;
    f.qflx_snofrz_lyr[OS(j)] = 0.0;
  }
  for (int j = -NLEVSNO + 1; j <= NLEVGRND; ++j) {   // all lanes walk the same level (absent snow layers idle)
    if (j < jtopl) continue;
    double tj = f.t_soisno[OS(j)];
    double ice = f.h2osoi_ice[OS(j)];
    double liq = f.h2osoi_liq[OS(j)];
    if (j == 0) ice = ice + ice0_delta;              // PhaseChangeH2osfc's update of h2osoi_ice(c,0)
    const double wice0 = ice, wliq0 = liq;
    const double wmass0 = ice + liq;
    (void)wliq0;
    const double fa = f.fact[OS(j)];
    int imelt = 0;
    double tinc = 0.0, supercool = 0.0;
    if (j <= 0) {                                    // :1276-1299
      if (ice > 0.0 && tj > tfrz) { imelt = 1; tinc = tfrz - tj; tj = tfrz; }
      if (liq > 0.0 && tj < tfrz) { imelt = 2; tinc = tfrz - tj; tj = tfrz; }
    } else {                                         // :1302-1357
      if (ice > 0. && tj > tfrz) { imelt = 1; tinc = tfrz - tj; tj = tfrz; }
      if (soilcrop) {
        if (tj < tfrz) {
          const double smp = hfus * (tfrz - tj) / (grav * tj) * 1000.0;
          supercool = f.watsat[OG(j)] * pow(smp / f.sucsat[OG(j)], -1.0 / f.bsw[OG(j)]);
          supercool = supercool * f.dz[OS(j)] * 1000.0;
        }
      }
      if (liq > supercool && tj < tfrz) { imelt = 2; tinc = tfrz - tj; tj = tfrz; }
      if (h2osno_no_layers > 0.0 && j == 1) {
        if (tj > tfrz) { imelt = 1; tinc = tfrz - tj; tj = tfrz; }
      }
    }
    double hm = 0.0, xm = 0.0;
    if (imelt > 0) {                                 // :1373-1399
      if (j == jtopl) {
        if (j > 0) hm = dhsdT * tinc - tinc / fa;
        else hm = frac_sno_eff * (dhsdT * tinc - tinc / fa);
        if (j == 1 && frac_h2osfc != 0.0) hm = hm - frac_h2osfc * (dhsdT * tinc);
      } else if (j == 1) {
        hm = (1.0 - frac_sno_eff - frac_h2osfc) * dhsdT * tinc - tinc / fa;
      } else {
        if (j < 1) hm = -frac_sno_eff * (tinc / fa);
        else hm = -tinc / fa;
      }
    }
    if (imelt == 1 && hm < 0.0) { hm = 0.0; imelt = 0; }      // :1403-1410
    if (imelt == 2 && hm > 0.0) { hm = 0.0; imelt = 0; }
    if (imelt > 0 && fabs(hm) > 0.0) {                        // :1414
      xm = hm * dtime / hfus;
      if (j == 1) {                                           // :1420-1440
        if (h2osno_no_layers > 0.0 && xm > 0.0) {
          const double temp1 = h2osno_no_layers;
          h2osno_no_layers = fmax(0.0, temp1 - xm);
          const double propor = h2osno_no_layers / temp1;
          snow_depth = propor * snow_depth;
          const double heatr0 = hm - hfus * (temp1 - h2osno_no_layers) / dtime;
          if (heatr0 > 0.0) { xm = heatr0 * dtime / hfus; hm = heatr0; }
          else { xm = 0.0; hm = 0.0; }
          qflx_snomelt = fmax(0.0, (temp1 - h2osno_no_layers)) / dtime;
          xmf = hfus * qflx_snomelt;
          qflx_snow_drain = qflx_snomelt;
        }
      }
      double heatr = 0.0;
      if (xm > 0.0) {                                         // :1443-1457
        ice = fmax(0.0, wice0 - xm);
        heatr = hm - hfus * (wice0 - ice) / dtime;
      } else if (xm < 0.0) {                                  // :1458-1469
        if (j <= 0) {
          ice = fmin(wmass0, wice0 - xm);
        } else {
          if (wmass0 < supercool) ice = 0.0;
          else ice = fmin(wmass0 - supercool, wice0 - xm);
        }
        heatr = hm - hfus * (wice0 - ice) / dtime;
      }
      liq = fmax(0.0, wmass0 - ice);                          // :1471
      if (fabs(heatr) > 0.0) {                                // :1474-1501
        if (j == jtopl) {
          if (j == 1) tj = tj + fa * heatr / (1.0 - (1.0 - frac_h2osfc) * fa * dhsdT);
          else tj = tj + (fa / frac_sno_eff) * heatr / (1.0 - fa * dhsdT);
        } else if (j == 1) {
          tj = tj + fa * heatr / (1.0 - (1.0 - frac_sno_eff - frac_h2osfc) * fa * dhsdT);
        } else {
          if (j > 0) tj = tj + fa * heatr;
          else if (frac_sno_eff > 0.0) tj = tj + (fa / frac_sno_eff) * heatr;
        }
        if (j <= 0) {
          if (liq * ice > 0.0) tj = tfrz;
        }
      }
      if (j >= 1) xmf = xmf + hfus * (wice0 - ice) / dtime + 0.0;   // + hfus*(wexice0-excess_ice)/dtime == 0
      else xmf = xmf + hfus * (wice0 - ice) / dtime;
      if (imelt == 1 && j < 1) {                              // :1513-1517
        const double q = fmax(0.0, (wice0 - ice)) / dtime;
        f.qflx_snomelt_lyr[OS(j)] = q;
        qflx_snomelt = qflx_snomelt + q;
        snomelt_accum = snomelt_accum + q * dtime * 1.e-3;
      }
      if (imelt == 2 && j < 1) {                              // :1520-1523
        const double q = fmax(0.0, (ice - wice0)) / dtime;
        f.qflx_snofrz_lyr[OS(j)] = q;
        qflx_snofrz = qflx_snofrz + q;
      }
    }
    if (j == jtopl) t_new_top = tj;
    if (j == 1) t_new_1 = tj;
    f.t_soisno[OS(j)] = tj;
    f.h2osoi_ice[OS(j)] = ice;
    f.h2osoi_liq[OS(j)] = liq;
    f.imelt[OS(j)] = imelt;
  }
  f.xmf[ci] = xmf;
  f.qflx_snomelt[ci] = qflx_snomelt;
  f.qflx_snofrz[ci] = qflx_snofrz;
  f.qflx_snow_drain[ci] = qflx_snow_drain;
  f.snomelt_accum[ci] = snomelt_accum;
  f.h2osno_no_layers[ci] = h2osno_no_layers;
  f.snow_depth[ci] = snow_depth;
  const double eflx_snomelt = qflx_snomelt * hfus;            // :1523-1534
  f.eflx_snomelt[ci] = eflx_snomelt;
  if (soilcrop) f.eflx_snomelt_r[ci] = eflx_snomelt;

  // ---- t_grnd :546-568 ------------------------------------------------------------------------
  double t_grnd;
  if (snl < 0) {
    if (frac_h2osfc != 0.0)
      t_grnd = frac_sno_eff * t_new_top + (1.0 - frac_sno_eff - frac_h2osfc) * t_new_1 + frac_h2osfc * t_h2osfc;
    else
      t_grnd = frac_sno_eff * t_new_top + (1.0 - frac_sno_eff) * t_new_1;
  } else {
    if (frac_h2osfc != 0.0) t_grnd = (1.0 - frac_h2osfc) * t_new_1 + frac_h2osfc * t_h2osfc;
    else t_grnd = t_new_1;
  }
  f.t_grnd[ci] = t_grnd;
#undef OS
#undef OZ
#undef OG
}

extern "C" int ctsm_b200_soiltemperature(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakep,
                                         const int32_t* filter_nolakep, int num_nolakec,
                                         const int32_t* filter_nolakec, const ctsm_soiltemperature_fields_t* hf,
                                         int mem, ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_nolakec < 0 || num_nolakep < 0 || (num_nolakec > 0 && !filter_nolakec) ||
      (num_nolakep > 0 && !filter_nolakep))
    return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  SoilTempDev d;
  const int32_t *dfc = filter_nolakec, *dfp = filter_nolakep;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_SOILTEMPERATURE
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_SOILTEMPERATURE
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
  const int ldp = hf->alloc.endp - hf->alloc.begp + 1;
  const int ldc = hf->alloc.endc - hf->alloc.begc + 1;
  if ((size_t)ldp > ctx->patchmask_cap) {
    if (ctx->d_patchmask) CUDA_TRY(cudaFree(ctx->d_patchmask));
    ctx->d_patchmask = nullptr; ctx->patchmask_cap = 0;
    CUDA_TRY(cudaMalloc(&ctx->d_patchmask, sizeof(int32_t) * (size_t)ldp));
    ctx->patchmask_cap = (size_t)ldp;
  }
  CUDA_TRY(cudaMemsetAsync(ctx->d_patchmask, 0, sizeof(int32_t) * (size_t)ldp, ctx->stream));
  if (num_nolakep > 0) {
    patchmask_kernel<<<grid_for(num_nolakep, 256), 256, 0, ctx->stream>>>(ctx->d_patchmask, hf->alloc.begp, num_nolakep, dfp);
    ctx->launches++;
  }
  SoilTempPrm p{ctx->prm.dtime, ctx->prm.snow_thermal_cond_method, ctx->prm.snow_thermal_cond_glc_method};
  if (num_nolakec > 0 && ctx->tune.soil_stream) {
    int rc = arena_reserve(ctx, ctx->arena_scratch, sizeof(double) * (size_t)ST_SCR_SLOTS * (size_t)num_nolakec);
    if (rc) return rc;
    soiltemp_stream_kernel<<<grid_for(num_nolakec, 128), 128, 0, ctx->stream>>>(d, p, hf->alloc.begc, ldc, hf->alloc.begp, ldp,
                                                                                 num_nolakec, dfc, ctx->d_patchmask,
                                                                                 (double*)ctx->arena_scratch.p, ctx->d_status);
    ctx->launches++;
  } else if (num_nolakec > 0) {
    soiltemp_kernel<<<grid_for(num_nolakec, 128), 128, 0, ctx->stream>>>(d, p, hf->alloc.begc, ldc, hf->alloc.begp, ldp,
                                                                          num_nolakec, dfc, ctx->d_patchmask, ctx->d_status);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}
