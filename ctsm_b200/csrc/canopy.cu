// canopy.cu — CanopyFluxes (+ PhotosynthesisHydraulicStress) on B200.
//
// Reference: src/biogeophys/CanopyFluxesMod.F90:191-1765 with
//   FrictionVelocity / MoninObukIni   FrictionVelocityMod.F90:754-1209
//   QSat                              QSatMod.F90:61-127
//   calc_effective_soilporosity, calc_volumetric_h2oliq, calc_root_moist_stress
//                                     SoilMoistStressMod.F90:70-514
//   PhotosynthesisHydraulicStress     PhotosynthesisMod.F90:2704-3811 (callees in phs.cuh)
//   TimeStepInit :1143, PhotosynthesisTotal :2065, truncate_small_values NumericsMod.F90:50
//   setExposedvegpFilter              src/main/filterMod.F90:595-648
//
// B200 mapping.  The reference runs ~25 filter loops per ITERATION pass over clump-sized scratch arrays and compacts
// the patch filter on the host side of every pass.  Here one call is a chain of small kernels on one stream:
//   canopy_mark / colprep  eff_porosity / h2osoi_liqvol for the columns that own an exposed-vegetation patch (the
//                          reference writes them through a duplicate-laden patch->column list; writes are idempotent);
//   canopy_zero_kernel     TimeStepInit zeroing and rb1 = 0 for every patch in bounds;
//   canopy_init_kernel     one thread per filter patch: everything before the ITERATION loop, the iteration-invariant
//                          part of PHS (root-soil conductances: 20 pow per patch, which the reference recomputes on
//                          every pass), the constant part of the patch's PHS record, and the first night / day list;
//   per ITERATION pass k:  canopy_close_kernel (closes pass k-1, convergence test, survivors -> list of pass k),
//                          canopy_fric_kernel, canopy_leaf_kernel (open pass k), then the PHS solve of pass k as lane
//                          tasks with queue refill: phs_ci_kernel x4 interleaved with phs_newton_kernel /
//                          phs_newton_quad_kernel x4, then canopy_phs_end_kernel (see the comments above the kernels);
//   canopy_final_kernel    one thread per filter patch: PHS outputs of the patch's last pass, energy-balance check,
//                          stem temperature, ground fluxes, 2 m diagnostics, longwave, dew update, totals.
// Uniform kernels (one thread per listed patch) work on the Fortran arrays (patch index fastest: coalesced) and a
// structure-of-arrays workspace indexed by filter position; the irregular task kernels work on one contiguous record per
// patch (PhsRec).  Survivor lists are appended with warp-aggregated atomics: order is irrelevant to the numerics because
// patches are independent.
// Roofline: FP64 pipe / latency (hundreds of pow/exp/log per ~3 KB of patch traffic), not HBM;
// DESIGN.md section 4 has the algorithmic bytes and the measured breakdown.
#include "surface_layer.cuh"
#include <stdlib.h>

struct CanopyDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_CANOPYFLUXES
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_CANOPYFLUXES
#undef CTSM_F
};

namespace {


// workspace slots (structure of arrays, one row of `stride` doubles per slot, indexed by filter position)
enum Slot {
  W_AIR, W_BIR, W_CIR, W_SA_LEAF, W_SA_STEM, W_SA_INT, W_FRS, W_CP_LEAF, W_CP_STEM, W_RSTEM, W_DAYL, W_UR, W_ZLDIS,
  W_TL_INI, W_TS_INI, W_DEL, W_EFEB, W_EFE, W_OBUOLD, W_NMOZ, W_FM, W_EL, W_QSATL, W_QSATLDT, W_DELQ, W_DTH, W_DQH,
  W_TLBEF, W_DT_VEG, W_TEMP1, W_TEMP2, W_TEMP12M, W_TEMP22M, W_WTG, W_WTA0, W_WTL0, W_WTSTEM0, W_WTAL, W_WTGQ, W_WTAQ0,
  W_WTLQ0, W_WTALQ, W_LW_LEAF, W_LW_STEM, W_ERR, W_NSLOT
};

struct CanopyPrm {
  double dtime;
  int itmax, use_undercanopy_stability, use_biomass_heat_storage, z0param_method, soil_resis_method, use_luna,
      medlyn, light_inhibit, modifyphoto_and_lmr_forcrop, human_fast, hydrstress;
  double lai_dl, z_dl, a_coef, a_exp, csoilc, cv, wind_min, zetamaxstable, leaf_mr_vcm;
  double act25, fnr, cp25_yr2000, kc25_coef, ko25_coef, fnps, theta_psii, theta_ip;
  double vcmaxha, jmaxha, tpuha, lmrha, kcha, koha, cpha, vcmaxhd, jmaxhd, tpuhd, lmrhd, lmrse;
  double tpu25ratio, kp25ratio, vcmaxse_sf, jmaxse_sf, tpuse_sf, jmax25top_sf;
  // per-member values of scalar parameters (perturbed-parameter ensembles; NULL: the scalar above); member = itype / (mxpft+1)
  const double *m_csoilc, *m_cv, *m_a_coef, *m_z_dl;
};

struct Geo {   // index bases / leading dimensions
  int begp0, begc0, begg0, ldp, ldc;
  int begp, endp, begc, endc;   // call bounds
  int npft;                     // length of the PFT parameter tables (ctsm_params_t::npft_table)
};

// Active list of one ITERATION pass: ONE list in (roughly) ascending filter order.  Survivors are appended per warp, so a
// warp's 32 entries stay a run of neighbouring patches and the uniform kernels read whole 32-byte sectors; splitting
// the list into night / day bins (an earlier layout) halved the useful bytes per sector of kernels that are HBM bound.
#define NBIN 1
// PHS task queues of one pass (hybrid_PHS has at most 4 outer passes, PhotosynthesisMod.F90:3896):
//   ci queue i   (i = 0..3): day patches about to run outer pass i+1 of the ci solve (queue 0 is filled by the close kernel);
//   newton queue i (i = 0..3): patches that leave ci pass i+1: either on to calcstress before outer pass i+2, or, when
//   the ci solve has ended, to the hybrid_PHS epilogue (getvegwp); queue 0 also holds the night patches, whose whole
//   PHS solve is one calcstress.
#define NQ_CI 4
#define NQ_NT 4
#define QROW (NBIN + 2 * (NQ_CI + NQ_NT) + 1 + 2 * NQ_NT)   // ints per pass: list count, queue counts, queue fetch heads, a spare,
                                                            // then per calcstress stage: runnable-solve count and fetch head (split path)
struct Lists {
  int* counts;          // [npass + 2][QROW]: {bin counts[NBIN], ci count[NQ_CI], nt count[NQ_NT], ci head[NQ_CI], nt head[NQ_NT]}
  int* list_a;          // ping: [NBIN][cap]
  int* list_b;          // pong
  int* q_ci;            // [NQ_CI][cap]
  int* q_nt;            // [NQ_NT][cap]
  int* colflag;         // per column (alloc-based): owns an exposed-veg patch in this call
  int* ejected;         // [cap] per filter position: 1 once the patch has left the bulk rounds for the tail kernel
  int* tail_list;       // [cap] ejected patches in ejection order: fi | TAIL_B (entry point, see canopy_tail_kernel)
  int* tail_count;      // entries in tail_list so far
  int* tail_end;        // [npass + 2] tail_count at the end of every bulk round (canopy_tail_mark_kernel)
  int* tail_head;       // [npass + 2] fetch head of every round's tail kernel
  int cap;
  __device__ __forceinline__ int* n_ci(int row, int i) const { return counts + (size_t)row * QROW + NBIN + i; }
  __device__ __forceinline__ int* n_nt(int row, int i) const { return counts + (size_t)row * QROW + NBIN + NQ_CI + i; }
};

// Everything the PHS task kernels need about one exposed-vegetation patch, as ONE contiguous record (array of
// structures): the task kernels pick patches in queue order, i.e. scattered, so a record a lane can stream with
// 16-byte loads beats the structure-of-arrays layout that suits the uniform kernels.
struct alignas(128) PhsRec {
  // --- what a calcstress task streams (first five 128-byte lines) ---
  double psi50[4], ck[4], kmax[4];                  // constant during the call (canopy_init_kernel)
  double laisun, laisha, elai, esai, tsai, htop, fdry, forc_rho, forc_pbot, cf;
  double ksum, ksmp, ksmpg, smpg_mean;
  double qsatl, qaf, gb_mol;                        // per ITERATION pass (canopy_leaf_kernel)
  double x[4];                                      // vegwp at PHS entry (night: x[sun] = 1, the reference's sentinel)
  double gs0sun, gs0sha;                            // conductances handed to calcstress (phs::HybridCarry)
  int flags, iter1, patch, pad;
  double Kv[NLEVSOI], Gv[NLEVSOI];                  // k_soil_root(p,:), 1000 z(c,:)
  // --- what a ci task streams ---
  double qe, theta_cj, theta_ip, medint, medslope, bbb, mbb, cair, oair, par[2];     // constant during the call
  double rh_can, vcmax[2], tpu[2], kp[2], lmr[2], je[2], cp, kc, ko;                 // per pass
  double x1sun, x1sha, bsun, bsha, b0sun, b0sha;    // rest of phs::HybridCarry
  // --- results of the solve ---
  double gs_sun, gs_sha, tran, xo[4];
  phs::CiOut o;
  // --- epilogues only ---
  double Sv[NLEVSOI];                               // smp_l(c,:)
  phs::Brent br;                                    // brent_PHS state (rare: lives here, not in registers)
  __device__ __forceinline__ double K(int j) const { return Kv[j]; }
  __device__ __forceinline__ double G(int j) const { return Gv[j]; }
  __device__ __forceinline__ double S(int j) const { return Sv[j]; }
};
enum { RF_C3 = 1, RF_MEDLYN = 2, RF_NIGHT = 4, RF_SOLVE = 8, RF_FINAL = 16 };
// tail_list entry: filter position, plus TAIL_B when the patch was ejected in the middle of a pass (by a calcstress task
// that exceeded its iteration budget): its fric / leaf part of the pass is done and its PHS solve restarts from scratch.
#define TAIL_B (1 << 30)

// warp-aggregated append of `item` to bin `bin` of list (counts row `row`)
__device__ __forceinline__ void bin_append(const Lists& L, int* list, int row, int bin, int item, unsigned group) {
  const unsigned peers = __match_any_sync(group, bin);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  int b0 = 0;
  if (lane == leader) b0 = atomicAdd(&L.counts[(size_t)row * QROW + bin], __popc(peers));
  b0 = __shfl_sync(peers, b0, leader);
  list[(size_t)bin * L.cap + b0 + __popc(peers & ((1u << lane) - 1))] = item;
}

// statement functions PhotosynthesisMod.F90:2918-2920
__device__ __forceinline__ double ft(double tl, double ha) { return dexp(ha / (rgas * 1.e-3 * (tfrz + 25.0)) * (1.0 - (tfrz + 25.0) / tl)); }
__device__ __forceinline__ double fth(double tl, double hd, double se, double sc) { return sc / (1.0 + dexp((-hd + se * tl) / (rgas * 1.e-3 * tl))); }
__device__ __forceinline__ double fth25(double hd, double se) { return 1.0 + dexp((-hd + se * (tfrz + 25.0)) / (rgas * 1.e-3 * (tfrz + 25.0))); }

#define PF(name) f.name[pp]
#define PF2(name, j0) f.name[(size_t)(j0) * g.ldp + pp]           /* j0 = level - lower bound */
#define CF(name) f.name[cc]
#define CF2(name, j0) f.name[(size_t)(j0) * g.ldc + cc]
#define WS(slot) ws[(size_t)(slot) * wstride + fi]

// ---------------------------------------------------------------------------------------------
__global__ void canopy_mark_kernel(CanopyDev f, Geo g, int fn, const int32_t* __restrict__ filterp, int* __restrict__ colflag) {
  const int fi = blockIdx.x * blockDim.x + threadIdx.x;
  if (fi >= fn) return;
  const int pp = filterp[fi] - g.begp0;
  colflag[PF(column) - g.begc0] = 1;
}

// calc_effective_soilporosity :104-113, calc_volumetric_h2oliq :205-215 (jtop = 1)
__global__ void __launch_bounds__(128)
canopy_colprep_kernel(CanopyDev f, Geo g, const int* __restrict__ colflag) {
  const int cc = (g.begc - g.begc0) + blockIdx.x * blockDim.x + threadIdx.x;
  if (cc > g.endc - g.begc0) return;
  if (!colflag[cc]) return;
  for (int j = 1; j <= NLEVGRND; ++j) {
    const double watsat = CF2(watsat, j - 1);
    const double dz = CF2(dz, j - SNOSOI_LO);
    const double vol_ice = fmin(watsat, CF2(h2osoi_ice, j - SNOSOI_LO) / (denice * dz));
    const double eff = watsat - vol_ice;
    CF2(eff_porosity, j - 1) = eff;
    CF2(h2osoi_liqvol, j - SNOSOI_LO) = fmin(eff, CF2(h2osoi_liq, j - SNOSOI_LO) / (dz * denh2o));
  }
}

// ---------------------------------------------------------------------------------------------
// everything before the ITERATION loop (CanopyFluxesMod.F90:656-1023) + iteration-invariant PHS (:3063-3114)
__global__ void __launch_bounds__(128)
canopy_zero_kernel(CanopyDev f, Geo g) {
  // TimeStepInit :1158-1172 and rb1(begp:endp) = 0 (:830): every patch in bounds
  const int pp = (g.begp - g.begp0) + blockIdx.x * blockDim.x + threadIdx.x;
  if (pp > g.endp - g.begp0) return;
  if (!PF(patch_lakpoi)) {
    PF(psnsun) = 0.0; PF(psnsun_wc) = 0.0; PF(psnsun_wj) = 0.0; PF(psnsun_wp) = 0.0;
    PF(psnsha) = 0.0; PF(psnsha_wc) = 0.0; PF(psnsha_wj) = 0.0; PF(psnsha_wp) = 0.0;
    PF(fpsn) = 0.0; PF(fpsn_wc) = 0.0; PF(fpsn_wj) = 0.0; PF(fpsn_wp) = 0.0;
  }
  PF(rb1) = 0.0;
}

__global__ void __launch_bounds__(128)
canopy_init_kernel(CanopyDev f, CanopyPrm prm, Geo g, int fn, const int32_t* __restrict__ filterp, double* __restrict__ ws,
                   int wstride, Lists L, PhsRec* __restrict__ rec, DevStatus* ds) {
  const int fi = blockIdx.x * blockDim.x + threadIdx.x;
  if (fi >= fn) return;
  const int pp = filterp[fi] - g.begp0;
  const int cc = PF(column) - g.begc0;
  const int gg = PF(gridcell) - g.begg0;
  const int ivt = PF(itype);
  const double elai = PF(elai), esai = PF(esai), htop = PF(htop);

  PF(dhsdt_canopy) = 0.0;
  // biomass heat storage :716-818
  double frs, sa_leaf, sa_stem, sa_int, cp_leaf, cp_stem, rstem;
  if (prm.use_biomass_heat_storage) {
    const double fbw = f.pft_fbw[ivt], nstem = f.pft_nstem[ivt];
    frs = (esai) / (elai + esai);
    if (elai > 0.0) frs = 0.1 * frs;
    const double dbh = f.pft_dbh[ivt];
    sa_leaf = 2.0 * elai;
    sa_stem = 1.0 * (nstem * (htop * rpi * dbh));
    if (!(f.pft_is_tree[ivt] || f.pft_is_shrub[ivt]) || dbh < 0.05) {
      frs = 0.0; sa_stem = 0.0; sa_leaf = sa_leaf + esai;
    } else if (elai < 0.1) {
      sa_leaf = sa_leaf + esai;
    }
    const double leaf_biomass = (1.e-3 * c_to_b / f.pft_slatop[ivt]) * fmax(0.01, 0.5 * sa_leaf) / (1.0 - fbw);
    const double carea_stem = rpi * ((dbh * 0.5) * (dbh * 0.5));
    const double stem_biomass = carea_stem * htop * 1.0 * nstem * f.pft_wood_density[ivt] / (1.0 - fbw);
    PF(leaf_biomass) = leaf_biomass;
    PF(stem_biomass) = stem_biomass;
    sa_int = 0.0 * fmin(sa_leaf, sa_stem);
    cp_leaf = leaf_biomass * (c_dry_biomass * (1.0 - fbw) + (fbw)*c_water);
    cp_stem = 1.0 * (stem_biomass * (c_dry_biomass * (1.0 - fbw) + (fbw)*c_water));
    rstem = f.pft_rstem_per_dbh[ivt] * dbh;
  } else {
    sa_leaf = (elai + esai); frs = 0.0; sa_stem = 0.0; sa_int = 0.0; cp_leaf = 0.0; cp_stem = 0.0; rstem = 0.0;
  }
  WS(W_FRS) = frs; WS(W_SA_LEAF) = sa_leaf; WS(W_SA_STEM) = sa_stem; WS(W_SA_INT) = sa_int;
  WS(W_CP_LEAF) = cp_leaf; WS(W_CP_STEM) = cp_stem; WS(W_RSTEM) = rstem;
  {
    const double dayl = f.dayl[gg], mx = f.max_dayl[gg];
    WS(W_DAYL) = fmin(1.0, fmax(0.01, (dayl * dayl) / (mx * mx)));      // :827
  }

  // calc_root_moist_stress_clm45default :377-431 (SMS btran, rresis, rootr)
  {
    const double smpsc = f.pft_smpsc[ivt], smpso = f.pft_smpso[ivt];
    double btran = 0.0;
    for (int j = 1; j <= NLEVGRND; ++j) {
      const double liqvol = CF2(h2osoi_liqvol, j - SNOSOI_LO);
      double rootr = 0.0;
      if (!(liqvol <= 0.0 || CF2(t_soisno, j - SNOSOI_LO) <= tfrz - 2.0)) {
        const double effp = CF2(eff_porosity, j - 1);
        const double s_node = fmax(liqvol / effp, 0.01);
        double smp_node = -CF2(sucsat, j - 1) * pw(s_node, -CF2(bsw, j - 1));
        smp_node = fmax(smpsc, smp_node);
        const double rresis = fmin((effp / CF2(watsat, j - 1)) * (smp_node - smpsc) / (smpso - smpsc), 1.0);
        PF2(rresis, j - 1) = rresis;
        rootr = PF2(rootfr, j - 1) * rresis;
        btran = btran + fmax(rootr, 0.0);
      }
      PF2(rootr, j - 1) = rootr;
    }
    for (int j = 1; j <= NLEVGRND; ++j) {
      if (btran > 0.0) PF2(rootr, j - 1) = PF2(rootr, j - 1) / btran;
      else PF2(rootr, j - 1) = 0.0;
    }
    PF(btran) = btran;
  }

  // aerodynamic parameters :900-948
  double displa, z0mv;
  if (prm.z0param_method == 1) {
    const double lt = fmin(elai + esai, tlsai_crit);
    const double egvf = (1.0 - alpha_aero * dexp(-lt)) / (1.0 - alpha_aero * dexp(-tlsai_crit));
    displa = egvf * PF(displa);
    z0mv = dexp(egvf * dlog(PF(z0mv)) + (1.0 - egvf) * dlog(CF(z0mg)));
  } else {
    double lt = fmax(1.e-5, elai + esai);
    displa = htop * (1.0 - (1.0 - dexp(-pw(cd1_param * lt, 0.5))) / pw(cd1_param * lt, 0.5));
    lt = fmin(lt, f.pft_z0v_LAImax[ivt]);
    const double zc = f.pft_z0v_c[ivt], cw = f.pft_z0v_cw[ivt];
    const double ini = pw(f.pft_z0v_Cs[ivt] + f.pft_z0v_Cr[ivt] * lt * 0.5, -0.5) * zc * lt * 0.25;
    double U = ini, delt = 2.0;
    while (delt > 1.e-4) {
      const double prev = U;
      U = ini * dexp(prev);
      delt = fabs(U - prev);
    }
    U = 4.0 * U / lt / zc;
    z0mv = htop * (1.0 - displa / htop) * dexp(-vkc * U + dlog(cw) - 1.0 + 1.0 / cw);
  }
  PF(displa) = displa; PF(z0mv) = z0mv; PF(z0hv) = z0mv; PF(z0qv) = z0mv;
  const double hgt_u = f.forc_hgt_u[gg] + z0mv + displa;
  PF(forc_hgt_u_patch) = hgt_u;
  PF(forc_hgt_t_patch) = f.forc_hgt_t[gg] + z0mv + displa;
  PF(forc_hgt_q_patch) = f.forc_hgt_q[gg] + z0mv + displa;

  // :951-995
  const double emv = PF(emv), emg = CF(emg);
  WS(W_AIR) = emv * (1.0 + (1.0 - emv) * (1.0 - emg)) * CF(forc_lwrad);
  WS(W_BIR) = -(2.0 - emv * (1.0 - emg)) * emv * sb;
  WS(W_CIR) = emv * emg * sb;
  const double t_veg = PF(t_veg);
  const QS q0 = qsat(t_veg, CF(forc_pbot), true);
  WS(W_QSATL) = q0.qs; WS(W_EL) = q0.es; WS(W_QSATLDT) = q0.qsdT;
  WS(W_NMOZ) = 0.0;
  const double thm = PF(thm), forc_q = CF(forc_q);
  const double taf = (CF(t_grnd) + thm) / 2.0;
  const double qaf = (forc_q + CF(qg)) / 2.0;
  PF(taf) = taf; PF(qaf) = qaf;
  const double fu = f.forc_u[gg], fvv = f.forc_v[gg];
  const double ur = fmax(prm.wind_min, sqrt(fu * fu + fvv * fvv));
  const double dth = thm - taf, dqh = forc_q - qaf;
  WS(W_UR) = ur; WS(W_DTH) = dth; WS(W_DQH) = dqh;
  WS(W_DELQ) = CF(qg) - qaf;
  const double dthv = dth * (1.0 + 0.61 * forc_q) + 0.61 * CF(forc_th) * dqh;
  const double zldis = hgt_u - displa;
  WS(W_ZLDIS) = zldis;
  if (zldis < 0.0) report_failure(ds, pp + g.begp0, CTSM_ERR_FORC_HGT, 0);      // :990-1002

  // MoninObukIni :1187-1207
  {
    const double thv = CF(thv);
    double um;
    if (dthv >= 0.0) um = fmax(ur, 0.1);
    else um = sqrt(ur * ur + 0.5 * 0.5);
    const double rib = grav * zldis * dthv / (thv * um * um);
    double zeta;
    if (rib >= 0.0) {
      zeta = rib * dlog(zldis / z0mv) / (1.0 - 5.0 * fmin(rib, 0.19));
      zeta = fmin(prm.zetamaxstable, fmax(zeta, 0.01));
    } else {
      zeta = rib * dlog(zldis / z0mv);
      zeta = fmax(-100.0, fmin(zeta, -0.01));
    }
    PF(um) = um;
    PF(obu) = zldis / zeta;
  }
  PF(num_iter) = 0.0;
  WS(W_TL_INI) = t_veg;
  WS(W_TS_INI) = PF(t_stem);
  WS(W_DEL) = 0.0; WS(W_EFEB) = 0.0; WS(W_OBUOLD) = 0.0; WS(W_FM) = 0.0;
  WS(W_WTLQ0) = 0.0; WS(W_WTALQ) = 0.0; WS(W_WTGQ) = 0.0; WS(W_WTAQ0) = 0.0;
  PF(eflx_sh_stem) = 0.0;

  // iteration-invariant part of PhotosynthesisHydraulicStress: root-soil interface conductance :3063-3114, and the
  // constant part of the patch's PHS record (segment parameters, leaf areas, root-zone vectors and their level sums)
  PhsRec& R = rec[fi];
  if (prm.hydrstress) {
    const double froot_carbon = PF(froot_carbon), tsl = PF(tsai) + PF(tlai);
    const double rr = f.pft_root_radius[ivt], rd = f.pft_root_density[ivt], frl = f.pft_froot_leaf[ivt], krmax = f.pft_krmax[ivt];
    const double psi50r = f.pft_psi50[(size_t)phs::ROOT * g.npft + ivt], ckr = f.pft_ck[(size_t)phs::ROOT * g.npft + ivt];
    double ksum = 0.0, ksmp = 0.0, ksmpg = 0.0, smpg = 0.0;
    for (int j = 1; j <= NLEVSOI; ++j) {
      const double rootfr = PF2(rootfr, j - 1);
      double rbd = c_to_b * froot_carbon * rootfr / CF2(dz, j - SNOSOI_LO);
      rbd = fmax(c_to_b * 1.0, rbd);
      const double area = rpi * (rr * rr);
      const double rld = rbd / (rd * area);
      const double rai = tsl * frl * rootfr;
      const double r_soil = sqrt(1. / (rpi * rld));
      double soil_c = fmin(CF2(hksat, j - 1), CF2(hk_l, j - 1)) / (1.e3 * r_soil);
      const double sm = CF2(smp_l, j - 1);
      const double fs = phs::plc(sm, psi50r, ckr);
      const double zj = CF2(z, j - SNOSOI_LO);
      double root_c = (fs * rai * krmax) / (0.25 + zj);
      soil_c = fmax(soil_c, 1.e-16);
      root_c = fmax(root_c, 1.e-16);
      PF2(root_conductance, j - 1) = root_c;
      PF2(soil_conductance, j - 1) = soil_c;
      const double rs_resis = 1.0 / soil_c + 1.0 / root_c;
      const double k = (rai * rootfr > 0.0 && j > 1) ? 1.0 / rs_resis : 0.0;
      PF2(k_soil_root, j - 1) = k;
      const double gr = 1000.0 * zj;
      R.Kv[j - 1] = k; R.Gv[j - 1] = gr; R.Sv[j - 1] = sm;
      ksum += k; ksmp += k * sm; ksmpg += k * (sm - gr); smpg += sm - gr;
    }
    R.ksum = ksum; R.ksmp = ksmp; R.ksmpg = ksmpg; R.smpg_mean = smpg / NLEVSOI;
  }
  const bool night = (PF2(parsun_z, 0) <= 0.0);
  if (prm.hydrstress) {
#pragma unroll
    for (int sgm = 0; sgm < 4; ++sgm) {
      R.psi50[sgm] = f.pft_psi50[(size_t)sgm * g.npft + ivt];
      R.ck[sgm] = f.pft_ck[(size_t)sgm * g.npft + ivt];
      R.kmax[sgm] = f.pft_kmax[(size_t)sgm * g.npft + ivt];
    }
    const double forc_pbot = CF(forc_pbot);
    R.laisun = PF(laisun); R.laisha = PF(laisha); R.elai = elai; R.esai = esai; R.tsai = PF(tsai); R.htop = htop; R.fdry = PF(fdry);
    R.forc_rho = CF(forc_rho); R.forc_pbot = forc_pbot;
    R.cf = forc_pbot / (rgas * 1.e-3 * thm) * 1.e06;
    const bool c3 = ((int)nearbyint(f.pft_c3psn[ivt]) == 1);
    R.qe = c3 ? 0.0 : 0.05;
    R.bbb = c3 ? 10000.0 : 40000.0;
    R.mbb = f.pft_mbbopt[ivt];
    R.medint = f.pft_medlynintercept[ivt]; R.medslope = f.pft_medlynslope[ivt];
    R.theta_cj = f.pft_theta_cj[ivt]; R.theta_ip = prm.theta_ip;
    R.cair = f.forc_pco2[gg]; R.oair = f.forc_po2[gg];
    R.par[0] = PF2(parsun_z, 0); R.par[1] = PF2(parsha_z, 0);
    R.flags = (c3 ? RF_C3 : 0) | (prm.medlyn ? RF_MEDLYN : 0) | (night ? RF_NIGHT : 0);
    R.patch = pp + g.begp0;
    R.iter1 = 1;
  }

  // first active list
  bin_append(L, L.list_a, 0, 0, fi, __activemask());
}

// ---------------------------------------------------------------------------------------------
// ITERATION loop (CanopyFluxesMod.F90:1028-1457).  One pass k is a chain of small kernels, each with a code footprint
// that stays resident in the SM's instruction cache (a single fused "step" kernel was 150 KB of straight-line FP64
// code and spent 3/4 of its cycles waiting for instruction lines, profiles/):
//   canopy_close_kernel   closes pass k-1 for every patch that ran it (leaf energy balance :1174-1435, convergence test
//                         :1439-1457) and appends the survivors to the night / day bin of pass k;
//   canopy_fric_kernel    FrictionVelocity :1033 and the aerodynamic / leaf boundary-layer resistances :1038-1122;
//   canopy_leaf_kernel    temperature-dependent leaf biochemistry (PhotosynthesisMod.F90:3118-3469) and the per-pass
//                         part of the patch's PHS record;
//   phs_ci_kernel / phs_newton_kernel   the ci / plant-water-potential solve (:3477-3807) as lane tasks;
//   canopy_phs_end_kernel what follows the solve (:3587-3807).
// Everything that crosses a kernel boundary lives in the patch fields it belongs to, in the SoA workspace or in the
// PHS record.
#define STEP_THREADS 128
struct ListSlot { int fi; bool live; };
// thread slots of one pass list: bins padded to whole warps
__device__ __forceinline__ int list_offsets(const Lists& L, int row, int* off) {
  off[0] = 0;
#pragma unroll
  for (int b = 0; b < NBIN; ++b) off[b + 1] = off[b] + ((L.counts[(size_t)row * QROW + b] + 31) & ~31);
  return off[NBIN];
}
__device__ __forceinline__ ListSlot list_slot(const Lists& L, int row, const int* off, const int* __restrict__ list, int t) {
  ListSlot sl;
  int bin = 0;
#pragma unroll
  for (int b = 1; b < NBIN; ++b) bin += (t >= off[b]) ? 1 : 0;
  const int idx = t - off[bin];
  sl.live = idx < L.counts[(size_t)row * QROW + bin];
  sl.fi = sl.live ? list[(size_t)bin * L.cap + idx] : 0;
  return sl;
}

template <bool ALL>
__device__ __forceinline__ void phs_outputs(const CanopyDev& f, const CanopyPrm& prm, const Geo& g, const PhsRec& R, int pp,
                                            DevStatus* ds);

struct CloseOut { bool keep, solve, night; };
// closes pass itlef0-1 of one patch (leaf energy balance :1174-1435, convergence test :1439-1457); `first`: nothing to
// close yet.  Shared by the bulk kernel (one thread per list entry) and the tail kernel (one thread per patch, all passes).
__device__ __forceinline__ CloseOut close_body(const CanopyDev& f, const CanopyPrm& prm, const Geo& g, const int itlef0,
                                               const bool first, const int fi, const int32_t* __restrict__ filterp,
                                               double* __restrict__ ws, const int wstride, DevStatus* ds) {
  const double dtime = prm.dtime;
  bool keep = false, solve = false, night = false;
  {
    {
      const int pp = filterp[fi] - g.begp0;
      const int cc = PF(column) - g.begc0;
      const double forc_pbot = CF(forc_pbot), forc_rho = CF(forc_rho), forc_q = CF(forc_q), t_grnd = CF(t_grnd);
      const double thm = PF(thm), elai = PF(elai), esai = PF(esai), emv = PF(emv);
      const double ur = WS(W_UR), zldis_u = WS(W_ZLDIS);
      keep = true;
      solve = PF(nrad) >= 1;
      night = (PF2(parsun_z, 0) <= 0.0);
      if (!first) {
        // ---- close pass itlef0-1 ----
        const double laisun = PF(laisun), laisha = PF(laisha);
        double t_veg = PF(t_veg);
        const double t_stem = PF(t_stem);
        double um = PF(um), obu, taf, qaf = PF(qaf);
        const double ustar = PF(ustar), temp1 = WS(W_TEMP1), temp2 = WS(W_TEMP2);
        const double tlbef = t_veg, del2 = WS(W_DEL);
        const double rah_a = PF(rah1), raw_a = PF(raw1), rah_b = PF(rah2), raw_b = rah_b, uaf = PF(uaf), rb = PF(rb1);
        const double qsatl = WS(W_QSATL);
        const double rssun = PF(rssun), rssha = PF(rssha), btran = PF(btran), qflx_tran_veg = PF(qflx_tran_veg);
      // ---- leaf energy balance :1174-1435 ----
      const double sa_leaf = WS(W_SA_LEAF), sa_stem = WS(W_SA_STEM), sa_int = WS(W_SA_INT), frs = WS(W_FRS);
      const double cp_leaf = WS(W_CP_LEAF), rstem = WS(W_RSTEM), tl_ini = WS(W_TL_INI);
      const double air = WS(W_AIR), bir = WS(W_BIR), cir = WS(W_CIR);
      const double qsatldT = WS(W_QSATLDT);
      const double fvn = (double)PF(frac_veg_nosno);
      const double wta = 1.0 / rah_a;
      const double wtl = sa_leaf / rb;
      const double wtg = 1.0 / rah_b;
      const double wtstem = sa_stem / (rstem + rb);
      const double wtshi = 1.0 / (wta + wtl + wtstem + wtg);
      const double wtl0 = wtl * wtshi, wtg0 = wtg * wtshi, wta0 = wta * wtshi, wtstem0 = wtstem * wtshi;
      const double wtga = wta0 + wtg0 + wtstem0;
      const double wtal = wta0 + wtl0 + wtstem0;
      const double lw_stem = sa_int * emv * sb * pow4(t_stem);
      double lw_leaf = sa_int * emv * sb * pow4(t_veg);
      const double fdry = PF(fdry), fwet = PF(fwet);
      double rppdry;
      if (fdry > 0.0) rppdry = fdry * rb * (laisun / (rb + rssun) + laisha / (rb + rssha)) / elai;
      else rppdry = 0.0;
      double efpot = forc_rho * ((elai + esai) / rb) * (qsatl - qaf);
      const double h2ocan = PF(liqcan) + PF(snocan);
      double rpp;
      double tran = qflx_tran_veg;
      if (prm.hydrstress) {                            // :1219-1230
        if (efpot > 0.0) {
          rpp = (btran > 0.0) ? rppdry + fwet : fwet;
          rpp = fmin(rpp, (tran + h2ocan / dtime) / efpot);
        } else {
          rpp = 1.0;
        }
      } else {                                         // :1231-1248: transpiration follows the potential
        if (efpot > 0.0) {
          if (btran > 0.0) { tran = efpot * rppdry; rpp = rppdry + fwet; }
          else { rpp = fwet; tran = 0.0; }
          rpp = fmin(rpp, (tran + h2ocan / dtime) / efpot);
        } else {
          rpp = 1.0;
          tran = 0.0;
        }
      }
      const double wtaq = fvn / raw_a;
      const double wtlq = fvn * (elai + esai) / rb * rpp;
      const double z_dl = prm.m_z_dl ? prm.m_z_dl[PF(itype) / (CTSM_MXPFT + 1)] : prm.z_dl;
      const double fsno_dl = CF(snow_depth) / z_dl;
      const double elai_dl = prm.lai_dl * (1.0 - fmin(fsno_dl, 1.0));
      const double rdl = (1.0 - dexp(-elai_dl)) / (0.004 * uaf);
      double wtgq = WS(W_WTGQ);
      if (WS(W_DELQ) < 0.0) {
        wtgq = fvn / (raw_b + rdl);
      } else {
        if (prm.soil_resis_method == 0) wtgq = CF(soilbeta) * fvn / (raw_b + rdl);
        if (prm.soil_resis_method == 1) wtgq = fvn / (raw_b + CF(soilresis));
      }
      const double wtsqi = 1.0 / (wtaq + wtlq + wtgq);
      const double wtgq0 = wtgq * wtsqi, wtlq0 = wtlq * wtsqi, wtaq0 = wtaq * wtsqi;
      const double wtgaq = wtaq0 + wtgq0;
      const double wtalq = wtaq0 + wtlq0;
      const double dc1 = forc_rho * cpair * wtl;
      const double dc2 = hvap * forc_rho * wtlq;
      const double qg = CF(qg);
      const double efsh = dc1 * (wtga * t_veg - wtg0 * t_grnd - wta0 * thm - wtstem0 * t_stem);
      double eflx_sh_stem = forc_rho * cpair * wtstem * ((wta0 + wtg0 + wtl0) * t_stem - wtg0 * t_grnd - wta0 * thm - wtl0 * t_veg);
      double efe = dc2 * (wtgaq * qsatl - wtgq0 * qg - wtaq0 * forc_q);
      const double efeb = WS(W_EFEB);
      double erre = 0.0;
      if (efe * efeb < 0.0) {
        const double efeold = efe;
        efe = 0.1 * efeold;
        erre = efe - efeold;
      }
      const int snl = CF(snl);
      const double frac_sno = CF(frac_sno_eff), frac_h2osfc = CF(frac_h2osfc);
      const double lw_grnd = (frac_sno * pow4(CF2(t_soisno, snl + 1 - SNOSOI_LO)) + (1.0 - frac_sno - frac_h2osfc) * pow4(CF2(t_soisno, 1 - SNOSOI_LO))
                              + frac_h2osfc * pow4(CF(t_h2osfc)));
      const double sabv = PF(sabv);
      const double tv3 = pow3(t_veg), tv4 = pow4(t_veg);
      double dt_veg = ((1.0 - frs) * (sabv + air + bir * tv4 + cir * lw_grnd) - efsh - efe - lw_leaf + lw_stem - (cp_leaf / dtime) * (t_veg - tl_ini))
                      / ((1.0 - frs) * (-4.0 * bir * tv3) + 4.0 * sa_int * emv * sb * tv3 + dc1 * wtga + dc2 * wtgaq * qsatldT + cp_leaf / dtime);
      t_veg = tlbef + dt_veg;
      const double dels = dt_veg;
      const double del = fabs(dels);
      double err = 0.0;
      const double tb3 = pow3(tlbef);
      if (del > 1.0) {
        dt_veg = 1.0 * dels / del;
        t_veg = tlbef + dt_veg;
        err = (1.0 - frs) * (sabv + air + bir * tb3 * (tlbef + 4.0 * dt_veg) + cir * lw_grnd)
              - sa_int * emv * sb * tb3 * (tlbef + 4.0 * dt_veg) + lw_stem - (efsh + dc1 * wtga * dt_veg)
              - (efe + dc2 * wtgaq * qsatldT * dt_veg) - (cp_leaf / dtime) * (t_veg - tl_ini);
      }
      efpot = forc_rho * ((elai + esai) / rb) * (wtgaq * (qsatl + qsatldT * dt_veg) - wtgq0 * qg - wtaq0 * forc_q);
      double qflx_evap_veg = rpp * efpot;
      if (!prm.hydrstress) {                           // :1357-1362
        tran = (efpot > 0.0 && btran > 0.0) ? efpot * rppdry : 0.0;
        PF(qflx_tran_veg) = tran;
      }
      const double ecidif = fmax(0.0, qflx_evap_veg - tran - h2ocan / dtime);
      qflx_evap_veg = fmin(qflx_evap_veg, tran + h2ocan / dtime);
      PF(qflx_evap_veg) = qflx_evap_veg;
      PF(eflx_sh_veg) = efsh + dc1 * wtga * dt_veg + err + erre + hvap * ecidif;
      eflx_sh_stem = eflx_sh_stem + forc_rho * cpair * wtstem * (-wtl0 * dt_veg);
      PF(eflx_sh_stem) = eflx_sh_stem;
      lw_leaf = sa_int * emv * sb * tb3 * (tlbef + 4.0 * dt_veg);
      const QS q1 = qsat(t_veg, forc_pbot, true);
      taf = wtg0 * t_grnd + wta0 * thm + wtl0 * t_veg + wtstem0 * t_stem;
      qaf = wtlq0 * q1.qs + wtgq0 * qg + forc_q * wtaq0;
      const double dth = thm - taf, dqh = forc_q - qaf;
      const double delq = wtalq * qg - wtlq0 * q1.qs - wtaq0 * forc_q;
      const double tstar = temp1 * dth, qstar = temp2 * dqh;
      const double thv = CF(thv);
      const double thvstar = tstar * (1.0 + 0.61 * forc_q) + 0.61 * CF(forc_th) * qstar;
      double zeta = zldis_u * vkc * grav * thvstar / ((ustar * ustar) * thv);
      if (zeta >= 0.0) {
        zeta = fmin(prm.zetamaxstable, fmax(zeta, 0.01));
        um = fmax(ur, 0.1);
      } else {
        zeta = fmax(-100.0, fmin(zeta, -0.01));
        double wc;
        if (ustar * thvstar > 0.0) { wc = 0.0; atomicAdd(&ds->n_warnings, 1); }
        else wc = 1.0 * pw(-grav * ustar * thvstar * 1000.0 / thv, 0.333);
        um = sqrt(ur * ur + wc * wc);
      }
      obu = zldis_u / zeta;
      double nmoz = WS(W_NMOZ);
      const double obuold = WS(W_OBUOLD);
      if (obuold * obu < 0.0) nmoz = nmoz + 1.0;
      if (nmoz >= 4.0) obu = zldis_u / (-0.01);
      PF(t_veg) = t_veg; PF(taf) = taf; PF(qaf) = qaf; PF(zeta) = zeta; PF(um) = um; PF(obu) = obu;
      WS(W_NMOZ) = nmoz; WS(W_OBUOLD) = obu;
      WS(W_QSATL) = q1.qs; WS(W_EL) = q1.es; WS(W_QSATLDT) = q1.qsdT;
      WS(W_DTH) = dth; WS(W_DQH) = dqh; WS(W_DELQ) = delq; WS(W_TLBEF) = tlbef; WS(W_DT_VEG) = dt_veg; WS(W_DEL) = del;
      WS(W_WTG) = wtg; WS(W_WTA0) = wta0; WS(W_WTL0) = wtl0; WS(W_WTSTEM0) = wtstem0; WS(W_WTAL) = wtal; WS(W_WTGQ) = wtgq;
      WS(W_WTAQ0) = wtaq0; WS(W_WTLQ0) = wtlq0; WS(W_WTALQ) = wtalq; WS(W_LW_LEAF) = lw_leaf; WS(W_LW_STEM) = lw_stem;
      WS(W_EFE) = efe; WS(W_ERR) = err;

      // convergence :1439-1457
      keep = true;
      const int it1 = itlef0;
      if (it1 > 2) {
        const double dele = fabs(efe - efeb);
        WS(W_EFEB) = efe;
        const double det = fmax(del, del2);
        PF(num_iter) = (double)it1;
        keep = !(det < 0.01 && dele < 0.1);
      }
      }
    }
  }
  CloseOut o; o.keep = keep; o.solve = solve; o.night = night;
  return o;
}

#ifndef CLOSE_MINBLOCKS
#define CLOSE_MINBLOCKS 4
#endif
// Bulk round `itlef0`: (phs_outputs of pass itlef0-1, then) close pass itlef0-1 for every listed patch and append the
// survivors to the list of pass itlef0 - or, once the list of this round is short (<= tail_max entries), EJECT them to
// the round's tail list, where canopy_tail_kernel runs each patch through all its remaining passes (no more bulk rounds).
__global__ void __launch_bounds__(STEP_THREADS, CLOSE_MINBLOCKS)
canopy_close_kernel(CanopyDev f, CanopyPrm prm, Geo g, int itlef0, int first, int last, const int32_t* __restrict__ filterp,
                    double* __restrict__ ws, int wstride, Lists L, const int* __restrict__ list_in,
                    int* __restrict__ list_out, PhsRec* __restrict__ rec, int tail_max, DevStatus* ds) {
  const int row = itlef0;
  int off[NBIN + 1];
  const int total = list_offsets(L, row, off);
  const bool to_tail = !first && !last && L.counts[(size_t)row * QROW] <= tail_max;
  for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
    const ListSlot sl = list_slot(L, row, off, list_in, base + threadIdx.x);
    const int fi = sl.fi;
    const bool live = sl.live && !(L.ejected[fi]);     // ejected mid-pass by a task kernel: the tail kernel owns it now
    CloseOut c; c.keep = false; c.solve = false; c.night = false;
    if (live) {
      if (!first && prm.hydrstress) phs_outputs<false>(f, prm, g, rec[fi], filterp[fi] - g.begp0, ds);   // what follows the PHS solve of pass itlef0-1
      c = close_body(f, prm, g, itlef0, first != 0, fi, filterp, ws, wstride, ds);
    }
    const bool go = c.keep && !last;
    const unsigned act = __activemask();
    // survivors -> list of pass itlef0; those that need a PHS solve also enter their first task queue: night patches the
    // calcstress queue 0 (their whole solve is one calcstress), day patches the ci queue 0
    {
      const bool gt = go && to_tail;
      const unsigned mt = __ballot_sync(act, gt);
      if (gt) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(mt) - 1;
        int b0 = 0;
        if (lane == leader) b0 = atomicAdd(L.tail_count, __popc(mt));
        b0 = __shfl_sync(mt, b0, leader);
        L.tail_list[b0 + __popc(mt & ((1u << lane) - 1))] = fi;       // entry A: fric/leaf of pass itlef0 first
        L.ejected[fi] = 1;
      }
    }
    const bool gb = go && !to_tail;
    const unsigned mk = __ballot_sync(act, gb);
    if (gb) bin_append(L, list_out, row + 1, 0, fi, mk);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const bool gq = gb && c.solve && (c.night == (q == 0)) && prm.hydrstress;
      const unsigned mq = __ballot_sync(act, gq);
      if (gq) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(mq) - 1;
        int b0 = 0;
        if (lane == leader) b0 = atomicAdd(q == 0 ? L.n_nt(row + 1, 0) : L.n_ci(row + 1, 0), __popc(mq));
        b0 = __shfl_sync(mq, b0, leader);
        (q == 0 ? L.q_nt : L.q_ci)[b0 + __popc(mq & ((1u << lane) - 1))] = fi;
      }
    }
  }
}

// opens pass itlef0 of one patch, first half: FrictionVelocity and the aerodynamic / leaf boundary-layer resistances
__device__ __forceinline__ void fric_body(const CanopyDev& f, const CanopyPrm& prm, const Geo& g, const int itlef0, const int fi,
                                          const int32_t* __restrict__ filterp, double* __restrict__ ws, const int wstride) {
  {
    {
      const int pp = filterp[fi] - g.begp0;
      const int cc = PF(column) - g.begc0;
      const int ivt = PF(itype);
      const double forc_pbot = CF(forc_pbot), t_grnd = CF(t_grnd);
      const double elai = PF(elai), esai = PF(esai), htop = PF(htop);
      const double ur = WS(W_UR);
      const double um = PF(um), obu = PF(obu), taf = PF(taf), qaf = PF(qaf);
      const double displa = PF(displa), z0mv = PF(z0mv);
      // FrictionVelocity :1033-1036
      const FricOut fo = friction_velocity(PF(forc_hgt_u_patch), PF(forc_hgt_t_patch), PF(forc_hgt_q_patch), displa, z0mv,
                                           z0mv, z0mv, obu, itlef0 + 1, ur, um, WS(W_FM));
      const double ustar = fo.ustar, temp1 = fo.temp1, temp2 = fo.temp2;
      PF(ustar) = ustar; PF(vds) = fo.vds; PF(u10_clm) = fo.u10_clm; PF(va) = um; PF(u10) = fo.u10; PF(fv) = ustar;
      WS(W_FM) = fo.fm; WS(W_TEMP1) = temp1; WS(W_TEMP2) = temp2; WS(W_TEMP12M) = fo.temp12m; WS(W_TEMP22M) = fo.temp22m;

      // :1038-1122
      const double ram1 = 1.0 / (ustar * ustar / um);
      const double rah_a = 1.0 / (temp1 * ustar);
      const double raw_a = 1.0 / (temp2 * ustar);
      const double uaf = um * sqrt(1.0 / (ram1 * um));
      const double uuc = fmin(0.4, (0.03 * um / ustar));
      const double dleaf = f.pft_dleaf[ivt];
      const int mem_ = ivt / (CTSM_MXPFT + 1);
      const double p_cv = prm.m_cv ? prm.m_cv[mem_] : prm.cv, p_a_coef = prm.m_a_coef ? prm.m_a_coef[mem_] : prm.a_coef;
      const double p_csoilc = prm.m_csoilc ? prm.m_csoilc[mem_] : prm.csoilc;
      const double cfl = p_cv / (sqrt(uaf) * sqrt(dleaf));
      const double rb = 1.0 / (cfl * uaf);
      const double w = dexp(-(elai + esai));
      const double csoilb = vkc / (p_a_coef * pw(CF(z0mg) * uaf / nu_param, prm.a_exp));
      const double ri = (grav * htop * (taf - t_grnd)) / (taf * (uaf * uaf));
      double csoilcn;
      if (prm.use_undercanopy_stability && (taf - t_grnd) > 0.0) {
        const double ricsoilc = p_csoilc / (1.00 + 0.5 * fmin(ri, 10.0));
        csoilcn = csoilb * w + ricsoilc * (1.0 - w);
      } else {
        csoilcn = csoilb * w + p_csoilc * (1.0 - w);
      }
      const double rah_b = prm.use_biomass_heat_storage ? 1.0 / (csoilcn * uuc) : 1.0 / (csoilcn * uaf);
      const double raw_b = rah_b;
      const double svpts = WS(W_EL);
      const double eah = forc_pbot * qaf / 0.622;
      PF(ram1) = ram1; PF(uaf) = uaf; PF(dleaf_patch) = dleaf; PF(rb1) = rb;
      PF(rh_af) = eah / svpts;
      PF(rah1) = rah_a; PF(raw1) = raw_a; PF(rah2) = rah_b; PF(raw2) = raw_b;
      PF(vpd) = fmax((svpts - eah), 50.0) * 0.001;
    }
  }
}

// Residency of the two uniform kernels that open a pass.  Measured at f02 on B200 (CanopyFluxes call, ms): unconstrained
// (150 / 201 registers) 106.1; 4 blocks of 128 threads (128 registers, 16 warps per SM, no spills) 103.0; 5 blocks (96
// registers, spills) 103.3; 80 / 64 registers 108.8 / 110.9.  Both kernels wait on HBM (long scoreboard): more warps help
// until the spills start.
#ifndef FRIC_MINBLOCKS
#define FRIC_MINBLOCKS 4
#endif
#ifndef LEAF_MINBLOCKS
#define LEAF_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(STEP_THREADS, FRIC_MINBLOCKS)
canopy_fric_kernel(CanopyDev f, CanopyPrm prm, Geo g, int itlef0, const int32_t* __restrict__ filterp,
                   double* __restrict__ ws, int wstride, Lists L, const int* __restrict__ list_in) {
  const int row = itlef0 + 1;
  int off[NBIN + 1];
  const int total = list_offsets(L, row, off);
  for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
    const ListSlot sl = list_slot(L, row, off, list_in, base + threadIdx.x);
    if (!sl.live) continue;
    fric_body(f, prm, g, itlef0, sl.fi, filterp, ws, wstride);
  }
}

// opens pass itlef0 of one patch, second half: temperature-dependent leaf biochemistry and the per-pass part of the
// patch's PHS record
__device__ __forceinline__ void leaf_body(const CanopyDev& f, const CanopyPrm& prm, const Geo& g, const int itlef0, const int fi,
                                          const int32_t* __restrict__ filterp, double* __restrict__ ws, const int wstride,
                                          PhsRec* __restrict__ rec, DevStatus* ds) {
  {
    {
      {
      const int pp = filterp[fi] - g.begp0;
      const int cc = PF(column) - g.begc0;
      const int gg = PF(gridcell) - g.begg0;
      const int ivt = PF(itype);
      const double forc_pbot = CF(forc_pbot), thm = PF(thm);
      const double t_veg = PF(t_veg), qaf = PF(qaf), rb = PF(rb1);
      const double svpts = WS(W_EL);
      const double eah = forc_pbot * qaf / 0.622;
      phs::Leaf Lf;
      const bool c3 = ((int)nearbyint(f.pft_c3psn[ivt]) == 1);
      const double crop = f.pft_crop[ivt];
      Lf.c3 = c3; Lf.medlyn = prm.medlyn != 0;
      Lf.qe = c3 ? 0.0 : 0.05;
      Lf.bbb = c3 ? 10000.0 : 40000.0;
      Lf.mbb = f.pft_mbbopt[ivt];
      Lf.medint = f.pft_medlynintercept[ivt]; Lf.medslope = f.pft_medlynslope[ivt];
      Lf.theta_cj = f.pft_theta_cj[ivt]; Lf.theta_ip = prm.theta_ip;
      Lf.cair = f.forc_pco2[gg]; Lf.oair = f.forc_po2[gg];
      {
        const double kc25 = prm.kc25_coef * forc_pbot, ko25 = prm.ko25_coef * forc_pbot;
        const double sco = 0.5 * 0.209 / prm.cp25_yr2000;
        const double cp25 = 0.5 * Lf.oair / sco;
        Lf.kc = kc25 * ft(t_veg, prm.kcha);
        Lf.ko = ko25 * ft(t_veg, prm.koha);
        Lf.cp = cp25 * ft(t_veg, prm.cpha);
      }
      PF(c3flag) = c3 ? 1 : 0; PF(qe) = Lf.qe; PF(kc) = Lf.kc; PF(ko) = Lf.ko; PF(cp) = Lf.cp; PF(gb_mol) = (1.0 / rb) * (forc_pbot / (rgas * 1.e-3 * thm) * 1.e06);
      const double t10 = PF(t_a10), dayl_factor = WS(W_DAYL);
      const double lnc = fmin(1.0 / (f.pft_slatop[ivt] * f.pft_leafcn[ivt]), 10.0);
      PF(lnca) = lnc;
      double vcmax25top = lnc * f.pft_flnr[ivt] * prm.fnr * prm.act25 * dayl_factor;
      vcmax25top = vcmax25top * f.pft_fnitr[ivt];
      const double jmax25top = ((2.59 - 0.035 * fmin(fmax((t10 - tfrz), 11.0), 35.0)) * vcmax25top) * prm.jmax25top_sf;
      const double tpu25top = prm.tpu25ratio * vcmax25top;
      const double kp25top = prm.kp25ratio * vcmax25top;
      PF(luvcmax25top) = vcmax25top; PF(lujmax25top) = jmax25top; PF(lutpu25top) = tpu25top;
      const double lmr25top = c3 ? vcmax25top * prm.leaf_mr_vcm : vcmax25top * 0.025;
      const int nrad = PF(nrad);
      const double par_sun = PF2(parsun_z, 0), par_sha = PF2(parsha_z, 0);
      Lf.par[0] = par_sun; Lf.par[1] = par_sha;
      double jmax[2] = {0.0, 0.0};
      Lf.vcmax[0] = Lf.vcmax[1] = Lf.tpu[0] = Lf.tpu[1] = Lf.kp[0] = Lf.kp[1] = 0.0;
      if (nrad >= 1) {
        const double ns[2] = {PF(vcmaxcintsun), PF(vcmaxcintsha)};
        const bool luna = prm.use_luna && c3 && crop == 0.0;
        const double vcmx25 = PF2(vcmx25_z, 0);
        const double lmrc = fth25(prm.lmrhd, prm.lmrse);
        const double tl_fac = fmin((0.2 * dexp(3.218 * PF2(tlai_z, 0))), 1.0);
#pragma unroll
        for (int s = 0; s < 2; ++s) {                  // :3350-3374
          double lmr25 = lmr25top * ns[s];
          if (luna) lmr25 = prm.leaf_mr_vcm * vcmx25;
          double lmr;
          if (c3) {
            lmr = lmr25 * ft(t_veg, prm.lmrha) * fth(t_veg, prm.lmrhd, prm.lmrse, lmrc);
          } else {
            lmr = lmr25 * pw2((t_veg - (tfrz + 25.0)) / 10.0);
            lmr = lmr / (1.0 + dexp(1.3 * (t_veg - (tfrz + 55.0))));
          }
          Lf.lmr[s] = lmr * tl_fac;
        }
        if (!(par_sun <= 0.0)) {                       // day :3393-3456
          double v25[2], j25[2], t25[2];
          if (luna) {
            const double jmx25 = PF2(jmx25_z, 0);
            v25[0] = v25[1] = vcmx25; j25[0] = j25[1] = jmx25;
            t25[0] = prm.tpu25ratio * v25[0]; t25[1] = prm.tpu25ratio * v25[1];
            if (ns[0] > 0.0) {
              v25[1] = v25[0] * ns[1] / ns[0];
              j25[1] = j25[0] * ns[1] / ns[0];
              t25[1] = t25[0] * ns[1] / ns[0];
            }
          } else {
#pragma unroll
            for (int s = 0; s < 2; ++s) { v25[s] = vcmax25top * ns[s]; j25[s] = jmax25top * ns[s]; t25[s] = tpu25top * ns[s]; }
          }
          const double tc = fmin(fmax((t10 - tfrz), 11.0), 35.0);
          const double vcmaxse = (668.39 - 1.07 * tc) * prm.vcmaxse_sf;
          const double jmaxse = (659.70 - 0.75 * tc) * prm.jmaxse_sf;
          const double tpuse = (668.39 - 1.07 * tc) * prm.tpuse_sf;
          const double vcmaxc = fth25(prm.vcmaxhd, vcmaxse), jmaxc = fth25(prm.jmaxhd, jmaxse), tpuc = fth25(prm.tpuhd, tpuse);
          const double fv_ = ft(t_veg, prm.vcmaxha) , hv_ = fth(t_veg, prm.vcmaxhd, vcmaxse, vcmaxc);
          const double fj_ = ft(t_veg, prm.jmaxha), hj_ = fth(t_veg, prm.jmaxhd, jmaxse, jmaxc);
          const double ftp = ft(t_veg, prm.tpuha), htp = fth(t_veg, prm.tpuhd, tpuse, tpuc);
          const double q10 = pw2((t_veg - (tfrz + 25.0)) / 10.0);
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            Lf.vcmax[s] = v25[s] * fv_ * hv_;
            jmax[s] = j25[s] * fj_ * hj_;
            Lf.tpu[s] = t25[s] * ftp * htp;
            if (!c3) {
              double v = v25[s] * q10;
              v = v / (1.0 + dexp(0.2 * ((tfrz + 15.0) - t_veg)));
              v = v / (1.0 + dexp(0.3 * (t_veg - (tfrz + 40.0))));
              Lf.vcmax[s] = v;
            }
            Lf.kp[s] = (kp25top * ns[s]) * q10;
          }
        }
        if (prm.light_inhibit && par_sun > 0.0) Lf.lmr[0] = Lf.lmr[0] * 0.67;      // :3461-3466
        if (prm.light_inhibit && par_sha > 0.0) Lf.lmr[1] = Lf.lmr[1] * 0.67;
        PF2(lmrsun_z, 0) = Lf.lmr[0]; PF2(lmrsha_z, 0) = Lf.lmr[1];
        PF2(vcmax_z_phs, 0) = Lf.vcmax[0]; PF2(vcmax_z_phs, 1) = Lf.vcmax[1];
        PF2(tpu_z_phs, 0) = Lf.tpu[0]; PF2(tpu_z_phs, 1) = Lf.tpu[1];
        PF2(kp_z_phs, 0) = Lf.kp[0]; PF2(kp_z_phs, 1) = Lf.kp[1];
        }
        // hand the patch to the PHS task kernels: the per-pass part of its record (PhotosynthesisMod.F90:3477-3585)
        {
          PhsRec& R = rec[fi];
          const double cfm = forc_pbot / (rgas * 1.e-3 * thm) * 1.e06;
          R.qsatl = WS(W_QSATL); R.qaf = qaf; R.gb_mol = (1.0 / rb) * cfm;
          R.cp = Lf.cp; R.kc = Lf.kc; R.ko = Lf.ko;
#pragma unroll
          for (int s = 0; s < 2; ++s) { R.vcmax[s] = Lf.vcmax[s]; R.tpu[s] = Lf.tpu[s]; R.kp[s] = Lf.kp[s]; R.lmr[s] = (nrad >= 1) ? Lf.lmr[s] : 0.0; }
          int flags = R.flags & ~(RF_SOLVE | RF_FINAL);
          if (nrad >= 1) {
            flags |= RF_SOLVE;
            const double gsmin = Lf.medlyn ? Lf.medint : Lf.bbb;
            if (par_sun <= 0.0) {                        // night :3492-3496: one calcstress at the minimum conductance
              R.x[0] = 1.0; R.x[1] = PF2(vegwp, 1); R.x[2] = PF2(vegwp, 2); R.x[3] = PF2(vegwp, 3);
              R.gs0sun = gsmin; R.gs0sha = gsmin;
            } else {                                     // day :3549-3585
              const double ceair = fmin(eah, svpts);
              double rh_can;
              if (!Lf.medlyn) rh_can = ceair / svpts;
              else { rh_can = fmax((svpts - ceair), 50.0) * 0.001; PF(vpd_can) = rh_can; }
              R.rh_can = rh_can;
              bool bad_quad = false;
#pragma unroll
              for (int s = 0; s < 2; ++s) {
                const double qabs = 0.5 * (1.0 - prm.fnps) * Lf.par[s] * 4.6;
                const phs::Quad q = phs::quadratic(prm.theta_psii, -(qabs + jmax[s]), qabs * jmax[s], &bad_quad);
                R.je[s] = fmin(q.r1, q.r2);
              }
              if (bad_quad) report_failure(ds, pp + g.begp0, CTSM_ERR_QUADRATIC, 0);
#pragma unroll
              for (int i = 0; i < 4; ++i) R.x[i] = PF2(vegwp, i);
              phs::HybridCarry H;
              phs::hybrid_carry_init(H, (c3 ? 0.7 : 0.4) * Lf.cair);
              R.x1sun = H.x1sun; R.x1sha = H.x1sha; R.gs0sun = H.gs0sun; R.gs0sha = H.gs0sha;
              R.bsun = H.bsun; R.bsha = H.bsha; R.b0sun = H.b0sun; R.b0sha = H.b0sha;
              R.iter1 = H.iter1;
            }
          }
          R.flags = flags;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(STEP_THREADS, LEAF_MINBLOCKS)
canopy_leaf_kernel(CanopyDev f, CanopyPrm prm, Geo g, int itlef0, const int32_t* __restrict__ filterp,
                   double* __restrict__ ws, int wstride, Lists L, const int* __restrict__ list_in, PhsRec* __restrict__ rec,
                   DevStatus* ds) {
  const int row = itlef0 + 1;
  int off[NBIN + 1];
  const int total = list_offsets(L, row, off);
  for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
    const ListSlot sl = list_slot(L, row, off, list_in, base + threadIdx.x);
    if (!sl.live) continue;
    leaf_body(f, prm, g, itlef0, sl.fi, filterp, ws, wstride, rec, ds);
  }
}

// ---------------------------------------------------------------------------------------------
// Photosynthesis WITHOUT plant hydraulic stress (use_hydrstress = .false., clm4_5 physics):
// PhotosynthesisMod.F90 Photosynthesis :1243-2063, hybrid :2251-2400, brent :2403-2514, ci_func :2551-2701, called for the
// sunlit and then the shaded leaves of every patch of the pass (CanopyFluxesMod.F90:1143-1166).  One thread per listed
// patch runs both phases with the reference's nested loops: this is the secondary configuration, a single leaf per solve
// has no calcstress and at most 40 + 20 ci_func evaluations, and the lists are the same survivor lists as on the PHS path.
struct NoPhsLeaf {            // what ci_func reads (one phase of one patch)
  bool c3, medlyn;
  double vcmax, tpu, kp, cp, kc, ko, qe, je, par, lmr, theta_cj, theta_ip, medint, medslope, bbb, mbb, cair, oair, rh_can,
      gb_mol, forc_pbot;
};
struct NoPhsOut { double ac, aj, ap, ag, an; };
// ci_func :2551-2701
__device__ __noinline__ double nophs_ci_func(const NoPhsLeaf& L, double ci, double& gs_mol, NoPhsOut& o, bool* bad) {
  if (L.c3) {
    o.ac = L.vcmax * fmax(ci - L.cp, 0.0) / (ci + L.kc * (1.0 + L.oair / L.ko));
    o.aj = L.je * fmax(ci - L.cp, 0.0) / (4.0 * ci + 8.0 * L.cp);
    o.ap = 3.0 * L.tpu;
  } else {
    o.ac = L.vcmax;
    o.aj = L.qe * L.par * 4.6;
    o.ap = L.kp * fmax(ci, 0.0) / L.forc_pbot;
  }
  phs::Quad q = phs::quadratic(L.theta_cj, -(o.ac + o.aj), o.ac * o.aj, bad);
  const double ai = fmin(q.r1, q.r2);
  q = phs::quadratic(L.theta_ip, -(ai + o.ap), ai * o.ap, bad);
  o.ag = fmax(0.0, fmin(q.r1, q.r2));
  o.an = o.ag - L.lmr;
  if (o.an < 0.0) return 0.0;
  double cs = L.cair - 1.4 / L.gb_mol * o.an * L.forc_pbot;
  cs = fmax(cs, phs::max_cs);
  if (L.medlyn) {
    const double term = 1.6 * o.an / (cs / L.forc_pbot * 1.e06);
    const double bq = -(2.0 * (L.medint * 1.e-06 + term) + ((L.medslope * term) * (L.medslope * term)) / (L.gb_mol * 1.e-06 * L.rh_can));
    const double cq = L.medint * L.medint * 1.e-12 + (2.0 * L.medint * 1.e-06 + term * (1.0 - L.medslope * L.medslope / L.rh_can)) * term;
    q = phs::quadratic(1.0, bq, cq, bad);
    gs_mol = fmax(q.r1, q.r2) * 1.e06;
  } else {
    const double bq = cs * (L.gb_mol - L.bbb) - L.mbb * o.an * L.forc_pbot;
    const double cq = -L.gb_mol * (cs * L.bbb + L.mbb * o.an * L.forc_pbot * L.rh_can);
    q = phs::quadratic(cs, bq, cq, bad);
    gs_mol = fmax(q.r1, q.r2);
  }
  return ci - L.cair + o.an * L.forc_pbot * (1.4 * gs_mol + 1.6 * L.gb_mol) / (L.gb_mol * gs_mol);
}
// brent :2403-2514
__device__ __noinline__ double nophs_brent(const NoPhsLeaf& L, double x1, double x2, double f1, double f2, double tol, double& gs_mol,
                                           NoPhsOut& o, bool* bad, bool* notbracketed) {
  double a = x1, b = x2, fa = f1, fb = f2, c, fc, d = 0.0, e = 0.0;
  if ((fa > 0.0 && fb > 0.0) || (fa < 0.0 && fb < 0.0)) *notbracketed = true;
  c = b; fc = fb;
  for (int iter = 0; iter < 20;) {
    ++iter;
    if ((fb > 0.0 && fc > 0.0) || (fb < 0.0 && fc < 0.0)) { c = a; fc = fa; d = b - a; e = d; }
    if (fabs(fc) < fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
    const double tol1 = 2.0 * 1.e-2 * fabs(b) + 0.5 * tol;
    const double xm = 0.5 * (c - b);
    if (fabs(xm) <= tol1 || fb == 0.0) return b;
    if (fabs(e) >= tol1 && fabs(fa) > fabs(fb)) {
      const double sv = fb / fa;
      double pv, qv;
      if (a == c) {
        pv = 2.0 * xm * sv;
        qv = 1.0 - sv;
      } else {
        qv = fa / fc;
        const double rv = fb / fc;
        pv = sv * (2.0 * xm * qv * (qv - rv) - (b - a) * (rv - 1.0));
        qv = (qv - 1.0) * (rv - 1.0) * (sv - 1.0);
      }
      if (pv > 0.0) qv = -qv;
      pv = fabs(pv);
      if (2.0 * pv < fmin(3.0 * xm * qv - fabs(tol1 * qv), fabs(e * qv))) { e = d; d = pv / qv; }
      else { d = xm; e = d; }
    } else {
      d = xm; e = d;
    }
    a = b; fa = fb;
    if (fabs(d) > tol1) b = b + d;
    else b = b + copysign(tol1, xm);
    fb = nophs_ci_func(L, b, gs_mol, o, bad);
    if (fb == 0.0) break;
  }
  return b;
}
// hybrid :2251-2400 (the caller does not use the ci it returns)
__device__ __noinline__ void nophs_hybrid(const NoPhsLeaf& L, double x0, double& gs_mol, NoPhsOut& o, bool* bad, bool* notbracketed) {
  double f0 = nophs_ci_func(L, x0, gs_mol, o, bad);
  if (f0 == 0.0) return;
  double minx = x0, minf = f0;
  double x1 = x0 * 0.99;
  double f1 = nophs_ci_func(L, x1, gs_mol, o, bad);
  if (f1 == 0.0) return;
  if (f1 < minf) { minx = x1; minf = f1; }
  for (int iter = 0;;) {
    ++iter;
    const double dx = -f1 * (x1 - x0) / (f1 - f0);
    const double x = x1 + dx;
    const double tol = fabs(x) * 1.e-2;
    if (fabs(dx) < tol) break;
    x0 = x1; f0 = f1; x1 = x;
    f1 = nophs_ci_func(L, x1, gs_mol, o, bad);
    if (f1 < minf) { minx = x1; minf = f1; }
    if (fabs(f1) <= 1.e-4) break;
    if (f1 * f0 < 0.0) { (void)nophs_brent(L, x0, x1, f0, f1, tol, gs_mol, o, bad, notbracketed); break; }
    if (iter > 40) { (void)nophs_ci_func(L, minx, gs_mol, o, bad); break; }
  }
}

// Photosynthesis :1243-2063 for one patch, one phase (0 = sun, 1 = sha)
__device__ __noinline__ void nophs_photosynthesis(const CanopyDev& f, const CanopyPrm& prm, const Geo& g, int pp, int phase,
                                                  double esat_tv, double eair, double rb, double btran, double dayl_factor,
                                                  DevStatus* ds) {
  const int cc = PF(column) - g.begc0;
  const int gg = PF(gridcell) - g.begg0;
  const int ivt = PF(itype);
  const bool medlyn = prm.medlyn != 0;
  const double forc_pbot = CF(forc_pbot), t_veg = PF(t_veg), t10 = PF(t_a10);
  const double cair = f.forc_pco2[gg], oair = f.forc_po2[gg];
  const bool c3 = ((int)nearbyint(f.pft_c3psn[ivt]) == 1);
  PF(c3flag) = c3 ? 1 : 0;
  NoPhsLeaf L;
  L.c3 = c3; L.medlyn = medlyn;
  L.qe = c3 ? 0.0 : 0.05;
  PF(qe) = L.qe;
  L.bbb = 0.0; L.mbb = 0.0;
  if (!medlyn) { L.bbb = fmax((c3 ? 10000.0 : 40000.0) * btran, 1.0); L.mbb = f.pft_mbbopt[ivt]; }      // :1462-1465
  {
    const double kc25 = prm.kc25_coef * forc_pbot, ko25 = prm.ko25_coef * forc_pbot;
    const double sco = 0.5 * 0.209 / prm.cp25_yr2000;
    const double cp25 = 0.5 * oair / sco;
    L.kc = kc25 * ft(t_veg, prm.kcha);
    L.ko = ko25 * ft(t_veg, prm.koha);
    L.cp = cp25 * ft(t_veg, prm.cpha);
    PF(kc) = L.kc; PF(ko) = L.ko; PF(cp) = L.cp;
  }
  const double lnc = 1.0 / (f.pft_slatop[ivt] * f.pft_leafcn[ivt]);                                     // :1514
  PF(lnca) = lnc;
  double vcmax25top = lnc * f.pft_flnr[ivt] * prm.fnr * prm.act25 * dayl_factor;
  vcmax25top = vcmax25top * f.pft_fnitr[ivt];
  const double jmax25top = ((2.59 - 0.035 * fmin(fmax((t10 - tfrz), 11.0), 35.0)) * vcmax25top) * prm.jmax25top_sf;
  const double tpu25top = prm.tpu25ratio * vcmax25top;
  const double kp25top = prm.kp25ratio * vcmax25top;
  const double lmr25top = c3 ? vcmax25top * prm.leaf_mr_vcm : vcmax25top * 0.025;
  const int nrad = PF(nrad);
  const double par_z = phase == 0 ? PF2(parsun_z, 0) : PF2(parsha_z, 0);
  const double lai_z = phase == 0 ? PF2(laisun_z, 0) : PF2(laisha_z, 0);
  const double nscaler = phase == 0 ? PF(vcmaxcintsun) : PF(vcmaxcintsha);
  const double cf = forc_pbot / (rgas * 1.e-3 * PF(thm)) * 1.e06;
  const double gb_mol = (1.0 / rb) * cf;
  PF(gb_mol) = gb_mol;
  double lmr_z = 0.0, rs_z = 0.0, psn_z = 0.0, wc_z = 0.0, wj_z = 0.0, wp_z = 0.0;
  if (nrad >= 1) {
    const double crop = f.pft_crop[ivt];
    const bool luna = prm.use_luna && c3 && crop == 0.0;
    double lmr25 = lmr25top * nscaler;
    if (luna) lmr25 = prm.leaf_mr_vcm * PF2(vcmx25_z, 0);
    if (c3) {
      lmr_z = lmr25 * ft(t_veg, prm.lmrha) * fth(t_veg, prm.lmrhd, prm.lmrse, fth25(prm.lmrhd, prm.lmrse));
    } else {
      lmr_z = lmr25 * pw2((t_veg - (tfrz + 25.0)) / 10.0);
      lmr_z = lmr_z / (1.0 + dexp(1.3 * (t_veg - (tfrz + 55.0))));
    }
    double vcmax_z = 0.0, jmax_z = 0.0, tpu_z = 0.0, kp_z = 0.0;
    if (!(par_z <= 0.0)) {
      double vcmax25, jmax25, tpu25;
      if (luna) {
        vcmax25 = PF2(vcmx25_z, 0);
        jmax25 = PF2(jmx25_z, 0);
        tpu25 = prm.tpu25ratio * vcmax25;
        if (phase == 1 && PF(vcmaxcintsun) > 0.0) {
          const double a = PF(vcmaxcintsha), b = PF(vcmaxcintsun);
          vcmax25 = vcmax25 * a / b; jmax25 = jmax25 * a / b; tpu25 = tpu25 * a / b;
        }
      } else {
        vcmax25 = vcmax25top * nscaler; jmax25 = jmax25top * nscaler; tpu25 = tpu25top * nscaler;
      }
      const double kp25 = kp25top * nscaler;
      const double tc = fmin(fmax((t10 - tfrz), 11.0), 35.0);
      const double vcmaxse = (668.39 - 1.07 * tc) * prm.vcmaxse_sf;
      const double jmaxse = (659.70 - 0.75 * tc) * prm.jmaxse_sf;
      const double tpuse = (668.39 - 1.07 * tc) * prm.tpuse_sf;
      const double vcmaxc = fth25(prm.vcmaxhd, vcmaxse), jmaxc = fth25(prm.jmaxhd, jmaxse), tpuc = fth25(prm.tpuhd, tpuse);
      vcmax_z = vcmax25 * ft(t_veg, prm.vcmaxha) * fth(t_veg, prm.vcmaxhd, vcmaxse, vcmaxc);
      jmax_z = jmax25 * ft(t_veg, prm.jmaxha) * fth(t_veg, prm.jmaxhd, jmaxse, jmaxc);
      tpu_z = tpu25 * ft(t_veg, prm.tpuha) * fth(t_veg, prm.tpuhd, tpuse, tpuc);
      if (!c3) {
        vcmax_z = vcmax25 * pw2((t_veg - (tfrz + 25.0)) / 10.0);
        vcmax_z = vcmax_z / (1.0 + dexp(0.2 * ((tfrz + 15.0) - t_veg)));
        vcmax_z = vcmax_z / (1.0 + dexp(0.3 * (t_veg - (tfrz + 40.0))));
      }
      kp_z = kp25 * pw2((t_veg - (tfrz + 25.0)) / 10.0);
    }
    vcmax_z = vcmax_z * btran;                                                                          // :1745-1746
    lmr_z = lmr_z * btran;
    if (prm.light_inhibit && par_z > 0.0) lmr_z = lmr_z * 0.67;
    PF2(vcmax_z, 0) = vcmax_z; PF2(tpu_z, 0) = tpu_z; PF2(kp_z, 0) = kp_z;
    L.vcmax = vcmax_z; L.tpu = tpu_z; L.kp = kp_z; L.lmr = lmr_z; L.par = par_z;
    L.theta_cj = f.pft_theta_cj[ivt]; L.theta_ip = prm.theta_ip;
    L.medint = f.pft_medlynintercept[ivt]; L.medslope = f.pft_medlynslope[ivt];
    L.cair = cair; L.oair = oair; L.gb_mol = gb_mol; L.forc_pbot = forc_pbot;
    NoPhsOut o;
    double ci_z, gs_sunsha;
    if (par_z <= 0.0) {                                                                                 // night :1781-1815
      o.ac = 0.0; o.aj = 0.0; o.ap = 0.0; o.ag = 0.0;
      o.an = o.ag - lmr_z;
      rs_z = fmin(2.e4, 1.0 / (medlyn ? L.medint : L.bbb) * cf);
      ci_z = 0.0;
      PF(rh_leaf) = 0.0;
      gs_sunsha = cf / rs_z;
    } else {                                                                                            // day :1817-2006
      const double ceair = fmin(eair, esat_tv);
      double rh_can;
      if (!medlyn) rh_can = ceair / esat_tv;
      else { rh_can = fmax((esat_tv - ceair), 50.0) * 0.001; PF(vpd_can) = rh_can; }
      L.rh_can = rh_can;
      bool bad = false, nb = false;
      const double qabs = 0.5 * (1.0 - prm.fnps) * par_z * 4.6;
      const phs::Quad q = phs::quadratic(prm.theta_psii, -(qabs + jmax_z), qabs * jmax_z, &bad);
      L.je = fmin(q.r1, q.r2);
      double gs_mol = PF2(gs_mol, 0);
      nophs_hybrid(L, (c3 ? 0.7 : 0.4) * cair, gs_mol, o, &bad, &nb);
      if (bad) report_failure(ds, pp + g.begp0, CTSM_ERR_QUADRATIC, 0);
      if (nb) report_failure(ds, pp + g.begp0, CTSM_ERR_BRENT, 0);
      if (o.an < 0.0) gs_mol = medlyn ? L.medint : L.bbb;
      PF2(gs_mol, 0) = gs_mol;
      gs_sunsha = gs_mol;
      const int near_noon = f.near_local_noon[gg];
      if (phase == 0) PF2(gs_mol_sun_ln, 0) = near_noon ? gs_mol : spval;
      else PF2(gs_mol_sha_ln, 0) = near_noon ? gs_mol : spval;
      double cs = cair - 1.4 / gb_mol * o.an * forc_pbot;
      cs = fmax(cs, phs::max_cs);
      ci_z = cair - o.an * forc_pbot * (1.4 * gs_mol + 1.6 * gb_mol) / (gb_mol * gs_mol);
      ci_z = fmax(ci_z, 1.e-06);
      const double gs = gs_mol / cf;
      rs_z = fmin(1.0 / gs, 2.e4);
      rs_z = rs_z / (phase == 0 ? PF(o3coefgsun) : PF(o3coefgsha));
      psn_z = o.ag;
      psn_z = psn_z * (phase == 0 ? PF(o3coefvsun) : PF(o3coefvsha));
      if (o.ac <= o.aj && o.ac <= o.ap) wc_z = psn_z;
      else if (o.aj < o.ac && o.aj <= o.ap) wj_z = psn_z;
      else if (o.ap < o.ac && o.ap < o.aj) wp_z = psn_z;
      if (gs_mol < 0.0) report_failure(ds, pp + g.begp0, CTSM_ERR_GS_NEG, 0);
      if (!medlyn) {
        const double hs = (gb_mol * ceair + gs_mol * esat_tv) / ((gb_mol + gs_mol) * esat_tv);
        PF(rh_leaf) = hs;
        const double gs_mol_err = L.mbb * fmax(o.an, 0.0) * hs / cs * forc_pbot + L.bbb;
        if (fabs(gs_mol - gs_mol_err) > 1.e-01) atomicAdd(&ds->n_warnings, 1);
      }
    }
    PF2(ac, 0) = o.ac; PF2(aj, 0) = o.aj; PF2(ap, 0) = o.ap; PF2(ag, 0) = o.ag; PF2(an, 0) = o.an;
    if (phase == 0) { PF2(lmrsun_z, 0) = lmr_z; PF2(rssun_z, 0) = rs_z; PF2(cisun_z, 0) = ci_z; PF2(psnsun_z, 0) = psn_z; PF2(gs_mol_sun, 0) = gs_sunsha; }
    else { PF2(lmrsha_z, 0) = lmr_z; PF2(rssha_z, 0) = rs_z; PF2(cisha_z, 0) = ci_z; PF2(psnsha_z, 0) = psn_z; PF2(gs_mol_sha, 0) = gs_sunsha; }
  }
  // canopy sums :2015-2058 (nlevcan = 1)
  double psn = 0.0, pwc = 0.0, pwj = 0.0, pwp = 0.0, lmr = 0.0, rs = 0.0;
  {
    double psncan = 0.0, cwc = 0.0, cwj = 0.0, cwp = 0.0, lmrcan = 0.0, gscan = 0.0, laican = 0.0;
    if (nrad >= 1) {
      psncan = psncan + psn_z * lai_z; cwc = cwc + wc_z * lai_z; cwj = cwj + wj_z * lai_z; cwp = cwp + wp_z * lai_z;
      lmrcan = lmrcan + lmr_z * lai_z;
      gscan = gscan + lai_z / (rb + rs_z);
      laican = laican + lai_z;
    }
    if (laican > 0.0) { psn = psncan / laican; pwc = cwc / laican; pwj = cwj / laican; pwp = cwp / laican; lmr = lmrcan / laican; rs = laican / gscan - rb; }
  }
  if (phase == 0) { PF(psnsun) = psn; PF(psnsun_wc) = pwc; PF(psnsun_wj) = pwj; PF(psnsun_wp) = pwp; PF(lmrsun) = lmr; PF(rssun) = rs; }
  else { PF(psnsha) = psn; PF(psnsha_wc) = pwc; PF(psnsha_wj) = pwj; PF(psnsha_wp) = pwp; PF(lmrsha) = lmr; PF(rssha) = rs; }
}

__global__ void __launch_bounds__(STEP_THREADS)
canopy_photosyn_kernel(CanopyDev f, CanopyPrm prm, Geo g, int itlef0, const int32_t* __restrict__ filterp,
                       double* __restrict__ ws, int wstride, Lists L, const int* __restrict__ list_in, DevStatus* ds) {
  const int row = itlef0 + 1;
  int off[NBIN + 1];
  const int total = list_offsets(L, row, off);
  for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
    const ListSlot sl = list_slot(L, row, off, list_in, base + threadIdx.x);
    if (!sl.live) continue;
    const int fi = sl.fi;
    const int pp = filterp[fi] - g.begp0;
    const int cc = PF(column) - g.begc0;
    const double svpts = WS(W_EL), eah = CF(forc_pbot) * PF(qaf) / 0.622;
    const double rb = PF(rb1), btran = PF(btran), dayl = WS(W_DAYL);
    nophs_photosynthesis(f, prm, g, pp, 0, svpts, eah, rb, btran, dayl, ds);
    nophs_photosynthesis(f, prm, g, pp, 1, svpts, eah, rb, btran, dayl, ds);
  }
}

// ---------------------------------------------------------------------------------------------
// PHS task kernels.  A lane carries one task (one calcstress solve / one outer pass of the ci solve) from a queue
// and is REFILLED from the queue when its task ends, so the 32 lanes of a warp stay busy although the tasks need
// anything between 1 and 51 Newton iterations (2 and 26 ci_func evaluations).  All scheduling decisions are
// warp-uniform (taken from ballots), so every __*_sync below is executed by the full warp.
#define TASK_THREADS 64
#ifndef NT_MINBLOCKS
#define NT_MINBLOCKS 8
#endif
#ifndef CI_MINBLOCKS
#define CI_MINBLOCKS 6
#endif
#ifndef REFILL_MIN
#define REFILL_MIN 8          // refill when at least this many lanes are idle (or nothing else is left to do)
#endif
#ifndef CI_REFILL_MIN
#define CI_REFILL_MIN REFILL_MIN   // the same threshold serves the ci tasks (measured: 4 / 8 / 16 within 1 %)
#endif
#ifndef NT_SPLIT_REFILL
#define NT_SPLIT_REFILL 8     // split path: refill the iterate kernel's lanes when this many are idle
#endif
#ifndef FIN_MIN
#define FIN_MIN 16            // run the calcstress epilogue when at least this many lanes are waiting (for it or for work)
#endif
#define QUAD_MAX 16384        // calcstress queues up to this size run four lanes per solve (phs_newton_quad_kernel)
constexpr unsigned FULL = 0xffffffffu;

enum LaneState { LS_IDLE = 0, LS_RUN = 1, LS_FIN = 2 };


// push `item` of the lanes in `mask` (warp-uniform) to a queue
__device__ __forceinline__ void queue_push(int* __restrict__ q, int* __restrict__ count, unsigned mask, bool mine, int item) {
  if (mask == 0) return;
  const int lane = threadIdx.x & 31;
  int b0 = 0;
  if (lane == 0) b0 = atomicAdd(count, __popc(mask));
  b0 = __shfl_sync(FULL, b0, 0);
  if (mine) q[b0 + __popc(mask & ((1u << lane) - 1))] = item;
}

// calcstress tasks (PhotosynthesisMod.F90:4490-4710).  Day patches return to the ci queue `q_out`; for night patches
// the task is the whole solve (:3492-3547) and leaves potentials, stress factors and transpiration in the record.
__global__ void __launch_bounds__(TASK_THREADS, NT_MINBLOCKS)
phs_newton_kernel(PhsRec* __restrict__ rec, const int* __restrict__ q_in, const int* __restrict__ n_in, int* __restrict__ head,
                  int* __restrict__ q_out, int* __restrict__ n_out, Lists L, int budget) {
  extern __shared__ double shm[];                       // [2][NLEVSOI][TASK_THREADS]: k_soil_root, 1000 z
  double* sk = shm + threadIdx.x;
  double* sgv = shm + (size_t)NLEVSOI * TASK_THREADS + threadIdx.x;
  const int n = *n_in;
  if (n <= QUAD_MAX) return;                            // small queues: phs_newton_quad_kernel
  const int lane = threadIdx.x & 31;
  int st = LS_IDLE, fi = 0;
  bool exhausted = false;
  phs::Newton N;
  phs::NewtonCtx P;                                     // what newton_step reads: registers + shared root-zone vectors
  P.sk = sk; P.sg = sgv; P.stride = TASK_THREADS;
  for (;;) {
    const unsigned idle = __ballot_sync(FULL, st == LS_IDLE);
    const unsigned run = __ballot_sync(FULL, st == LS_RUN);
    const unsigned fin = __ballot_sync(FULL, st == LS_FIN);
    if (!exhausted && (__popc(idle) >= REFILL_MIN || (run | fin) == 0)) {
      int base = 0;
      if (lane == 0) base = atomicAdd(head, __popc(idle));
      base = __shfl_sync(FULL, base, 0);
      if (base + __popc(idle) >= n) exhausted = true;
      if (st == LS_IDLE) {
        const int my = base + __popc(idle & ((1u << lane) - 1));
        if (my < n) {
          fi = q_in[my];
          const PhsRec& R = rec[fi];
          if (R.flags & RF_FINAL) {
            st = LS_FIN;                                // the ci solve has ended: only the epilogue is left
          } else {
#pragma unroll
            for (int s = 0; s < 4; ++s) { P.psi50[s] = R.psi50[s]; P.ck[s] = R.ck[s]; }
            P.laisha = R.laisha; P.ksum = R.ksum; P.ksmp = R.ksmp;
#pragma unroll 4
            for (int j = 0; j < NLEVSOI; ++j) { sk[j * TASK_THREADS] = R.Kv[j]; sgv[j * TASK_THREADS] = R.Gv[j]; }
            const double xin[4] = {R.x[0], R.x[1], R.x[2], R.x[3]};
            st = phs::newton_begin(N, R, xin, R.gs0sun, R.gs0sha) ? LS_RUN : LS_FIN;
          }
        }
      }
      continue;
    }
    if ((run | fin) == 0) break;
    if (fin != 0 && (__popc(fin) + __popc(idle) >= FIN_MIN || run == 0)) {
      bool day = false;
      if (st == LS_FIN) {
        PhsRec& R = rec[fi];                            // the epilogues read the patch straight from its record
        if (R.flags & RF_FINAL) {
          // hybrid_PHS epilogue :4048-4062: potentials and transpiration at the converged conductances
          double x[4];
          double sf = phs::getvegwp(R, x, R.gs_sun, R.gs_sha);
          if (sf < 0.0) sf = 0.0;
          R.tran = sf;
#pragma unroll
          for (int i = 0; i < 4; ++i) R.xo[i] = x[i];
        } else {
          double tran = 0.0;
          const phs::Stress so = phs::newton_finish(N, R, R.gs0sun, R.gs0sha, &tran);
          R.bsun = so.bsun; R.bsha = so.bsha;
          if (R.flags & RF_NIGHT) {
#pragma unroll
            for (int i = 0; i < 4; ++i) R.xo[i] = N.x[i];
            R.tran = tran;
          } else {
            day = true;
          }
        }
        st = LS_IDLE;
      }
      const unsigned pm = __ballot_sync(FULL, day);
      queue_push(q_out, n_out, pm, day, fi);
      continue;
    }
    bool ej = false;
    if (st == LS_RUN) {
      if (!phs::newton_step(N, P)) st = LS_FIN;
      else if (N.iter >= budget) { ej = true; st = LS_IDLE; }     // a straggler: hand the patch to the tail kernel
    }
    const unsigned em = __ballot_sync(FULL, ej);
    if (em) {
      queue_push(L.tail_list, L.tail_count, em, ej, fi | TAIL_B);
      if (ej) L.ejected[fi] = 1;
    }
  }
}

// Small calcstress queues (the late ITERATION passes: a few thousand hard patches, and every pass waits for the
// slowest of them, 51 Newton iterations) are latency bound, so four lanes share one solve: each lane evaluates the
// Weibull curve of one plant segment (60 % of the dependent chain of an iteration) and the quad exchanges them by
// shuffle; everything else is computed redundantly, i.e. identically, by the four lanes.
struct WeibullQuad {
  unsigned qmask;
  int sgm, base;
  __device__ __forceinline__ void operator()(const double* x, const double* psi50, const double* ck, phs::Weibull* w) const {
    double xs = x[0], ps = psi50[0], cs = ck[0];
#pragma unroll
    for (int s = 1; s < 4; ++s) if (sgm == s) { xs = x[s]; ps = psi50[s]; cs = ck[s]; }
    const phs::Weibull mine = phs::weibull(xs, ps, cs, true);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      w[s].v = __shfl_sync(qmask, mine.v, base + s);
      w[s].d = __shfl_sync(qmask, mine.d, base + s);
    }
  }
};

__global__ void __launch_bounds__(128)
phs_newton_quad_kernel(PhsRec* __restrict__ rec, const int* __restrict__ q_in, const int* __restrict__ n_in,
                       int* __restrict__ q_out, int* __restrict__ n_out, Lists L, int budget) {
  const int n = *n_in;
  if (n <= 0 || n > QUAD_MAX) return;                   // large queues: phs_newton_kernel
  const int lane = threadIdx.x & 31;
  WeibullQuad wb;
  wb.sgm = lane & 3; wb.base = lane & ~3; wb.qmask = 0xFu << wb.base;
  const int quad = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, nquads = (gridDim.x * blockDim.x) >> 2;
  for (int task = quad; task < n; task += nquads) {
    const int fi = q_in[task];
    PhsRec& R = rec[fi];
    const int flags = R.flags;
    const double gs0sun = R.gs0sun, gs0sha = R.gs0sha;
    if (flags & RF_FINAL) {
      double x[4];
      double sf = phs::getvegwp(R, x, R.gs_sun, R.gs_sha);
      if (sf < 0.0) sf = 0.0;
      __syncwarp(wb.qmask);
      if (wb.sgm == 0) {
        R.tran = sf;
#pragma unroll
        for (int i = 0; i < 4; ++i) R.xo[i] = x[i];
      }
      continue;
    }
    phs::Newton N;
    const double xin[4] = {R.x[0], R.x[1], R.x[2], R.x[3]};
    bool ej = false;
    if (phs::newton_begin(N, R, xin, gs0sun, gs0sha)) {
      while (phs::newton_step(N, R, wb)) {
        if (N.iter >= budget) { ej = true; break; }             // quad-uniform: the four lanes carry the same solve
      }
    }
    if (ej) {
      if (wb.sgm == 0) { L.tail_list[atomicAdd(L.tail_count, 1)] = fi | TAIL_B; L.ejected[fi] = 1; }
      continue;
    }
    double tran = 0.0;
    const phs::Stress so = phs::newton_finish(N, R, gs0sun, gs0sha, &tran);
    __syncwarp(wb.qmask);
    if (wb.sgm == 0) {
      R.bsun = so.bsun; R.bsha = so.bsha;
      if (flags & RF_NIGHT) {
#pragma unroll
        for (int i = 0; i < 4; ++i) R.xo[i] = N.x[i];
        R.tran = tran;
      } else {
        q_out[atomicAdd(n_out, 1)] = fi;
      }
    }
  }
}

// Split path for LARGE calcstress queues.  In phs_newton_kernel the prologue (newton_begin: getqflx, 25 % of a mean task)
// and the epilogue (newton_finish: two Weibull curves, gs_from_qflx, for some a getvegwp with its 20-level sum; 40 %) run
// for whichever lanes happen to wait, at 8..16 of 32 lanes.  They are uniform per task, so they get their own kernels with
// one thread per queue entry (all lanes busy); the persistent kernel in between only iterates, and its refill is a
// coalesced load of the 10-double solver state plus the patch's root-zone vectors.  Same functions, same arithmetic.
struct NtState {          // structure of arrays over queue positions
  double* v;              // [10][cap]: x[4], qsun, qsha, ls, lh, tk, grav1
  int* w;                 // [cap]: iter | both << 8 | sha_only << 9 | night << 10 | flag << 11 | run << 12 | ejected << 13
  int cap;
};
enum { NW_BOTH = 1 << 8, NW_SHA = 1 << 9, NW_NIGHT = 1 << 10, NW_FLAG = 1 << 11, NW_RUN = 1 << 12, NW_EJECT = 1 << 13 };
__device__ __forceinline__ void nt_store(const NtState& S, int t, const phs::Newton& N, int extra) {
  const size_t c = (size_t)S.cap;
#pragma unroll
  for (int i = 0; i < 4; ++i) S.v[(size_t)i * c + t] = N.x[i];
  S.v[4 * c + t] = N.qsun; S.v[5 * c + t] = N.qsha; S.v[6 * c + t] = N.ls; S.v[7 * c + t] = N.lh; S.v[8 * c + t] = N.tk;
  S.v[9 * c + t] = N.grav1;
  S.w[t] = (N.iter & 0xff) | (N.both ? NW_BOTH : 0) | (N.sha_only ? NW_SHA : 0) | (N.night ? NW_NIGHT : 0) | (N.flag ? NW_FLAG : 0) | extra;
}
__device__ __forceinline__ int nt_load(const NtState& S, int t, phs::Newton& N) {
  const size_t c = (size_t)S.cap;
#pragma unroll
  for (int i = 0; i < 4; ++i) N.x[i] = S.v[(size_t)i * c + t];
  N.qsun = S.v[4 * c + t]; N.qsha = S.v[5 * c + t]; N.ls = S.v[6 * c + t]; N.lh = S.v[7 * c + t]; N.tk = S.v[8 * c + t];
  N.grav1 = S.v[9 * c + t];
  const int w = S.w[t];
  N.iter = w & 0xff; N.both = (w & NW_BOTH) != 0; N.sha_only = (w & NW_SHA) != 0; N.night = (w & NW_NIGHT) != 0; N.flag = (w & NW_FLAG) != 0;
  return w;
}

__global__ void __launch_bounds__(128)
nt_begin_kernel(const PhsRec* __restrict__ rec, const int* __restrict__ q_in, const int* __restrict__ n_in, NtState S,
                int* __restrict__ run_list, int* __restrict__ n_run) {
  const int n = *n_in;
  if (n <= QUAD_MAX) return;                            // small queues: phs_newton_quad_kernel
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const int t = base + threadIdx.x;
    bool run = false;
    if (t < n) {
      const PhsRec& R = rec[q_in[t]];
      if (!(R.flags & RF_FINAL)) {
        phs::Newton N;
        const double xin[4] = {R.x[0], R.x[1], R.x[2], R.x[3]};
        run = phs::newton_begin(N, R, xin, R.gs0sun, R.gs0sha);
        nt_store(S, t, N, run ? NW_RUN : 0);
      }
    }
    const unsigned act = __activemask();
    const unsigned m = __ballot_sync(act, run);
    if (run) {
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      int b0 = 0;
      if (lane == leader) b0 = atomicAdd(n_run, __popc(m));
      b0 = __shfl_sync(m, b0, leader);
      run_list[b0 + __popc(m & ((1u << lane) - 1))] = t;
    }
  }
}

__global__ void __launch_bounds__(TASK_THREADS, NT_MINBLOCKS)
nt_iter_kernel(const PhsRec* __restrict__ rec, const int* __restrict__ q_in, const int* __restrict__ n_in,
               const int* __restrict__ run_list, const int* __restrict__ n_run, int* __restrict__ head, NtState S, Lists L,
               int budget) {
  extern __shared__ double shm[];                       // [2][NLEVSOI][TASK_THREADS]: k_soil_root, 1000 z
  double* sk = shm + threadIdx.x;
  double* sgv = shm + (size_t)NLEVSOI * TASK_THREADS + threadIdx.x;
  if (*n_in <= QUAD_MAX) return;
  const int n = *n_run;
  const int lane = threadIdx.x & 31;
  bool running = false, exhausted = (n <= 0);
  int t = 0, fi = 0;
  phs::Newton N;
  phs::NewtonCtx P;
  P.sk = sk; P.sg = sgv; P.stride = TASK_THREADS;
  for (;;) {
    const unsigned idle = __ballot_sync(FULL, !running);
    if (!exhausted && (__popc(idle) >= NT_SPLIT_REFILL || idle == FULL)) {
      int base = 0;
      if (lane == 0) base = atomicAdd(head, __popc(idle));
      base = __shfl_sync(FULL, base, 0);
      if (base + __popc(idle) >= n) exhausted = true;
      if (!running) {
        const int my = base + __popc(idle & ((1u << lane) - 1));
        if (my < n) {
          t = run_list[my];
          fi = q_in[t];
          const PhsRec& R = rec[fi];
          (void)nt_load(S, t, N);
#pragma unroll
          for (int s = 0; s < 4; ++s) { P.psi50[s] = R.psi50[s]; P.ck[s] = R.ck[s]; }
          P.laisha = R.laisha; P.ksum = R.ksum; P.ksmp = R.ksmp;
#pragma unroll 4
          for (int j = 0; j < NLEVSOI; ++j) { sk[j * TASK_THREADS] = R.Kv[j]; sgv[j * TASK_THREADS] = R.Gv[j]; }
          running = true;
        }
      }
      continue;
    }
    if (idle == FULL) break;
    bool ej = false;
    if (running) {
      bool more = phs::newton_step(N, P);
      if (more && N.iter >= budget) { ej = true; more = false; }                  // a straggler: to the tail kernel
      if (!more) { nt_store(S, t, N, ej ? NW_EJECT : 0); running = false; }
    }
    const unsigned em = __ballot_sync(FULL, ej);
    if (em) {
      queue_push(L.tail_list, L.tail_count, em, ej, fi | TAIL_B);
      if (ej) L.ejected[fi] = 1;
    }
  }
}

__global__ void __launch_bounds__(128)
nt_finish_kernel(PhsRec* __restrict__ rec, const int* __restrict__ q_in, const int* __restrict__ n_in, NtState S,
                 int* __restrict__ q_out, int* __restrict__ n_out) {
  const int n = *n_in;
  if (n <= QUAD_MAX) return;
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const int t = base + threadIdx.x;
    bool day = false;
    int fi = 0;
    if (t < n) {
      fi = q_in[t];
      PhsRec& R = rec[fi];
      if (R.flags & RF_FINAL) {
        double x[4];                                    // hybrid_PHS epilogue :4048-4062
        double sf = phs::getvegwp(R, x, R.gs_sun, R.gs_sha);
        if (sf < 0.0) sf = 0.0;
        R.tran = sf;
#pragma unroll
        for (int i = 0; i < 4; ++i) R.xo[i] = x[i];
      } else {
        phs::Newton N;
        const int w = nt_load(S, t, N);
        if (!(w & NW_EJECT)) {
          double tran = 0.0;
          const phs::Stress so = phs::newton_finish(N, R, R.gs0sun, R.gs0sha, &tran);
          R.bsun = so.bsun; R.bsha = so.bsha;
          if (R.flags & RF_NIGHT) {
#pragma unroll
            for (int i = 0; i < 4; ++i) R.xo[i] = N.x[i];
            R.tran = tran;
          } else {
            day = true;
          }
        }
      }
    }
    const unsigned act = __activemask();
    const unsigned m = __ballot_sync(act, day);
    if (day) {
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      int b0 = 0;
      if (lane == leader) b0 = atomicAdd(n_out, __popc(m));
      b0 = __shfl_sync(m, b0, leader);
      q_out[b0 + __popc(m & ((1u << lane) - 1))] = fi;
    }
  }
}

// what a ci task reads of its patch, and how one outer pass of hybrid_PHS starts from / ends in the patch's record
// (shared by the ci task kernel and the nested solve of the tail kernel)
struct CiPatch { double gb_mol, forc_pbot; };
__device__ __forceinline__ void ci_load(const PhsRec& R, phs::Leaf& L, CiPatch& P) {
  P.gb_mol = R.gb_mol; P.forc_pbot = R.forc_pbot;
  L.c3 = (R.flags & RF_C3) != 0; L.medlyn = (R.flags & RF_MEDLYN) != 0;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    L.vcmax[s] = R.vcmax[s]; L.tpu[s] = R.tpu[s]; L.kp[s] = R.kp[s]; L.lmr[s] = R.lmr[s]; L.je[s] = R.je[s]; L.par[s] = R.par[s];
  }
  L.cp = R.cp; L.kc = R.kc; L.ko = R.ko; L.qe = R.qe; L.theta_cj = R.theta_cj; L.theta_ip = R.theta_ip;
  L.medint = R.medint; L.medslope = R.medslope; L.bbb = R.bbb; L.mbb = R.mbb; L.cair = R.cair; L.oair = R.oair;
  L.rh_can = R.rh_can;
}
__device__ __forceinline__ void ci_begin_from_record(const PhsRec& R, phs::CiLane& C) {
  phs::HybridCarry H;
  H.x1sun = R.x1sun; H.x1sha = R.x1sha; H.bsun = R.bsun; H.bsha = R.bsha; H.b0sun = R.b0sun; H.b0sha = R.b0sha;
  phs::ci_task_begin(C, H);
}
// bottom of the outer pass :4034-4046; returns true when hybrid_PHS is finished (only its epilogue is left)
__device__ __forceinline__ bool ci_end_to_record(PhsRec& R, const phs::CiLane& C, bool bad, bool nb, DevStatus* ds) {
  phs::HybridCarry H;
  H.gs0sun = R.gs0sun; H.gs0sha = R.gs0sha; H.iter1 = R.iter1;
  const bool lastpass = phs::ci_task_end(C, H);
  R.x1sun = H.x1sun; R.x1sha = H.x1sha; R.gs0sun = H.gs0sun; R.gs0sha = H.gs0sha;
  R.b0sun = C.bsun; R.b0sha = C.bsha; R.iter1 = H.iter1;
  R.gs_sun = C.gs_sun; R.gs_sha = C.gs_sha;
  R.o = C.o;
  if (bad) report_failure(ds, R.patch, CTSM_ERR_QUADRATIC, 0);
  if (nb) report_failure(ds, R.patch, CTSM_ERR_BRENT, 0);
  if (lastpass) R.flags |= RF_FINAL;
  return lastpass;
}

// ci tasks: one outer pass of hybrid_PHS (:3925-4046) per task, one ci_func evaluation per scheduler round.
__global__ void __launch_bounds__(TASK_THREADS, CI_MINBLOCKS)
phs_ci_kernel(PhsRec* __restrict__ rec, const int* __restrict__ q_in, const int* __restrict__ n_in, int* __restrict__ head,
              int* __restrict__ q_out, int* __restrict__ n_out, DevStatus* ds) {
  const int n = *n_in;
  const int lane = threadIdx.x & 31;
  int st = LS_IDLE, fi = 0;
  bool exhausted = (n <= 0);
  phs::CiLane C;
  phs::Leaf L;
  CiPatch P;                                            // all ci_func reads of the patch besides the leaf state
  bool bad = false, nb = false;
  for (;;) {
    const unsigned idle = __ballot_sync(FULL, st == LS_IDLE);
    if (!exhausted && (__popc(idle) >= CI_REFILL_MIN)) {
      int base = 0;
      if (lane == 0) base = atomicAdd(head, __popc(idle));
      base = __shfl_sync(FULL, base, 0);
      if (base + __popc(idle) >= n) exhausted = true;
      if (st == LS_IDLE) {
        const int my = base + __popc(idle & ((1u << lane) - 1));
        if (my < n) {
          fi = q_in[my];
          const PhsRec& R = rec[fi];
          if (R.flags & RF_SOLVE) {                     // (nrad < 1: nothing to solve, :3477)
            ci_load(R, L, P);
            ci_begin_from_record(R, C);
            bad = false; nb = false;
            st = LS_RUN;
          }
        }
      }
      continue;
    }
    if (idle == FULL) break;                            // nothing running and the queue is exhausted
    bool push = false;
    if (st == LS_RUN) {
      PhsRec& R = rec[fi];
      if (!phs::ci_step(C, R.br, P, L, &bad, &nb)) {
        (void)ci_end_to_record(R, C, bad, nb, ds);
        push = true;
        st = LS_IDLE;
      }
    }
    const unsigned pm = __ballot_sync(FULL, push);
    queue_push(q_out, n_out, pm, push, fi);
  }
}

// the part of PhotosynthesisHydraulicStress that follows the solve (:3497-3547 night, :3587-3714 day, canopy sums
// :3724-3807) for one patch.  Between passes only what the next pass reads is stored (ALL = false: vegwp, rssun, rssha,
// btran, qflx_tran_veg); the patch's record keeps the state of its last pass, from which canopy_final_kernel writes the
// full set of outputs once (ALL = true) - the reference overwrites them on every pass and the last one survives.
template <bool ALL>
__device__ __forceinline__ void phs_outputs(const CanopyDev& f, const CanopyPrm& prm, const Geo& g, const PhsRec& R, int pp,
                                            DevStatus* ds) {
  const int gg = PF(gridcell) - g.begg0;
  const int ivt = PF(itype);
  const double forc_pbot = R.forc_pbot, cfm = R.cf, gb_mol = R.gb_mol, rb = PF(rb1);
  const bool medlyn = (R.flags & RF_MEDLYN) != 0;
  const double crop = f.pft_crop[ivt];
  const int nrad = PF(nrad);
  const double lmr_z[2] = {R.lmr[0], R.lmr[1]};
  double bsun = 0.0, bsha = 0.0, rs_z[2] = {0.0, 0.0}, psn_z[2] = {0.0, 0.0};
  double wc_z[2] = {0.0, 0.0}, wj_z[2] = {0.0, 0.0}, wp_z[2] = {0.0, 0.0};
  double qflx_tran_veg = PF(qflx_tran_veg);
  if (nrad >= 1) {
    const bool scale_an = (crop == 0.0 || !prm.modifyphoto_and_lmr_forcrop);
    const int near_noon = f.near_local_noon[gg];
    const double gsmin = medlyn ? R.medint : R.bbb;
    phs::CiOut co;
    double gs_mol[2], an[2], ci_z[2];
    bsun = R.bsun; bsha = R.bsha;
    const double bb[2] = {bsun, bsha};
    if (R.flags & RF_NIGHT) {                          // night :3497-3547
      qflx_tran_veg = R.tran;
      const bool pd = f.local_time_lt_noon[gg] != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const double xi = R.xo[i]; PF2(vegwp, i) = xi; if (ALL) PF2(vegwp_pd, i) = pd ? xi : spval; }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        co.ac[s] = co.aj[s] = co.ap[s] = co.ag[s] = 0.0;
        an[s] = scale_an ? 0.0 - bb[s] * lmr_z[s] : 0.0 - lmr_z[s];
        rs_z[s] = fmin(2.e4, 1.0 / (fmax(bb[s] * gsmin, 1.0)) * cfm);
        ci_z[s] = 0.0;
        gs_mol[s] = cfm / rs_z[s];
      }
    } else {                                           // day :3587-3711
      qflx_tran_veg = R.tran;
      co = R.o;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const double xi = R.xo[i]; PF2(vegwp, i) = xi; if (ALL) { PF2(vegwp_ln, i) = near_noon ? xi : spval; PF2(vegwp_pd, i) = spval; } }
      gs_mol[0] = R.gs_sun; gs_mol[1] = R.gs_sha;
      const double cair = R.cair;
      const double o3g[2] = {PF(o3coefgsun), PF(o3coefgsha)}, o3v[2] = {PF(o3coefvsun), PF(o3coefvsha)};
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        an[s] = co.an[s];
        if (an[s] < 0.0) gs_mol[s] = fmax(bb[s] * gsmin, 1.0);
        ci_z[s] = cair - an[s] * forc_pbot * (1.4 * gs_mol[s] + 1.6 * gb_mol) / (gb_mol * gs_mol[s]);
        ci_z[s] = fmax(ci_z[s], 1.e-06);
        const double gs = gs_mol[s] / cfm;
        rs_z[s] = fmin(1.0 / gs, 2.e4);
        rs_z[s] = rs_z[s] / o3g[s];
        psn_z[s] = co.ag[s] * o3v[s];
        if (co.ac[s] <= co.aj[s] && co.ac[s] <= co.ap[s]) wc_z[s] = psn_z[s];
        else if (co.aj[s] < co.ac[s] && co.aj[s] <= co.ap[s]) wj_z[s] = psn_z[s];
        else if (co.ap[s] < co.ac[s] && co.ap[s] < co.aj[s]) wp_z[s] = psn_z[s];
      }
      if (ALL) { PF2(gs_mol_sun_ln, 0) = near_noon ? gs_mol[0] : spval; PF2(gs_mol_sha_ln, 0) = near_noon ? gs_mol[1] : spval; }
      if (!ALL && (gs_mol[0] < 0.0 || gs_mol[1] < 0.0)) report_failure(ds, pp + g.begp0, CTSM_ERR_GS_NEG, 0);
    }
    if (ALL) {
      PF2(ac_phs, 0) = co.ac[0]; PF2(ac_phs, 1) = co.ac[1]; PF2(aj_phs, 0) = co.aj[0]; PF2(aj_phs, 1) = co.aj[1];
      PF2(ap_phs, 0) = co.ap[0]; PF2(ap_phs, 1) = co.ap[1]; PF2(ag_phs, 0) = co.ag[0]; PF2(ag_phs, 1) = co.ag[1];
      PF2(an_sun, 0) = an[0]; PF2(an_sha, 0) = an[1];
      PF2(gs_mol_sun, 0) = gs_mol[0]; PF2(gs_mol_sha, 0) = gs_mol[1];
      PF2(cisun_z, 0) = ci_z[0]; PF2(cisha_z, 0) = ci_z[1];
      PF2(rssun_z, 0) = rs_z[0]; PF2(rssha_z, 0) = rs_z[1];
      PF2(psnsun_z, 0) = psn_z[0]; PF2(psnsha_z, 0) = psn_z[1];
    }
  }
  // canopy sums :3724-3807 (nlevcan = 1)
  {
    const bool scale_lmr = (crop == 0.0 && prm.modifyphoto_and_lmr_forcrop);
    const double lz[2] = {nrad >= 1 ? PF2(laisun_z, 0) : 0.0, nrad >= 1 ? PF2(laisha_z, 0) : 0.0};
    const double bb[2] = {bsun, bsha};
    double psn[2], pwc[2], pwj[2], pwp[2], lmr[2], rs[2], lai[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      double a = 0.0, b_ = 0.0, c_ = 0.0, d_ = 0.0, e_ = 0.0, gsc = 0.0, ll = 0.0;
      if (nrad >= 1) {
        a = a + psn_z[s] * lz[s]; b_ = b_ + wc_z[s] * lz[s]; c_ = c_ + wj_z[s] * lz[s]; d_ = d_ + wp_z[s] * lz[s];
        e_ = scale_lmr ? e_ + lmr_z[s] * lz[s] * bb[s] : e_ + lmr_z[s] * lz[s];
        gsc = gsc + lz[s] / (rb + rs_z[s]);
        ll = ll + lz[s];
      }
      lai[s] = ll;
      if (ll > 0.0) { psn[s] = a / ll; pwc[s] = b_ / ll; pwj[s] = c_ / ll; pwp[s] = d_ / ll; lmr[s] = e_ / ll; rs[s] = ll / gsc - rb; }
      else { psn[s] = 0.0; pwc[s] = 0.0; pwj[s] = 0.0; pwp[s] = 0.0; lmr[s] = 0.0; rs[s] = 0.0; }
    }
    if (ALL) {
      PF(psnsun) = psn[0]; PF(psnsun_wc) = pwc[0]; PF(psnsun_wj) = pwj[0]; PF(psnsun_wp) = pwp[0]; PF(lmrsun) = lmr[0];
      PF(psnsha) = psn[1]; PF(psnsha_wc) = pwc[1]; PF(psnsha_wj) = pwj[1]; PF(psnsha_wp) = pwp[1]; PF(lmrsha) = lmr[1];
    }
    PF(rssun) = rs[0]; PF(rssha) = rs[1];
    double btran;
    if (lai[1] + lai[0] > 0.0) btran = bsun * (lai[0] / (lai[0] + lai[1])) + bsha * (lai[1] / (lai[0] + lai[1]));
    else btran = bsun;
    PF(btran) = btran;
    if (ALL) { PF(bsun) = bsun; PF(bsha) = bsha; }
    PF(qflx_tran_veg) = qflx_tran_veg;
  }
}

// ---------------------------------------------------------------------------------------------
// Tail of the ITERATION loop.  A bulk round is a chain of ~15 grid-wide kernels, each as long as its slowest task, so
// rounds that serve a few thousand patches (the late passes: 1 % of the patches need more than 20 of the 41 passes) and
// tasks that run far longer than the rest (0.5 % of the calcstress solves hit the 50-iteration cap) would set the pace
// of the whole call.  Such patches are EJECTED from the bulk rounds - all survivors once a round's list is short
// (canopy_close_kernel), single patches whose calcstress exceeds its budget (task kernels) - and one thread per patch
// runs each through all of its remaining passes with the nested loops of the reference, concurrently with the bulk
// rounds on a second stream.  Same device functions, same arithmetic, same order as the bulk kernels: results do not
// depend on where a patch was ejected (tests/test_gpu_canopy.py compares bulk-only, tail-only and mixed runs bit for bit).

// the PHS record of a day patch as canopy_leaf_kernel left it (start of hybrid_PHS :3893-3915)
__device__ __forceinline__ void phs_restart(PhsRec& R) {
  int flags = R.flags & ~RF_FINAL;
  R.flags = flags;
  if ((flags & RF_SOLVE) && !(flags & RF_NIGHT)) {
    phs::HybridCarry H;
    phs::hybrid_carry_init(H, ((flags & RF_C3) ? 0.7 : 0.4) * R.cair);
    R.x1sun = H.x1sun; R.x1sha = H.x1sha; R.gs0sun = H.gs0sun; R.gs0sha = H.gs0sha;
    R.bsun = H.bsun; R.bsha = H.bsha; R.b0sun = H.b0sun; R.b0sha = H.b0sha;
    R.iter1 = H.iter1;
  }
}

// calcstress :4490-4710 for one patch, straight from its record
__device__ __noinline__ void calcstress_record(PhsRec& R) {
  phs::Newton N;
  const double xin[4] = {R.x[0], R.x[1], R.x[2], R.x[3]};
  const double gs0sun = R.gs0sun, gs0sha = R.gs0sha;
  if (phs::newton_begin(N, R, xin, gs0sun, gs0sha)) {
    while (phs::newton_step(N, R)) {}
  }
  double tran = 0.0;
  const phs::Stress so = phs::newton_finish(N, R, gs0sun, gs0sha, &tran);
  R.bsun = so.bsun; R.bsha = so.bsha;
  if (R.flags & RF_NIGHT) {
#pragma unroll
    for (int i = 0; i < 4; ++i) R.xo[i] = N.x[i];
    R.tran = tran;
  }
}

// the PHS solve of one pass for one patch (PhotosynthesisMod.F90:3477-3585 -> hybrid_PHS :3815-4064): the sequence the
// task kernels spread over ci(1) -> calcstress -> ci(2) -> ... -> epilogue, as nested loops
__device__ __noinline__ void phs_solve_nested(PhsRec& R, DevStatus* ds) {
  const int flags = R.flags;
  if (!(flags & RF_SOLVE)) return;                       // nrad < 1
  if (flags & RF_NIGHT) { calcstress_record(R); return; }
  phs::Leaf L;
  CiPatch P;
  ci_load(R, L, P);
  for (;;) {
    phs::CiLane C;
    bool bad = false, nb = false;
    ci_begin_from_record(R, C);
    while (phs::ci_step(C, R.br, P, L, &bad, &nb)) {}
    if (ci_end_to_record(R, C, bad, nb, ds)) break;
    calcstress_record(R);
  }
  double x[4];                                           // hybrid_PHS epilogue :4048-4062
  double sf = phs::getvegwp(R, x, R.gs_sun, R.gs_sha);
  if (sf < 0.0) sf = 0.0;
  R.tran = sf;
#pragma unroll
  for (int i = 0; i < 4; ++i) R.xo[i] = x[i];
}

__device__ __noinline__ void fric_call(const CanopyDev& f, const CanopyPrm& prm, const Geo& g, int itlef0, int fi,
                                       const int32_t* filterp, double* ws, int wstride) {
  fric_body(f, prm, g, itlef0, fi, filterp, ws, wstride);
}
__device__ __noinline__ void leaf_call(const CanopyDev& f, const CanopyPrm& prm, const Geo& g, int itlef0, int fi,
                                       const int32_t* filterp, double* ws, int wstride, PhsRec* rec, DevStatus* ds) {
  leaf_body(f, prm, g, itlef0, fi, filterp, ws, wstride, rec, ds);
}
__device__ __noinline__ bool close_call(const CanopyDev& f, const CanopyPrm& prm, const Geo& g, int itlef0, int fi,
                                        const int32_t* filterp, double* ws, int wstride, PhsRec* rec, DevStatus* ds) {
  if (prm.hydrstress) phs_outputs<false>(f, prm, g, rec[fi], filterp[fi] - g.begp0, ds);
  return close_body(f, prm, g, itlef0, false, fi, filterp, ws, wstride, ds).keep;
}

// One launch instead of three for SHORT lists (late rounds: a few thousand patches, every kernel as long as one patch's
// dependent chain): close pass itlef0-1 of a listed patch and, if it survives, open pass itlef0 for it (fric, leaf) in
// the same thread.  Same functions as canopy_close / _fric / _leaf_kernel; patches are independent, so the order
// "all close, all fric, all leaf" against "close, fric, leaf per patch" changes no operand.  PHS configuration only;
// not used together with the tail kernel.
__global__ void __launch_bounds__(STEP_THREADS)
canopy_round_small_kernel(CanopyDev f, CanopyPrm prm, Geo g, int itlef0, const int32_t* __restrict__ filterp,
                          double* __restrict__ ws, int wstride, Lists L, const int* __restrict__ list_in,
                          int* __restrict__ list_out, PhsRec* __restrict__ rec, DevStatus* ds) {
  const int row = itlef0;
  int off[NBIN + 1];
  const int total = list_offsets(L, row, off);
  for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
    const ListSlot sl = list_slot(L, row, off, list_in, base + threadIdx.x);
    const int fi = sl.fi;
    const bool live = sl.live && !(L.ejected[fi]);
    CloseOut c; c.keep = false; c.solve = false; c.night = false;
    if (live) {
      phs_outputs<false>(f, prm, g, rec[fi], filterp[fi] - g.begp0, ds);
      c = close_body(f, prm, g, itlef0, false, fi, filterp, ws, wstride, ds);
    }
    const bool gb = c.keep;
    const unsigned act = __activemask();
    const unsigned mk = __ballot_sync(act, gb);
    if (gb) bin_append(L, list_out, row + 1, 0, fi, mk);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const bool gq = gb && c.solve && (c.night == (q == 0));
      const unsigned mq = __ballot_sync(act, gq);
      if (gq) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(mq) - 1;
        int b0 = 0;
        if (lane == leader) b0 = atomicAdd(q == 0 ? L.n_nt(row + 1, 0) : L.n_ci(row + 1, 0), __popc(mq));
        b0 = __shfl_sync(mq, b0, leader);
        (q == 0 ? L.q_nt : L.q_ci)[b0 + __popc(mq & ((1u << lane) - 1))] = fi;
      }
    }
    if (gb) {
#ifdef ROUND_SMALL_CALLS
      fric_call(f, prm, g, itlef0, fi, filterp, ws, wstride);
      leaf_call(f, prm, g, itlef0, fi, filterp, ws, wstride, rec, ds);
#else
      fric_body(f, prm, g, itlef0, fi, filterp, ws, wstride);
      leaf_body(f, prm, g, itlef0, fi, filterp, ws, wstride, rec, ds);
#endif
    }
  }
}

// one thread: where the tail list stands at the end of bulk round `round`
__global__ void canopy_tail_mark_kernel(Lists L, int round) { L.tail_end[round] = *L.tail_count; }

#define TAIL_THREADS 64
__global__ void __launch_bounds__(TAIL_THREADS)
canopy_tail_kernel(CanopyDev f, CanopyPrm prm, Geo g, int round, int npass, const int32_t* __restrict__ filterp, double* ws,
                   int wstride, Lists L, PhsRec* rec, int lanes, DevStatus* ds) {
  const int beg = round > 0 ? L.tail_end[round - 1] : 0;
  const int n = L.tail_end[round] - beg;
  if (n <= 0) return;
  const int lane = threadIdx.x & 31;
  // a warp carries `lanes` patches (<= 32): every patch takes its own path through the nest, so fewer patches per warp
  // trade idle lanes for less serialisation
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&L.tail_head[round], lanes);
    base = __shfl_sync(FULL, base, 0);
    if (base >= n) break;
    const int t = base + lane;
    if (lane < lanes && t < n) {
      const int e = L.tail_list[beg + t];
      const int fi = e & ~TAIL_B;
      int k = round;
      bool open = !(e & TAIL_B);
      if (!open) phs_restart(rec[fi]);
      for (;;) {
        if (open) {
          fric_call(f, prm, g, k, fi, filterp, ws, wstride);
          leaf_call(f, prm, g, k, fi, filterp, ws, wstride, rec, ds);
        }
        open = true;
        phs_solve_nested(rec[fi], ds);
        ++k;
        if (!close_call(f, prm, g, k, fi, filterp, ws, wstride, rec, ds) || k == npass) break;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// after the ITERATION loop (CanopyFluxesMod.F90:1464-1760)
__global__ void __launch_bounds__(128)
canopy_final_kernel(CanopyDev f, CanopyPrm prm, Geo g, int fn, const int32_t* __restrict__ filterp,
                    const double* __restrict__ ws, int wstride, const PhsRec* __restrict__ rec, DevStatus* ds) {
  const int fi = blockIdx.x * blockDim.x + threadIdx.x;
  if (fi >= fn) return;
  const int pp = filterp[fi] - g.begp0;
  if (prm.hydrstress) phs_outputs<true>(f, prm, g, rec[fi], pp, ds);        // PHS outputs of the patch's last pass
  const int cc = PF(column) - g.begc0;
  const int gg = PF(gridcell) - g.begg0;
  const double dtime = prm.dtime;
  const double forc_rho = CF(forc_rho), forc_q = CF(forc_q), t_grnd = CF(t_grnd), thm = PF(thm);
  const double emv = PF(emv), emg = CF(emg), forc_lwrad = CF(forc_lwrad);
  const int snl = CF(snl);
  const double tsn = CF2(t_soisno, snl + 1 - SNOSOI_LO), ts1 = CF2(t_soisno, 1 - SNOSOI_LO), th2o = CF(t_h2osfc);
  const double frac_sno = CF(frac_sno_eff), frac_h2osfc = CF(frac_h2osfc);
  const double lw_grnd = (frac_sno * pow4(tsn) + (1.0 - frac_sno - frac_h2osfc) * pow4(ts1) + frac_h2osfc * pow4(th2o));
  const double frs = WS(W_FRS), tb = WS(W_TLBEF), dt_veg = WS(W_DT_VEG), tsi = WS(W_TS_INI), tl_ini = WS(W_TL_INI);
  const double air = WS(W_AIR), bir = WS(W_BIR), cir = WS(W_CIR), lw_leaf = WS(W_LW_LEAF), lw_stem = WS(W_LW_STEM);
  const double cp_leaf = WS(W_CP_LEAF), cp_stem = WS(W_CP_STEM);
  const double tb3 = pow3(tb), tsi3 = pow3(tsi), tsi4 = pow4(tsi);
  const double sabv = PF(sabv), tv = PF(t_veg);
  const double qflx_evap_veg = PF(qflx_evap_veg), qflx_tran_veg = PF(qflx_tran_veg);
  const double err = (1.0 - frs) * (sabv + air + bir * tb3 * (tb + 4.0 * dt_veg) + cir * lw_grnd) - lw_leaf + lw_stem
                     - PF(eflx_sh_veg) - hvap * qflx_evap_veg - ((tv - tl_ini) * cp_leaf / dtime);
  double dt_stem = 0.0;
  double tstem = PF(t_stem);
  if (prm.use_biomass_heat_storage) {
    if (PF(stem_biomass) > 0.0)
      dt_stem = (frs * (sabv + air + bir * tsi4 + cir * lw_grnd) - PF(eflx_sh_stem) + lw_leaf - lw_stem)
                / (cp_stem / dtime - frs * bir * 4.0 * tsi3);
    PF(dhsdt_canopy) = dt_stem * cp_stem / dtime + (tv - tl_ini) * cp_leaf / dtime;
    tstem = tstem + dt_stem;
    PF(t_stem) = tstem;
  }
  const double wtal = WS(W_WTAL), wtl0 = WS(W_WTL0), wta0 = WS(W_WTA0), wtstem0 = WS(W_WTSTEM0), wtg = WS(W_WTG);
  const double wtgq = WS(W_WTGQ), wtalq = WS(W_WTALQ), wtlq0 = WS(W_WTLQ0), wtaq0 = WS(W_WTAQ0), qsatl = WS(W_QSATL);
  const double delt = wtal * t_grnd - wtl0 * tv - wta0 * thm - wtstem0 * tstem;
  const double ram1 = PF(ram1);
  PF(taux) = -forc_rho * f.forc_u[gg] / ram1;
  PF(tauy) = -forc_rho * f.forc_v[gg] / ram1;
  PF(eflx_sh_grnd) = cpair * forc_rho * wtg * delt;
  PF(eflx_sh_snow) = cpair * forc_rho * wtg * (wtal * tsn - wtl0 * tv - wta0 * thm - wtstem0 * tstem);
  PF(eflx_sh_soil) = cpair * forc_rho * wtg * (wtal * ts1 - wtl0 * tv - wta0 * thm - wtstem0 * tstem);
  PF(eflx_sh_h2osfc) = cpair * forc_rho * wtg * (wtal * th2o - wtl0 * tv - wta0 * thm - wtstem0 * tstem);
  PF(qflx_evap_soi) = forc_rho * wtgq * WS(W_DELQ);
  PF(qflx_ev_snow) = forc_rho * wtgq * (wtalq * CF(qg_snow) - wtlq0 * qsatl - wtaq0 * forc_q);
  PF(qflx_ev_soil) = forc_rho * wtgq * (wtalq * CF(qg_soil) - wtlq0 * qsatl - wtaq0 * forc_q);
  PF(qflx_ev_h2osfc) = forc_rho * wtgq * (wtalq * CF(qg_h2osfc) - wtlq0 * qsatl - wtaq0 * forc_q);
  const double temp1 = WS(W_TEMP1), temp2 = WS(W_TEMP2);
  const double t_ref2m = thm + temp1 * WS(W_DTH) * (1.0 / WS(W_TEMP12M) - 1.0 / temp1);
  PF(t_ref2m) = t_ref2m; PF(t_ref2m_r) = t_ref2m;
  const double q_ref2m = forc_q + temp2 * WS(W_DQH) * (1.0 / WS(W_TEMP22M) - 1.0 / temp2);
  PF(q_ref2m) = q_ref2m;
  const QS q2 = qsat(t_ref2m, CF(forc_pbot), false);
  const double rh = fmin(100.0, q_ref2m / q2.qs * 100.0);
  PF(rh_ref2m) = rh; PF(rh_ref2m_r) = rh;
  PF(vpd_ref2m) = q2.es * (1.0 - rh / 100.0);
  if (prm.human_fast) {                                   // fast human stress indices :1550-1570 (HumanIndexMod.F90)
    const double tc = t_ref2m - tfrz;                     // KtoC :1205
    PF(tc_ref2m) = tc;
    const double vap = (rh / 100.0) * q2.es;              // VaporPres :1243
    PF(vap_ref2m) = vap;
    if (rh < 0.0 || rh > 100.0) report_failure(ds, pp + g.begp0, CTSM_ERR_RH, 0);       // Wet_BulbS :1016-1022
    const double wbt = tc * atan(0.151977 * sqrt(rh + 8.313659)) + atan(tc + rh) - atan(rh - 1.676331)
                       + 0.00391838 * pow(rh, (3.0 / 2.0)) * atan(0.023101 * rh) - 4.686035;
    PF(wbt_ref2m) = wbt; PF(wbt_ref2m_r) = wbt;
    const double tf = (tc) * 9.0 / 5.0 + 32.0;            // HeatIndex :1039-1095
    double hi;
    if (tf < 68.0) hi = tf;
    else hi = -42.379 + 2.04901523 * tf + 10.14333127 * rh + (-0.22475541 * tf * rh) + (-6.83783e-3 * (tf * tf))
              + (-5.481717e-2 * (rh * rh)) + 1.22874e-3 * (tf * tf) * rh + 8.5282e-4 * tf * (rh * rh)
              + (-1.99e-6 * (tf * tf) * (rh * rh));
    hi = (hi - 32.0) * 5.0 / 9.0;
    PF(nws_hi_ref2m) = hi; PF(nws_hi_ref2m_r) = hi;
    const double at = tc + 3.30 * vap / 1000.0 - 0.70 * PF(u10_clm) - 4.0;              // AppTemp :555
    PF(appar_temp_ref2m) = at; PF(appar_temp_ref2m_r) = at;
    const double sw = 0.567 * (tc) + 0.393 * vap / 100.0 + 3.94;                        // swbgt :596
    PF(swbgt_ref2m) = sw; PF(swbgt_ref2m_r) = sw;
    const double hx = tc + ((5.0 / 9.0) * (vap / 100.0 - 10.0));                        // hmdex :637
    PF(humidex_ref2m) = hx; PF(humidex_ref2m_r) = hx;
    const double Tc = fmin(tc, 50.0);                                                   // dis_coiS :715-761
    double rhl = fmin(rh, 99.0);
    rhl = fmax(rhl, 5.0);
    const double rh_min = Tc * (-2.27) + 27.7;
    const double dc = (Tc < -20.0 || rhl < rh_min) ? Tc : 0.5 * wbt + 0.5 * Tc;
    PF(discomf_index_ref2mS) = dc; PF(discomf_index_ref2mS_r) = dc;
  }
  PF(dlrad) = (1.0 - emv) * emg * forc_lwrad + emv * emg * sb * tb3 * (tb + 4.0 * dt_veg) * (1.0 - frs)
              + emv * emg * sb * tsi3 * (tsi + 4.0 * dt_stem) * frs;
  PF(ulrad) = ((1.0 - emg) * (1.0 - emv) * (1.0 - emv) * forc_lwrad
               + emv * (1.0 + (1.0 - emg) * (1.0 - emv)) * sb * tb3 * (tb + 4.0 * dt_veg) * (1.0 - frs)
               + emv * (1.0 + (1.0 - emg) * (1.0 - emv)) * sb * tsi3 * (tsi + 4.0 * dt_stem) * frs
               + emg * (1.0 - emv) * sb * lw_grnd);
  PF(t_skin) = emv * tv + (1.0 - emv) * sqrt(sqrt(lw_grnd));
  const double cgrnds = PF(cgrnds) + cpair * forc_rho * wtg * wtal;
  const double cgrndl = PF(cgrndl) + forc_rho * wtgq * wtalq * CF(dqgdT);
  PF(cgrnds) = cgrnds; PF(cgrndl) = cgrndl;
  PF(cgrnd) = cgrnds + cgrndl * CF(htvp);
  // dew :1615-1641
  double snocan = PF(snocan), liqcan = PF(liqcan);
  const double base = snocan;
  if (tv > tfrz) {
    if ((qflx_evap_veg - qflx_tran_veg) * dtime > liqcan) snocan = fmax(0.0, snocan + liqcan + (qflx_tran_veg - qflx_evap_veg) * dtime);
    liqcan = fmax(0.0, liqcan + (qflx_tran_veg - qflx_evap_veg) * dtime);
  } else if (tv <= tfrz) {
    if ((qflx_evap_veg - qflx_tran_veg) * dtime > snocan) liqcan = liqcan + snocan + (qflx_tran_veg - qflx_evap_veg) * dtime;
    snocan = fmax(0.0, snocan + (qflx_tran_veg - qflx_evap_veg) * dtime);
  }
  if (fabs(snocan) < 1.e-10 * fabs(base)) snocan = 0.0;
  PF(snocan) = snocan; PF(liqcan) = liqcan;
  // PhotosynthesisTotal :2125-2131, iwue :1661-1676
  const double laisun = PF(laisun), laisha = PF(laisha);
  const double fpsn = PF(psnsun) * laisun + PF(psnsha) * laisha;
  PF(fpsn) = fpsn;
  PF(fpsn_wc) = PF(psnsun_wc) * laisun + PF(psnsha_wc) * laisha;
  PF(fpsn_wj) = PF(psnsun_wj) * laisun + PF(psnsha_wj) * laisha;
  PF(fpsn_wp) = PF(psnsun_wp) * laisun + PF(psnsha_wp) * laisha;
  double iwue = spval;
  if (f.near_local_noon[gg] && fpsn > 0.0) {
    const double gs = 1.e-6 * (laisun * PF2(gs_mol_sun, 0) + laisha * PF2(gs_mol_sha, 0));
    if (gs > 0.0) iwue = fpsn / gs;
  }
  PF(iwue_ln) = iwue;
  if (prm.use_luna) {                                     // Acc24_Climate_LUNA, LunaMod.F90:695-724 (call :1704)
    const double tvd = PF(t_veg_day);
    if (tvd != spval) {                                   // not the first day
      if (sabv > 0) { PF(t_veg_day) = tvd + tv; PF(ndaysteps) = PF(ndaysteps) + 1; }
      else { PF(t_veg_night) = PF(t_veg_night) + tv; PF(nnightsteps) = PF(nnightsteps) + 1; }
      if (PF(nrad) >= 1) {                                // nlevcan = 1
        const double tlaii = PF2(laisun_z, 0) + PF2(laisha_z, 0);
        if (tlaii > 0.0) {
          const double TRad = PF2(parsun_z, 0);           // :715 overrides the lai-weighted mean of :714
          PF2(par24d_z, 0) = PF2(par24d_z, 0) + dtime * TRad;
          if (TRad > PF2(par24x_z, 0)) PF2(par24x_z, 0) = TRad;
        }
      }
      PF(fpsn24) = PF(fpsn24) + dtime * fpsn;
    }
  }
  if (fabs(err) > 0.1) atomicAdd(&ds->n_warnings, 1);       // :1746-1760
}

// ---------------------------------------------------------------------------------------------
// setExposedvegpFilter (filterMod.F90:595-648): order-preserving two-way split
#define SPLIT_BLOCK 256
#define SPLIT_ITEMS 8
__global__ void __launch_bounds__(SPLIT_BLOCK)
split_count_kernel(int n, const int32_t* __restrict__ filt, const int32_t* __restrict__ fv, int begp, int sign, int* __restrict__ blockc) {
  __shared__ int sh[SPLIT_BLOCK / 32];
  const int base = blockIdx.x * SPLIT_BLOCK * SPLIT_ITEMS + threadIdx.x * SPLIT_ITEMS;
  int c = 0;
  for (int k = 0; k < SPLIT_ITEMS; ++k) {
    const int i = base + k;
    if (i < n && sign * fv[filt[i] - begp] > 0) ++c;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int k = 0; k < SPLIT_BLOCK / 32; ++k) s += sh[k];
    blockc[blockIdx.x] = s;
  }
}
__global__ void split_scan_kernel(int nblocks, int* __restrict__ blockc, int* __restrict__ total) {
  // single block: exclusive scan of the per-block counts (nblocks is small: n / 2048)
  __shared__ int carry;
  __shared__ int sh[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < nblocks) ? blockc[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) blockc[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(SPLIT_BLOCK)
split_scatter_kernel(int n, const int32_t* __restrict__ filt, const int32_t* __restrict__ fv, int begp, int sign,
                     const int* __restrict__ blockc, int32_t* __restrict__ out_yes, int32_t* __restrict__ out_no) {
  __shared__ int sh[SPLIT_BLOCK];
  const int base = blockIdx.x * SPLIT_BLOCK * SPLIT_ITEMS + threadIdx.x * SPLIT_ITEMS;
  int c = 0;
  int32_t v[SPLIT_ITEMS];
  bool y[SPLIT_ITEMS];
  for (int k = 0; k < SPLIT_ITEMS; ++k) {
    const int i = base + k;
    v[k] = (i < n) ? filt[i] : 0;
    y[k] = (i < n) && sign * fv[v[k] - begp] > 0;
    if (y[k]) ++c;
  }
  sh[threadIdx.x] = c;
  __syncthreads();
  for (int o = 1; o < SPLIT_BLOCK; o <<= 1) {
    const int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  int yes_pos = blockc[blockIdx.x] + sh[threadIdx.x] - c;           // exposed entries before this thread
  for (int k = 0; k < SPLIT_ITEMS; ++k) {
    const int i = base + k;
    if (i >= n) break;
    if (y[k]) { out_yes[yes_pos] = v[k]; ++yes_pos; }
    else out_no[i - yes_pos] = v[k];                                 // entries before i that are not exposed
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
static int reserve_ints(ctsm_b200_ctx* ctx, ctsm_b200_ctx::Arena& a, size_t n) { return arena_reserve(ctx, a, sizeof(int) * (n > 0 ? n : 1)); }

extern "C" int ctsm_b200_canopyfluxes(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_exposedvegp,
                                      const int32_t* filter_exposedvegp, const ctsm_canopyfluxes_fields_t* hf, int mem,
                                      ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_exposedvegp < 0 || (num_exposedvegp > 0 && !filter_exposedvegp)) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  CanopyDev d;
  const int32_t* dfilter = filter_exposedvegp;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_CANOPYFLUXES
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_CANOPYFLUXES
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_exposedvegp, num_exposedvegp, &dfilter);
    if (rc) return rc;
  }
  const ctsm_params_t& p = ctx->prm;
  CanopyPrm cp;
  cp.dtime = p.dtime; cp.itmax = p.itmax_canopy_fluxes; cp.use_undercanopy_stability = p.use_undercanopy_stability;
  cp.use_biomass_heat_storage = p.use_biomass_heat_storage; cp.z0param_method = p.z0param_method;
  cp.soil_resis_method = p.soil_resis_method; cp.use_luna = p.use_luna; cp.medlyn = (p.stomatalcond_mtd == 2);
  cp.light_inhibit = p.light_inhibit; cp.modifyphoto_and_lmr_forcrop = p.modifyphoto_and_lmr_forcrop;
  cp.human_fast = (p.calc_human_stress_indices == 1);
  cp.hydrstress = p.use_hydrstress;
  cp.lai_dl = p.lai_dl; cp.z_dl = p.z_dl; cp.a_coef = p.a_coef; cp.a_exp = p.a_exp; cp.csoilc = p.csoilc; cp.cv = p.cv;
  cp.wind_min = p.wind_min; cp.zetamaxstable = p.zetamaxstable; cp.leaf_mr_vcm = p.leaf_mr_vcm;
  cp.act25 = p.act25; cp.fnr = p.fnr; cp.cp25_yr2000 = p.cp25_yr2000; cp.kc25_coef = p.kc25_coef; cp.ko25_coef = p.ko25_coef;
  cp.fnps = p.fnps; cp.theta_psii = p.theta_psii; cp.theta_ip = p.theta_ip;
  cp.vcmaxha = p.vcmaxha; cp.jmaxha = p.jmaxha; cp.tpuha = p.tpuha; cp.lmrha = p.lmrha; cp.kcha = p.kcha; cp.koha = p.koha;
  cp.cpha = p.cpha; cp.vcmaxhd = p.vcmaxhd; cp.jmaxhd = p.jmaxhd; cp.tpuhd = p.tpuhd; cp.lmrhd = p.lmrhd; cp.lmrse = p.lmrse;
  cp.tpu25ratio = p.tpu25ratio; cp.kp25ratio = p.kp25ratio; cp.vcmaxse_sf = p.vcmaxse_sf; cp.jmaxse_sf = p.jmaxse_sf;
  cp.tpuse_sf = p.tpuse_sf; cp.jmax25top_sf = p.jmax25top_sf;
  cp.m_csoilc = ctx->member.csoilc; cp.m_cv = ctx->member.cv; cp.m_a_coef = ctx->member.a_coef; cp.m_z_dl = ctx->member.z_dl;

  Geo g;
  g.begp0 = hf->alloc.begp; g.begc0 = hf->alloc.begc; g.begg0 = hf->alloc.begg;
  g.ldp = hf->alloc.endp - hf->alloc.begp + 1; g.ldc = hf->alloc.endc - hf->alloc.begc + 1;
  g.begp = bounds->begp; g.endp = bounds->endp; g.begc = bounds->begc; g.endc = bounds->endc;
  g.npft = p.npft_table;
  const int fn = num_exposedvegp;
  const int npb = g.endp - g.begp + 1, ncb = g.endc - g.begc + 1;
  if (npb <= 0) return finish_call(ctx, mem, st);

  // workspace: [W_NSLOT][wstride] doubles, the PHS records [fn], and int scratch {colflag[ldc],
  // list_a/list_b[fn], ci queues 0..3 [fn], calcstress queues 0..3 [fn], ejected[fn], tail_list[fn], counters}
  const int wstride = (fn + 31) & ~31;
  const int npass = p.itmax_canopy_fluxes + 1;
  {
    const int rce = ensure_round_events(ctx, npass + 2);
    if (rce) return rce;
  }
  const size_t n_counts = (size_t)QROW * (size_t)(npass + 2);
  const size_t n_tailc = 2 * (size_t)(npass + 2) + 2;        // tail_end[], tail_head[], tail_count
  const size_t ws_bytes = (sizeof(double) * (size_t)W_NSLOT * (size_t)(wstride > 0 ? wstride : 32) + 127) & ~(size_t)127;
  const size_t rec_bytes = (sizeof(PhsRec) * (size_t)(fn > 0 ? fn : 1) + 127) & ~(size_t)127;
  const bool nt_split = ctx->tune.nt_split != 0 && fn > QUAD_MAX;
  const size_t nts_bytes = nt_split ? ((sizeof(double) * 10 + sizeof(int) * 2) * (size_t)fn + 256) : 0;   // NtState + run list
  int rc = arena_reserve(ctx, ctx->arena_scratch, ws_bytes + rec_bytes + nts_bytes);
  if (rc) return rc;
  const size_t nq = (size_t)(2 * NBIN + NQ_CI + NQ_NT + 2);
  rc = reserve_ints(ctx, ctx->arena_ints, (size_t)g.ldc + nq * (size_t)fn + n_counts + n_tailc + 64);
  if (rc) return rc;
  double* ws = (double*)ctx->arena_scratch.p;
  PhsRec* rec = (PhsRec*)((char*)ctx->arena_scratch.p + ws_bytes);
  NtState nts;
  nts.v = (double*)((char*)ctx->arena_scratch.p + ws_bytes + rec_bytes);
  nts.w = (int*)(nts.v + (size_t)10 * fn);
  nts.cap = fn;
  int* run_list = nts.w + fn;
  int* ip = (int*)ctx->arena_ints.p;
  Lists L;
  L.colflag = ip; ip += g.ldc;
  L.list_a = ip; ip += (size_t)NBIN * fn;
  L.list_b = ip; ip += (size_t)NBIN * fn;
  L.q_ci = ip; ip += (size_t)NQ_CI * fn;
  L.q_nt = ip; ip += (size_t)NQ_NT * fn;
  L.tail_list = ip; ip += fn;
  L.ejected = ip; ip += fn;                                  // zeroed together with the counters that follow it
  L.counts = ip; ip += n_counts;
  L.tail_end = ip; ip += npass + 2;
  L.tail_head = ip; ip += npass + 2;
  L.tail_count = ip; ip += 2;
  L.cap = fn;
  ctx->dbg_counts = L.counts; ctx->dbg_tail_end = L.tail_end; ctx->dbg_npass = npass;
  cudaStream_t s = ctx->stream;
  CUDA_TRY(cudaMemsetAsync(L.colflag, 0, sizeof(int) * (size_t)g.ldc, s));
  CUDA_TRY(cudaMemsetAsync(L.ejected, 0, sizeof(int) * ((size_t)fn + n_counts + n_tailc), s));
  if (fn > 0) {
    canopy_mark_kernel<<<grid_for(fn, 256), 256, 0, s>>>(d, g, fn, dfilter, L.colflag);
    canopy_colprep_kernel<<<grid_for(ncb, 128), 128, 0, s>>>(d, g, L.colflag);
    ctx->launches += 2;
  }
  canopy_zero_kernel<<<grid_for(npb, 128), 128, 0, s>>>(d, g);
  ctx->launches++;
  if (fn > 0) {
    canopy_init_kernel<<<grid_for(fn, 128), 128, 0, s>>>(d, cp, g, fn, dfilter, ws, wstride, L, rec, ctx->d_status);
    ctx->launches++;
    const size_t shbytes = sizeof(double) * 2 * NLEVSOI * TASK_THREADS;
    int sms = 148, occ_n = 1, occ_c = 1, occ_s = 1, occ_t = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    CUDA_TRY(cudaFuncSetAttribute(phs_newton_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shbytes));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_n, phs_newton_kernel, TASK_THREADS, shbytes);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, phs_ci_kernel, TASK_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, canopy_close_kernel, STEP_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, canopy_tail_kernel, TAIL_THREADS, 0);
    int occ_i = 1;
    CUDA_TRY(cudaFuncSetAttribute(nt_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shbytes));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_i, nt_iter_kernel, TASK_THREADS, shbytes);
    if (occ_i < 1) occ_i = 1;
    if (occ_n < 1) occ_n = 1;
    if (occ_c < 1) occ_c = 1;
    if (occ_s < 1) occ_s = 1;
    if (occ_t < 1) occ_t = 1;
    // persistent grids: whole waves of resident blocks on the 148 SMs
    const int need_t = grid_for(fn, TASK_THREADS);
    const int grid_n = need_t < sms * occ_n ? need_t : sms * occ_n;
    const int grid_c = need_t < sms * occ_c ? need_t : sms * occ_c;
    const int grid_i = need_t < sms * occ_i ? need_t : sms * occ_i;
    const int need_u = grid_for(fn, 128);
    const int grid_u = need_u < sms * 16 ? need_u : sms * 16;
    const int need_q = grid_for((fn < QUAD_MAX ? fn : QUAD_MAX) * 4, 128);
    const int grid_q = need_q < sms * 4 ? need_q : sms * 4;
    int grid_s = grid_for(fn + 32 * NBIN, STEP_THREADS);
    if (grid_s > sms * occ_s * 2) grid_s = sms * occ_s * 2;
    // tail policy (ctsm_b200_set_tuning / CTSM_B200_TAIL_MAX, _NT_BUDGET, _TAIL_LANES): see canopy_tail_kernel
    const int tail_max = ctx->tune.tail_max, nt_budget = ctx->tune.nt_budget > 0 ? ctx->tune.nt_budget : (1 << 30);
    const int tail_lanes = ctx->tune.tail_lanes < 1 ? 1 : ctx->tune.tail_lanes > 32 ? 32 : ctx->tune.tail_lanes;
    const bool use_tail = tail_max > 0 || ctx->tune.nt_budget > 0;
    const int tail_cap = fn < tail_max + (fn >> 6) + 1024 ? fn : tail_max + (fn >> 6) + 1024;   // what one round is expected to eject
    const int need_tail = grid_for(tail_cap, (TAIL_THREADS / 32) * tail_lanes);
    const int grid_tail = need_tail < sms * occ_t ? need_tail : sms * occ_t;
    cudaStream_t s2 = ctx->stream2;
    int *lin = L.list_a, *lout = L.list_b;
    const size_t cap = (size_t)fn;
    // close(0, first) only builds the list of pass 0; fric/leaf(k) open pass k; the task kernels solve its PHS system;
    // close(k+1) closes pass k; close(npass, last) only closes.  The host stops issuing rounds once it has seen (two
    // rounds late, through a pinned copy) that a round's list was empty.
    // Inside a resident window only small calls do this: a host that waits on the device could not queue the next clump's
    // uploads meanwhile (and a large filter practically always holds a patch that uses all itmax+1 passes).
    const bool early_exit = fn <= 65536 || !ctx->window_open;
    int rounds = 0;
    bool small_queues = false;      // every calcstress queue of the round is known to be below the split path's threshold
    for (int itlef = 0; itlef <= npass; ++itlef) {
      if (itlef >= 2 && early_exit) {
        // the host stays at most two rounds ahead of the device, so that it can stop when the lists have run empty and
        // leave out the split-path kernels (three launches per calcstress stage that would return at once) when the list -
        // an upper bound of every queue of the round, and lists only shrink - is below their threshold
        CUDA_TRY(cudaEventSynchronize(ctx->ev_round[itlef - 2]));
        if (ctx->h_counts[itlef - 2] == 0) break;
        small_queues = ctx->h_counts[itlef - 2] <= QUAD_MAX;
      }
      // short lists of the PHS configuration: close + fric + leaf in one launch (canopy_round_small_kernel)
      const bool fused_round = small_queues && cp.hydrstress && !use_tail && itlef > 0 && itlef < npass;
      if (fused_round)
        canopy_round_small_kernel<<<grid_s, STEP_THREADS, 0, s>>>(d, cp, g, itlef, dfilter, ws, wstride, L, lin, lout, rec, ctx->d_status);
      else
        canopy_close_kernel<<<grid_s, STEP_THREADS, 0, s>>>(d, cp, g, itlef, itlef == 0, itlef == npass, dfilter, ws, wstride,
                                                            L, lin, lout, rec, tail_max, ctx->d_status);
      ctx->launches++;
      rounds = itlef + 1;
      if (itlef < npass) {
        // the list this round works on: copy its length to the host for the early exit above
        if (early_exit) {
          CUDA_TRY(cudaMemcpyAsync(&ctx->h_counts[itlef], L.counts + (size_t)(itlef + 1) * QROW, sizeof(int), cudaMemcpyDeviceToHost, s));
          CUDA_TRY(cudaEventRecord(ctx->ev_round[itlef], s));
        }
        if (!fused_round) canopy_fric_kernel<<<grid_s, STEP_THREADS, 0, s>>>(d, cp, g, itlef, dfilter, ws, wstride, L, lout);
        if (!cp.hydrstress) {
          // use_hydrstress = .false.: Photosynthesis for the sunlit, then the shaded leaves (no ci / calcstress task kernels)
          canopy_photosyn_kernel<<<grid_s, STEP_THREADS, 0, s>>>(d, cp, g, itlef, dfilter, ws, wstride, L, lout, ctx->d_status);
          ctx->launches += 2;
          int* t = lin; lin = lout; lout = t;
          continue;
        }
        if (!fused_round) {
          canopy_leaf_kernel<<<grid_s, STEP_THREADS, 0, s>>>(d, cp, g, itlef, dfilter, ws, wstride, L, lout, rec, ctx->d_status);
          ctx->launches += 2;
        }
        const int row = itlef + 1;
        int* crow = L.counts + (size_t)row * QROW;
        int* n_ci = crow + NBIN;
        int* n_nt = n_ci + NQ_CI;
        int* h_ci = n_nt + NQ_NT;
        int* h_nt = h_ci + NQ_CI;
        int* spare = h_nt + NQ_NT;
        for (int i = 0; i < NQ_CI; ++i) {
          // ci(i) hands every patch to calcstress queue i: on to outer pass i + 2, or (solve finished) to the epilogue
          phs_ci_kernel<<<grid_c, TASK_THREADS, 0, s>>>(rec, L.q_ci + (size_t)i * cap, n_ci + i, h_ci + i,
                                                        L.q_nt + (size_t)i * cap, n_nt + i, ctx->d_status);
          // newton(i) feeds ci queue i + 1; newton(3) holds epilogue tasks only and pushes nothing
          const bool lastq = (i + 1 == NQ_CI);
          int* qo = L.q_ci + (size_t)(lastq ? 0 : i + 1) * cap;
          int* no = lastq ? spare : n_ci + i + 1;
          if (small_queues) {
            // (nothing: the four-lane kernel below takes the whole queue)
          } else if (nt_split) {
            // large queues: prologue / iterations / epilogue as three kernels (see NtState); each exits at once on a small queue
            int* n_run = spare + 1 + i;
            int* h_run = spare + 1 + NQ_NT + i;
            nt_begin_kernel<<<grid_u, 128, 0, s>>>(rec, L.q_nt + (size_t)i * cap, n_nt + i, nts, run_list, n_run);
            nt_iter_kernel<<<grid_i, TASK_THREADS, shbytes, s>>>(rec, L.q_nt + (size_t)i * cap, n_nt + i, run_list, n_run, h_run, nts, L,
                                                                 nt_budget);
            nt_finish_kernel<<<grid_u, 128, 0, s>>>(rec, L.q_nt + (size_t)i * cap, n_nt + i, nts, qo, no);
            ctx->launches += 3;
          } else {
            phs_newton_kernel<<<grid_n, TASK_THREADS, shbytes, s>>>(rec, L.q_nt + (size_t)i * cap, n_nt + i, h_nt + i, qo, no, L, nt_budget);
            ctx->launches++;
          }
          phs_newton_quad_kernel<<<grid_q, 128, 0, s>>>(rec, L.q_nt + (size_t)i * cap, n_nt + i, qo, no, L, nt_budget);
          ctx->launches += 2;
        }
        if (use_tail) {
          // everything ejected during this round (survivors of a short list, calcstress stragglers) runs to completion
          // on the second stream while the next bulk rounds proceed
          canopy_tail_mark_kernel<<<1, 1, 0, s>>>(L, itlef);
          CUDA_TRY(cudaEventRecord(ctx->ev_tail[itlef], s));
          CUDA_TRY(cudaStreamWaitEvent(s2, ctx->ev_tail[itlef], 0));
          canopy_tail_kernel<<<grid_tail, TAIL_THREADS, 0, s2>>>(d, cp, g, itlef, npass, dfilter, ws, wstride, L, rec, tail_lanes,
                                                                ctx->d_status);
          ctx->launches += 2;
        }
      }
      int* t = lin; lin = lout; lout = t;
    }
    if (use_tail && rounds > 0) {
      CUDA_TRY(cudaEventRecord(ctx->ev_join, s2));
      CUDA_TRY(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    }
    canopy_final_kernel<<<grid_for(fn, 128), 128, 0, s>>>(d, cp, g, fn, dfilter, ws, wstride, rec, ctx->d_status);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

// list length and cumulative tail entries of every round of the last CanopyFluxes call (diagnostic; synchronises)
extern "C" int ctsm_b200_canopy_round_stats(ctsm_b200_ctx* ctx, int32_t* list_len, int32_t* tail_end, int cap) {
  if (!ctx || !list_len || !tail_end) return -1;
  if (!ctx->dbg_counts) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  const int n = ctx->dbg_npass + 1 < cap ? ctx->dbg_npass + 1 : cap;
  std::vector<int> c((size_t)QROW * (size_t)(n + 1)), t((size_t)n);
  cudaMemcpy(c.data(), ctx->dbg_counts, sizeof(int) * c.size(), cudaMemcpyDeviceToHost);
  cudaMemcpy(t.data(), ctx->dbg_tail_end, sizeof(int) * t.size(), cudaMemcpyDeviceToHost);
  for (int r = 0; r < n; ++r) { list_len[r] = c[(size_t)(r + 1) * QROW]; tail_end[r] = t[r]; }
  return n;
}

// order-preserving split of a filter by the sign of an integer flag array (flag(beg:end), sign * flag > 0 goes to `yes`)
static int split_filter(ctsm_b200_ctx* ctx, int beg, int end, int sign, int num_nolakeurbanp, const int32_t* filter_nolakeurbanp,
                        const int32_t* frac_veg_nosno, int32_t* filter_exposedvegp, int32_t* num_exposedvegp,
                        int32_t* filter_noexposedvegp, int32_t* num_noexposedvegp, int mem);

extern "C" int ctsm_b200_set_exposedvegp_filter(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakeurbanp,
                                                const int32_t* filter_nolakeurbanp, const int32_t* frac_veg_nosno,
                                                int32_t* filter_exposedvegp, int32_t* num_exposedvegp,
                                                int32_t* filter_noexposedvegp, int32_t* num_noexposedvegp, int mem) {
  if (!bounds) return CTSM_ERR_BAD_ARG;
  return split_filter(ctx, bounds->begp, bounds->endp, 1, num_nolakeurbanp, filter_nolakeurbanp, frac_veg_nosno, filter_exposedvegp,
                      num_exposedvegp, filter_noexposedvegp, num_noexposedvegp, mem);
}

// BuildSnowFilter, SnowHydrologyMod.F90:3975-4010: snowc where col%snl < 0
extern "C" int ctsm_b200_build_snow_filter(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_nolakec, const int32_t* filter_nolakec,
                                           const int32_t* snl, int alloc_begc, int alloc_endc, int32_t* filter_snowc, int32_t* num_snowc,
                                           int32_t* filter_nosnowc, int32_t* num_nosnowc, int mem) {
  if (!ctx || !bounds || !snl || alloc_endc < alloc_begc) return CTSM_ERR_BAD_ARG;
  // Inside a resident window the host copy of col%snl may be stale (the snow-layer update has written it on the device and its
  // download is still in flight): where the device mirror of the array is fresh over the call bounds, the split reads the mirror.
  if (mem != CTSM_MEM_DEVICE && ctx->window_open) {
    auto it = ctx->mirrors.find((const void*)snl);
    if (it != ctx->mirrors.end() && it->second.d) {
      bool covered = false;
      for (const auto& iv : it->second.fresh)
        if (iv.first <= bounds->begc && iv.second >= bounds->endc) { covered = true; break; }
      if (covered)
        return split_filter(ctx, alloc_begc, alloc_endc, -1, num_nolakec, filter_nolakec, (const int32_t*)it->second.d, filter_snowc,
                            num_snowc, filter_nosnowc, num_nosnowc, CTSM_MEM_HOST | 0x100);
    }
  }
  return split_filter(ctx, alloc_begc, alloc_endc, -1, num_nolakec, filter_nolakec, snl, filter_snowc, num_snowc, filter_nosnowc,
                      num_nosnowc, mem);
}

static int split_filter(ctsm_b200_ctx* ctx, int beg, int end, int sign, int num_nolakeurbanp, const int32_t* filter_nolakeurbanp,
                        const int32_t* frac_veg_nosno, int32_t* filter_exposedvegp, int32_t* num_exposedvegp,
                        int32_t* filter_noexposedvegp, int32_t* num_noexposedvegp, int mem) {
  if (!ctx || num_nolakeurbanp < 0 || !frac_veg_nosno || !num_exposedvegp || !num_noexposedvegp ||
      (num_nolakeurbanp > 0 && (!filter_nolakeurbanp || !filter_exposedvegp || !filter_noexposedvegp)))
    return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  const bool flag_on_device = (mem & 0x100) != 0;            // (host lists, device flags: ctsm_b200_build_snow_filter inside a window)
  mem &= 0xff;
  const int n = num_nolakeurbanp;
  *num_exposedvegp = 0; *num_noexposedvegp = 0;
  if (n == 0) return CTSM_OK;
  cudaStream_t s = ctx->stream;
  const int np = flag_on_device ? 0 : end - beg + 1;
  const int nblocks = grid_for(n, SPLIT_BLOCK * SPLIT_ITEMS);
  const int32_t *dfilt = filter_nolakeurbanp, *dfv = frac_veg_nosno;
  int32_t *dyes = filter_exposedvegp, *dno = filter_noexposedvegp;
  int rc = reserve_ints(ctx, ctx->arena_ints, (size_t)nblocks + 8 + (mem != CTSM_MEM_DEVICE ? 3 * (size_t)n + (size_t)np : 0));
  if (rc) return rc;
  int* ip = (int*)ctx->arena_ints.p;
  int* blockc = ip; ip += nblocks;
  int* total = ip; ip += 8;
  if (mem != CTSM_MEM_DEVICE) {
    int32_t* a = ip; ip += n;
    int32_t* b = ip; ip += n;
    int32_t* c = ip; ip += n;
    int32_t* e = ip; ip += np;
    CUDA_TRY(cudaMemcpyAsync(a, filter_nolakeurbanp, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
    if (!flag_on_device) CUDA_TRY(cudaMemcpyAsync(e, frac_veg_nosno, sizeof(int32_t) * (size_t)np, cudaMemcpyHostToDevice, s));
    dfilt = a; dyes = b; dno = c; dfv = flag_on_device ? frac_veg_nosno : e;
  }
  split_count_kernel<<<nblocks, SPLIT_BLOCK, 0, s>>>(n, dfilt, dfv, beg, sign, blockc);
  split_scan_kernel<<<1, 1024, 0, s>>>(nblocks, blockc, total);
  split_scatter_kernel<<<nblocks, SPLIT_BLOCK, 0, s>>>(n, dfilt, dfv, beg, sign, blockc, dyes, dno);
  ctx->launches += 3;
  int htotal = 0;
  CUDA_TRY(cudaMemcpyAsync(&htotal, total, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  *num_exposedvegp = htotal;
  *num_noexposedvegp = n - htotal;
  if (mem != CTSM_MEM_DEVICE) {
    if (htotal > 0) CUDA_TRY(cudaMemcpyAsync(filter_exposedvegp, dyes, sizeof(int32_t) * (size_t)htotal, cudaMemcpyDeviceToHost, s));
    if (n - htotal > 0) CUDA_TRY(cudaMemcpyAsync(filter_noexposedvegp, dno, sizeof(int32_t) * (size_t)(n - htotal), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
  }
  return CTSM_OK;
}
