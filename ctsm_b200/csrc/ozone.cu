// ozone.cu — CalcOzoneUptake (OzoneMod.F90:356-511; called inside CanopyFluxes, CanopyFluxesMod.F90:1690) and CalcOzoneStress
// (:514-782; clm_driver.F90:690) on B200.  SURVEY.md section 8f rank 4.  Per-patch maps over the exposed-vegetation filter, one thread
// per filter entry, HBM-bound (11 doubles in, 3 out per patch for the uptake; 2 in, 4 out for the stress).
#include "common.cuh"
#include <vector>

struct OzoneDev {
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) ctype* name;
#define CTSM_FIELDS_OZONE
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_OZONE
#undef CTSM_F
};

namespace {
constexpr double ko3 = 1.67, lai_thresh = 0.5, o3_flux_threshold = 0.8;                        // OzoneMod.F90:95-101
constexpr double SHR_CONST_RGAS = 6.02214e26 * 1.38065e-23;

// CalcOzoneUptakeOnePoint :470-509
__device__ __forceinline__ double uptake_one_point(double forc_ozone, double forc_pbot, double forc_th, double rs, double rb, double ram,
                                                   double tlai, double tlai_old, double evergreen, double leaf_long, int dtime,
                                                   double o3uptake) {
  const double o3concnmolm3 = forc_ozone * 1.e9 * (forc_pbot / (forc_th * SHR_CONST_RGAS * 0.001));
  const double o3flux = o3concnmolm3 / (ko3 * rs + rb + ram);
  const double o3fluxcrit = (o3flux < o3_flux_threshold) ? 0.0 : o3flux - o3_flux_threshold;
  const double dtimeh = dtime / 3600.0;
  const double o3fluxperdt = o3fluxcrit * dtime * 0.000001;
  if (tlai > lai_thresh) {
    const double heal = (tlai - tlai_old > 0) ? fmax(0.0, (((tlai - tlai_old) / tlai) * o3fluxperdt)) : 0.0;
    const double leafturn = (evergreen == 1) ? 1.0 / (leaf_long * 365.0 * 24.0) : 0.0;
    const double decay = o3uptake * leafturn * dtimeh;
    return fmax(0.0, o3uptake + o3fluxperdt - decay - heal);
  }
  return 0.0;
}

__global__ void __launch_bounds__(256)
ozone_uptake_kernel(OzoneDev f, int begp0, int begc0, int begg0, int dtime, int numf, const int32_t* __restrict__ filterp) {
  const int fp = blockIdx.x * blockDim.x + threadIdx.x;
  if (fp >= numf) return;
  const int pp = filterp[fp] - begp0;
  const int c = f.column[pp] - begc0, g = f.gridcell[pp] - begg0, t = f.itype[pp];
  const double o3 = f.forc_o3[g], pbot = f.forc_pbot[c], th = f.forc_th[c], rb = f.rb1[pp], ram = f.ram1[pp], tlai = f.tlai[pp];
  const double tlai_old = f.tlai_old[pp], eg = f.pft_evergreen[t], ll = f.pft_leaf_long[t];
  f.o3uptakesha[pp] = uptake_one_point(o3, pbot, th, f.rssha[pp], rb, ram, tlai, tlai_old, eg, ll, dtime, f.o3uptakesha[pp]);
  f.o3uptakesun[pp] = uptake_one_point(o3, pbot, th, f.rssun[pp], rb, ram, tlai, tlai_old, eg, ll, dtime, f.o3uptakesun[pp]);
  f.tlai_old[pp] = tlai;
}

// the intercept / slope tables :103-133: needleleaf (pft_type <= 3), broadleaf (woody), nonwoody
__constant__ double photoInt[3] = {0.8390, 0.8752, 0.8021}, photoSlope[3] = {0.0, 0.0, -0.0009};
__constant__ double condInt[3] = {0.7823, 0.9125, 0.7511}, condSlope[3] = {0.0048, 0.0, 0.0};
__constant__ double jmaxInt[3] = {1.0, 1.0, 1.0}, jmaxSlope[3] = {0.0, -0.0037, 0.0};

// threads [0, num_exposed) exposed patches, the rest patches without exposed vegetation (coefficients = 1)
__global__ void __launch_bounds__(256)
ozone_stress_kernel(OzoneDev f, int begp0, int method, int num_exposed, const int32_t* __restrict__ filter_exposed, int num_noexposed,
                    const int32_t* __restrict__ filter_noexposed) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_exposed + num_noexposed) return;
  if (t >= num_exposed) {
    const int pp = filter_noexposed[t - num_exposed] - begp0;
    if (method == 1) { f.o3coefvsha[pp] = 1.0; f.o3coefgsha[pp] = 1.0; f.o3coefvsun[pp] = 1.0; f.o3coefgsun[pp] = 1.0; }
    else { f.o3coefjmaxsha[pp] = 1.0; f.o3coefjmaxsun[pp] = 1.0; }
    return;
  }
  const int pp = filter_exposed[t] - begp0;
  const int ty = f.itype[pp];
  const int k = ty > 3 ? (f.pft_woody[ty] == 0 ? 2 : 1) : 0;
  const double usha = f.o3uptakesha[pp], usun = f.o3uptakesun[pp];
  if (method == 1) {
    f.o3coefvsha[pp] = usha == 0.0 ? 1.0 : fmax(0.0, fmin(1.0, photoInt[k] + photoSlope[k] * usha));
    f.o3coefgsha[pp] = usha == 0.0 ? 1.0 : fmax(0.0, fmin(1.0, condInt[k] + condSlope[k] * usha));
    f.o3coefvsun[pp] = usun == 0.0 ? 1.0 : fmax(0.0, fmin(1.0, photoInt[k] + photoSlope[k] * usun));
    f.o3coefgsun[pp] = usun == 0.0 ? 1.0 : fmax(0.0, fmin(1.0, condInt[k] + condSlope[k] * usun));
  } else {
    f.o3coefjmaxsha[pp] = usha == 0.0 ? 1.0 : fmax(0.0, fmin(1.0, jmaxInt[k] + jmaxSlope[k] * usha));
    f.o3coefjmaxsun[pp] = usun == 0.0 ? 1.0 : fmax(0.0, fmin(1.0, jmaxInt[k] + jmaxSlope[k] * usun));
  }
}
}  // namespace

extern "C" int ctsm_b200_calc_ozone_uptake(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_exposedvegp,
                                           const int32_t* filter_exposedvegp, const ctsm_ozone_fields_t* hf, int mem, ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_exposedvegp < 0 || (num_exposedvegp > 0 && !filter_exposedvegp)) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  OzoneDev d;
  const int32_t* dfp = filter_exposedvegp;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_OZONE
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_OZONE
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, mem == CTSM_MEM_HOST);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_exposedvegp, num_exposedvegp, &dfp);
    if (rc) return rc;
  }
  if (num_exposedvegp > 0) {
    ozone_uptake_kernel<<<grid_for(num_exposedvegp, 256), 256, 0, ctx->stream>>>(d, hf->alloc.begp, hf->alloc.begc, hf->alloc.begg,
                                                                                 (int)ctx->prm.dtime, num_exposedvegp, dfp);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}

extern "C" int ctsm_b200_calc_ozone_stress(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int num_exposedvegp,
                                           const int32_t* filter_exposedvegp, int num_noexposedvegp, const int32_t* filter_noexposedvegp,
                                           int stress_method, int is_time_to_run_luna, const ctsm_ozone_fields_t* hf, int mem,
                                           ctsm_status_t* st) {
  if (!ctx || !bounds || !hf || num_exposedvegp < 0 || num_noexposedvegp < 0 || (num_exposedvegp > 0 && !filter_exposedvegp) ||
      (num_noexposedvegp > 0 && !filter_noexposedvegp) || (stress_method != 1 && stress_method != 2))
    return CTSM_ERR_BAD_ARG;
  if (stress_method == 2 && !is_time_to_run_luna) { if (st) memset(st, 0, sizeof *st); return CTSM_OK; }   // OzoneMod.F90:697
  CUDA_TRY(cudaSetDevice(ctx->device));
  OzoneDev d;
  const int32_t *dfe = filter_exposedvegp, *dfn = filter_noexposedvegp;
  std::vector<StageField> fl;
#define CTSM_F(name, ctype, sub, lev, intent, us, usn, ref) \
  d.name = hf->name;                                        \
  fl.push_back(StageField{(void**)&d.name, (void*)hf->name, (int)sizeof(ctype), SUB_##sub, lev_shape(#lev).n, INTENT_##intent});
#define CTSM_FIELDS_OZONE
#include "../../include/ctsm_b200_fields.def"
#undef CTSM_FIELDS_OZONE
#undef CTSM_F
  for (auto& s : fl) if (!s.host_ptr) return CTSM_ERR_BAD_ARG;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_begin(ctx, fl, hf->alloc, *bounds, true);      // (the method not chosen leaves its coefficients untouched: keep them)
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter_exposedvegp, num_exposedvegp, &dfe);
    if (rc) return rc;
    rc = stage_filter(ctx, ctx->arena_filter1, filter_noexposedvegp, num_noexposedvegp, &dfn);
    if (rc) return rc;
  }
  const int n = num_exposedvegp + num_noexposedvegp;
  if (n > 0) {
    ozone_stress_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(d, hf->alloc.begp, stress_method, num_exposedvegp, dfe,
                                                                   num_noexposedvegp, dfn);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) {
    int rc = stage_end(ctx, fl, hf->alloc, *bounds);
    if (rc) return rc;
  }
  return finish_call(ctx, mem, st);
}
