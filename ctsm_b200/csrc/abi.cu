// abi.cu — lifecycle, status plumbing and host<->device staging of the C ABI
// declared in include/ctsm_b200.h.  No compute here.
#include "common.cuh"
#include "../../include/ctsm_b200_defaults.h"

#include <stdlib.h>

static const char* kVersion = "ctsm_b200 0.2.0 (sm_100a, fp64, -fmad=false)";

static char g_last_cuda_error[256] = "";
int cuda_fail(cudaError_t e, const char* file, int line) {
  snprintf(g_last_cuda_error, sizeof g_last_cuda_error, "%s at %s:%d: %s", cudaGetErrorName(e), file, line, cudaGetErrorString(e));
  fprintf(stderr, "ctsm_b200: CUDA error %s\n", g_last_cuda_error);
  (void)cudaGetLastError();      // clear the sticky launch-error slot where that is possible
  if (e == cudaErrorMemoryAllocation) return CTSM_ERR_NOMEM;
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return CTSM_ERR_NO_DEVICE;
  return CTSM_ERR_CUDA;
}
extern "C" const char* ctsm_b200_last_cuda_error(void) { return g_last_cuda_error; }
static void free_member_params(ctsm_b200_ctx* ctx);

extern "C" const char* ctsm_b200_version(void) { return kVersion; }

extern "C" void ctsm_b200_default_params(ctsm_params_t* p) { ctsm_default_params_fill(p); }

extern "C" int ctsm_b200_init(const ctsm_params_t* p, ctsm_b200_ctx** out) {
  if (!p || !out) return CTSM_ERR_BAD_ARG;
  *out = nullptr;
  if (p->abi_version != CTSM_B200_ABI_VERSION) return CTSM_ERR_BAD_ARG;
  if (p->nlevsno != CTSM_NLEVSNO || p->nlevgrnd != CTSM_NLEVGRND || p->nlevsoi != CTSM_NLEVSOI) {
    fprintf(stderr, "ctsm_b200_init: kernels are compiled for nlevsno=%d nlevgrnd=%d nlevsoi=%d\n", CTSM_NLEVSNO,
            CTSM_NLEVGRND, CTSM_NLEVSOI);
    return CTSM_ERR_BAD_ARG;
  }
  if ((p->use_hydrstress != 0 && p->use_hydrstress != 1) || (p->z0param_method != 1 && p->z0param_method != 2) ||
      (p->stomatalcond_mtd != 1 && p->stomatalcond_mtd != 2) || p->itmax_canopy_fluxes < 1)
    return CTSM_ERR_BAD_ARG;
  if (p->calc_human_stress_indices != 0 && p->calc_human_stress_indices != 1) return CTSM_ERR_BAD_ARG;
  if (p->npft_table < CTSM_MXPFT + 1 || p->npft_table % (CTSM_MXPFT + 1) != 0) return CTSM_ERR_BAD_ARG;
  if (p->upper_boundary_condition != 1 || (p->lower_boundary_condition != 1 && p->lower_boundary_condition != 2))
    return CTSM_ERR_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    fprintf(stderr, "ctsm_b200_init: no CUDA device (there is no CPU fallback)\n");
    return CTSM_ERR_NO_DEVICE;
  }
  if (p->device < 0 || p->device >= ndev) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(p->device));
  ctsm_b200_ctx* ctx = new ctsm_b200_ctx();
  ctx->prm = *p;
  ctx->device = p->device;
  ctx->launches = 0;
  ctx->stream = nullptr; ctx->d_status = nullptr; ctx->h_status = nullptr;
  if (const char* e = getenv("CTSM_B200_TAIL_MAX")) ctx->tune.tail_max = atoi(e);
  if (const char* e = getenv("CTSM_B200_NT_BUDGET")) ctx->tune.nt_budget = atoi(e);
  if (const char* e = getenv("CTSM_B200_TAIL_LANES")) ctx->tune.tail_lanes = atoi(e);
  if (const char* e = getenv("CTSM_B200_NT_SPLIT")) ctx->tune.nt_split = atoi(e);
  if (const char* e = getenv("CTSM_B200_SOIL_STREAM")) ctx->tune.soil_stream = atoi(e);
  if (const char* e = getenv("CTSM_B200_SW_WARP")) ctx->tune.sw_warp = atoi(e);
  if (const char* e = getenv("CTSM_B200_SINK_WARP")) ctx->tune.sink_warp = atoi(e);
  const int rc = [&]() -> int {
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc(&ctx->d_status, sizeof(DevStatus)));
    CUDA_TRY(cudaMallocHost(&ctx->h_status, sizeof(DevStatus)));
    DevStatus init; init.key = ~0ULL; init.n_warnings = 0; init.pad = 0;
    CUDA_TRY(cudaMemcpyAsync(ctx->d_status, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return CTSM_OK;
  }();
  if (rc != CTSM_OK) {           // nothing of a half-built context survives
    ctsm_b200_finalize(ctx);
    return rc;
  }
  *out = ctx;
  return CTSM_OK;
}

extern "C" int ctsm_b200_finalize(ctsm_b200_ctx* ctx) {
  if (!ctx) return CTSM_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
  cudaFree(ctx->arena_fields.p); cudaFree(ctx->arena_filter0.p); cudaFree(ctx->arena_filter1.p);
  cudaFree(ctx->arena_scratch.p); cudaFree(ctx->arena_ints.p); cudaFree(ctx->d_patchmask);
  cudaFree(ctx->d_status);
  if (ctx->h_status) cudaFreeHost(ctx->h_status);
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  if (ctx->s_h2d) { cudaStreamSynchronize(ctx->s_h2d); cudaStreamDestroy(ctx->s_h2d); }
  if (ctx->s_d2h) { cudaStreamSynchronize(ctx->s_d2h); cudaStreamDestroy(ctx->s_d2h); }
  for (auto& kv : ctx->mirrors) cudaFree(kv.second.d);
  for (auto& c : ctx->win_chunks) cudaFree(c.p);
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  for (auto q : ctx->bal_pinned) cudaFreeHost(q);
  for (auto q : ctx->bal_dev) cudaFree(q);
  free_member_params(ctx);
  for (auto e : ctx->ev_round) cudaEventDestroy(e);
  for (auto e : ctx->ev_tail) cudaEventDestroy(e);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return CTSM_OK;
}

extern "C" int ctsm_b200_set_soil_tuning(ctsm_b200_ctx* ctx, int soil_stream) {
  if (!ctx) return CTSM_ERR_BAD_ARG;
  if (soil_stream >= 0) ctx->tune.soil_stream = soil_stream;
  return CTSM_OK;
}

extern "C" int ctsm_b200_set_sink_tuning(ctsm_b200_ctx* ctx, int sink_warp) {
  if (!ctx) return CTSM_ERR_BAD_ARG;
  if (sink_warp >= 0) ctx->tune.sink_warp = sink_warp;
  return CTSM_OK;
}

extern "C" int ctsm_b200_set_soilwater_tuning(ctsm_b200_ctx* ctx, int sw_warp) {
  if (!ctx) return CTSM_ERR_BAD_ARG;
  if (sw_warp >= 0) ctx->tune.sw_warp = sw_warp;
  return CTSM_OK;
}

extern "C" int ctsm_b200_set_tuning(ctsm_b200_ctx* ctx, int tail_max, int nt_budget, int tail_lanes, int nt_split) {
  if (!ctx) return CTSM_ERR_BAD_ARG;
  if (nt_split >= 0) ctx->tune.nt_split = nt_split;
  if (tail_max >= 0) ctx->tune.tail_max = tail_max;
  if (nt_budget >= 0) ctx->tune.nt_budget = nt_budget;
  if (tail_lanes >= 0) ctx->tune.tail_lanes = tail_lanes;
  return CTSM_OK;
}

static void free_member_params(ctsm_b200_ctx* ctx) {
  ctsm_b200_ctx::MemberPrm& m = ctx->member;
  cudaFree(m.e_ice); cudaFree(m.csoilc); cudaFree(m.cv); cudaFree(m.a_coef); cudaFree(m.z_dl); cudaFree(m.col_member);
  m = ctsm_b200_ctx::MemberPrm();
}

extern "C" int ctsm_b200_set_member_params(ctsm_b200_ctx* ctx, int nmember, const double* e_ice, const double* csoilc,
                                           const double* cv, const double* a_coef, const double* z_dl,
                                           const int32_t* col_member, int begc, int endc) {
  if (!ctx || nmember < 0) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  free_member_params(ctx);
  if (nmember == 0) return CTSM_OK;
  if (nmember * (CTSM_MXPFT + 1) != ctx->prm.npft_table) return CTSM_ERR_BAD_ARG;      // same member count as the PFT tables
  if (e_ice && (!col_member || endc < begc)) return CTSM_ERR_BAD_ARG;
  ctsm_b200_ctx::MemberPrm& m = ctx->member;
  m.n = nmember;
  const double* src[5] = {e_ice, csoilc, cv, a_coef, z_dl};
  double** dst[5] = {&m.e_ice, &m.csoilc, &m.cv, &m.a_coef, &m.z_dl};
  for (int k = 0; k < 5; ++k) {
    if (!src[k]) continue;
    CUDA_TRY(cudaMalloc(dst[k], sizeof(double) * (size_t)nmember));
    CUDA_TRY(cudaMemcpy(*dst[k], src[k], sizeof(double) * (size_t)nmember, cudaMemcpyHostToDevice));
  }
  if (col_member && endc >= begc) {
    const size_t n = (size_t)(endc - begc + 1);
    for (size_t i = 0; i < n; ++i)
      if (col_member[i] < 0 || col_member[i] >= nmember) { free_member_params(ctx); return CTSM_ERR_BAD_ARG; }
    CUDA_TRY(cudaMalloc(&m.col_member, sizeof(int32_t) * n));
    CUDA_TRY(cudaMemcpy(m.col_member, col_member, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
    m.begc = begc; m.endc = endc;
  }
  return CTSM_OK;
}

// events and the pinned count buffer of CanopyFluxes' rounds (grown on first use, itmax is a run-time parameter)
int ensure_round_events(ctsm_b200_ctx* ctx, int n) {
  while ((int)ctx->ev_round.size() < n) {
    cudaEvent_t a = nullptr, b = nullptr;
    CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    ctx->ev_round.push_back(a);
    CUDA_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    ctx->ev_tail.push_back(b);
  }
  if (ctx->h_counts_cap < n) {
    if (ctx->h_counts) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); CUDA_TRY(cudaFreeHost(ctx->h_counts)); ctx->h_counts = nullptr; }
    CUDA_TRY(cudaMallocHost(&ctx->h_counts, sizeof(int) * (size_t)n));
    ctx->h_counts_cap = n;
  }
  return CTSM_OK;
}

extern "C" void* ctsm_b200_stream(ctsm_b200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" void* ctsm_b200_stream_of(ctsm_b200_ctx* ctx, int kind) {
  if (!ctx) return nullptr;
  switch (kind) {
    case 0: return (void*)ctx->stream;
    case 1: return (void*)ctx->stream2;
    case 2: return (void*)ctx->s_h2d;
    case 3: return (void*)ctx->s_d2h;
    default: return nullptr;
  }
}
extern "C" int64_t ctsm_b200_launch_count(const ctsm_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int ctsm_b200_host_register(void* ptr, uint64_t bytes) {
  if (!ptr || !bytes) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return CTSM_OK;
}
extern "C" int ctsm_b200_host_unregister(void* ptr) {
  if (!ptr) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaHostUnregister(ptr));
  return CTSM_OK;
}

static const char* message_for(int code) {
  switch (code) {
    case CTSM_ERR_DGBSV: return "BandDiagonal ERROR: dgbsv returned error code";
    case CTSM_ERR_DGTSV: return "soilwater_moisture_form:: problem with the lapack solver";
    case CTSM_ERR_FORC_HGT: return "CanopyFluxes: forcing height is below canopy height";
    case CTSM_ERR_GS_NEG: return "PhotosynthesisHydraulicStress: negative stomatal conductance";
    case CTSM_ERR_BRENT: return "brent_PHS: root must be bracketed";
    case CTSM_ERR_QUADRATIC: return "quadratic solution error: b^2 - 4ac is negative";
    case CTSM_ERR_SNOW_NEGATIVE: return "In UpdateState_TopLayerFluxes, h2osoi_ice has gone significantly negative";
    case CTSM_ERR_RH: return "ERROR RH is negative / greater than a hundred (Wet_BulbS)";
    case CTSM_ERR_URBAN: return "urban column in filter is outside the ctsm_b200 hot path";
    case CTSM_ERR_BALANCE: return "BalanceCheck: balance error exceeds threshold / c2g: sumwt is greater than 1.0";
    default: return "";
  }
}

void decode_status(const DevStatus& ds, ctsm_status_t* st) {
  if (!st) return;
  memset(st, 0, sizeof *st);
  st->n_warnings = ds.n_warnings;
  if (ds.key == ~0ULL) return;
  st->subgrid_index = (int32_t)(ds.key >> 32);
  st->code = (int32_t)((ds.key >> 24) & 0xff);
  int info = (int)(ds.key & 0xffffff);
  if (info & 0x800000) info |= ~0xffffff;   // sign-extend
  st->info = info;
  if (st->code == CTSM_ERR_BALANCE) st->subgrid_level = CTSM_SUBGRID_GRIDCELL;     /* device-side: the c2g weight check */
  else st->subgrid_level = (st->code == CTSM_ERR_FORC_HGT || st->code == CTSM_ERR_GS_NEG || st->code == CTSM_ERR_BRENT ||
                       st->code == CTSM_ERR_QUADRATIC || st->code == CTSM_ERR_RH) ? CTSM_SUBGRID_PATCH : CTSM_SUBGRID_COLUMN;
  snprintf(st->msg, sizeof st->msg, "%s", message_for(st->code));
  if (st->code == CTSM_ERR_SNOW_NEGATIVE && info == 3)
    snprintf(st->msg, sizeof st->msg, "ERROR: capping procedure failed (negative mass remaining)");
  if (st->code == CTSM_ERR_SNOW_NEGATIVE && info == 2)
    snprintf(st->msg, sizeof st->msg, "In RenewCondensation, h2osoi_ice has gone significantly negative");
  if (st->code == CTSM_ERR_SNOW_NEGATIVE && info == 1)
    snprintf(st->msg, sizeof st->msg, "In UpdateState_TopLayerFluxes, h2osoi_liq has gone significantly negative");
}

extern "C" int ctsm_b200_sync(ctsm_b200_ctx* ctx, ctsm_status_t* st) {
  if (!ctx) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  DevStatus got = *ctx->h_status;
  if (got.key != ~0ULL || got.n_warnings != 0) {   // re-arm the record
    DevStatus init; init.key = ~0ULL; init.n_warnings = 0; init.pad = 0;
    *ctx->h_status = init;
    CUDA_TRY(cudaMemcpyAsync(ctx->d_status, ctx->h_status, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  ctsm_status_t local;
  decode_status(got, &local);
  if (st) *st = local;
  if (!ctx->window_open && !ctx->bal_pending.empty()) {      // BalanceCheck calls issued since the last synchronisation
    ctsm_status_t bst;
    memset(&bst, 0, sizeof bst);
    const int rb = balance_finish_pending(ctx, &bst);
    if (local.code == CTSM_OK && rb != CTSM_OK) {
      bst.n_warnings = local.n_warnings;
      if (st) *st = bst;
      return rb;
    }
  }
  return local.code;
}

// DEVICE calls are asynchronous (status is collected by ctsm_b200_sync);
// HOST calls are synchronous and return the status of this call.
int finish_call(ctsm_b200_ctx* ctx, int mem, ctsm_status_t* st) {
  // a kernel that failed to launch (bad configuration, missing image) must never pass silently
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (st) memset(st, 0, sizeof *st);
    return cuda_fail(e, __FILE__, __LINE__);
  }
  if (mem == CTSM_MEM_DEVICE || ctx->window_open) {     // asynchronous: ctsm_b200_sync / ctsm_b200_host_window_end collect the status
    if (st) memset(st, 0, sizeof *st);
    return CTSM_OK;
  }
  return ctsm_b200_sync(ctx, st);
}

// Grow-only arenas.  Growing replaces the allocation, so work that may still read the old one is drained first
// (steady state never grows: the sizes of a model run do not change after the first step).
int arena_reserve(ctsm_b200_ctx* ctx, ctsm_b200_ctx::Arena& a, size_t bytes) {
  if (bytes <= a.cap) return CTSM_OK;
  if (a.p) {
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->stream2) CUDA_TRY(cudaStreamSynchronize(ctx->stream2));
    CUDA_TRY(cudaFree(a.p));
  }
  a.p = nullptr; a.cap = 0;
  size_t want = bytes + bytes / 8 + 4096;
  CUDA_TRY(cudaMalloc(&a.p, want));
  a.cap = want;
  return CTSM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Resident window for CTSM_MEM_HOST callers.  Outside a window every HOST call is self-contained: upload all fields,
// run, download, synchronise - 36 GB over PCIe for one f02 step, most of it repeated (fields shared by routines) or
// only there so that values outside the filters survive the download.  Inside a window (one model step, or any run
// of hot-path calls between which the host does not touch the arrays)
//   * every host array has a persistent device mirror (created and filled once, at its first use ever);
//   * a call uploads, on the H2D stream, only the index ranges of its IN / INOUT fields that are not yet fresh on
//     the device in this window (uploaded or written by an earlier call), and never an OUT field;
//   * calls return as soon as their work is queued; kernels wait for their uploads through an event, the D2H stream
//     downloads a call's OUT / INOUT fields over its bounds as soon as its kernels are done - so uploads of the next
//     clump, kernels of this one and downloads of the previous one overlap when the host calls clump after clump;
//   * ctsm_b200_host_window_end waits for everything and returns the first failure, like ctsm_b200_sync.
static void fresh_missing(const std::vector<std::pair<int, int>>& fresh, int b, int e, std::vector<std::pair<int, int>>& out) {
  int cur = b;
  for (const auto& iv : fresh) {
    if (iv.second < cur) continue;
    if (iv.first > e) break;
    if (iv.first > cur) out.push_back({cur, iv.first - 1});
    if (iv.second + 1 > cur) cur = iv.second + 1;
    if (cur > e) break;
  }
  if (cur <= e) out.push_back({cur, e});
}
static void fresh_add(std::vector<std::pair<int, int>>& fresh, int b, int e) {
  std::vector<std::pair<int, int>> out;
  bool placed = false;
  for (const auto& iv : fresh) {
    if (iv.second + 1 < b) out.push_back(iv);
    else if (iv.first > e + 1) { if (!placed) { out.push_back({b, e}); placed = true; } out.push_back(iv); }
    else { if (iv.first < b) b = iv.first; if (iv.second > e) e = iv.second; }
  }
  if (!placed) out.push_back({b, e});
  fresh.swap(out);
}

int window_event(ctsm_b200_ctx* ctx, cudaEvent_t* ev) {
  if (ctx->ev_used == ctx->ev_pool.size()) {
    cudaEvent_t e = nullptr;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->ev_pool.push_back(e);
  }
  *ev = ctx->ev_pool[ctx->ev_used++];
  return CTSM_OK;
}

extern "C" int ctsm_b200_host_window_begin(ctsm_b200_ctx* ctx) {
  if (!ctx || ctx->window_open) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (!ctx->s_h2d) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
  if (!ctx->s_d2h) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));          // whatever ran before the window is finished
  for (auto& kv : ctx->mirrors) kv.second.fresh.clear();
  ctx->ev_used = 0; ctx->win_chunk = 0; ctx->win_off = 0; ctx->win_h2d_bytes = 0; ctx->win_d2h_bytes = 0;
  ctx->bal_pending.clear(); ctx->bal_pinned_used = 0;
  ctx->window_open = true;
  return CTSM_OK;
}

extern "C" int ctsm_b200_host_window_end(ctsm_b200_ctx* ctx, ctsm_status_t* st) {
  if (!ctx || !ctx->window_open) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaStreamSynchronize(ctx->s_h2d));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->s_d2h));
  ctx->window_open = false;
  return ctsm_b200_sync(ctx, st);        // device-side first failure, then the deferred BalanceCheck decisions
}

extern "C" int ctsm_b200_host_window_bytes(const ctsm_b200_ctx* ctx, uint64_t* h2d, uint64_t* d2h) {
  if (!ctx || !h2d || !d2h) return CTSM_ERR_BAD_ARG;
  *h2d = ctx->win_h2d_bytes; *d2h = ctx->win_d2h_bytes;
  return CTSM_OK;
}

extern "C" int ctsm_b200_host_invalidate(ctsm_b200_ctx* ctx, const void* host_ptr) {
  if (!ctx || ctx->window_open) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (!host_ptr) {
    for (auto& kv : ctx->mirrors) CUDA_TRY(cudaFree(kv.second.d));
    ctx->mirrors.clear();
    return CTSM_OK;
  }
  auto it = ctx->mirrors.find(host_ptr);
  if (it != ctx->mirrors.end()) { CUDA_TRY(cudaFree(it->second.d)); ctx->mirrors.erase(it); }
  return CTSM_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int copy_range(void* dst, const void* src, int es, size_t ld, int nlev, size_t off, size_t ncall,
                      cudaMemcpyKind kind, cudaStream_t s) {
  if (ncall == 0) return CTSM_OK;
  char* d = (char*)dst + off * es;
  const char* h = (const char*)src + off * es;
  if (ncall == ld) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)es * ld * nlev, kind, s));
  } else {
    CUDA_TRY(cudaMemcpy2DAsync(d, ld * es, h, ld * es, ncall * es, nlev, kind, s));
  }
  return CTSM_OK;
}

// Allocates the device mirrors (same layout and leading dimension as the host
// arrays), points the device-side struct members at them and uploads IN/INOUT
// (and, when preserve_out, OUT) fields over the call's bounds.
static int stage_begin_window(ctsm_b200_ctx* ctx, std::vector<StageField>& fl, const ctsm_bounds_t& alloc, const ctsm_bounds_t& call) {
  cudaStream_t up = ctx->s_h2d;
  std::vector<std::pair<int, int>> need;
  for (auto& f : fl) {
    if (!f.host_ptr) return CTSM_ERR_BAD_ARG;
    const int a0 = sub_beg(alloc, f.sub), a1 = sub_end(alloc, f.sub, ctx->prm.npft_table);
    const size_t ld = (size_t)(a1 - a0 + 1);
    const size_t bytes = (size_t)f.elem_size * ld * f.nlev;
    ctsm_b200_ctx::Mirror& m = ctx->mirrors[f.host_ptr];
    if (!m.d || m.bytes != bytes) {
      // first use of this host array: create the mirror and fill it with the array as it stands (once; this is what
      // keeps elements no routine writes - outside the filters - equal to the host's when the field is downloaded)
      if (m.d) CUDA_TRY(cudaFree(m.d));
      m.d = nullptr; m.fresh.clear();
      CUDA_TRY(cudaMalloc(&m.d, bytes > 0 ? bytes : 8));
      m.bytes = bytes;
      CUDA_TRY(cudaMemcpyAsync(m.d, f.host_ptr, bytes, cudaMemcpyHostToDevice, up));
      ctx->win_h2d_bytes += bytes;
      m.fresh.push_back({a0, a1});
    }
    *f.dev_slot = m.d;
    const int c0 = sub_beg(call, f.sub), c1 = sub_end(call, f.sub, ctx->prm.npft_table);
    if (c1 < c0) continue;
    if (f.intent & INTENT_IN) {
      need.clear();
      fresh_missing(m.fresh, c0, c1, need);
      for (const auto& iv : need) {
        const int rc = copy_range(m.d, f.host_ptr, f.elem_size, ld, f.nlev, (size_t)(iv.first - a0), (size_t)(iv.second - iv.first + 1),
                                  cudaMemcpyHostToDevice, up);
        if (rc) return rc;
        ctx->win_h2d_bytes += (uint64_t)f.elem_size * (uint64_t)(iv.second - iv.first + 1) * (uint64_t)f.nlev;
      }
    }
    fresh_add(m.fresh, c0, c1);
  }
  cudaEvent_t ev;
  int rc = window_event(ctx, &ev);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(ev, up));
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ev, 0));
  return CTSM_OK;
}

int stage_begin(ctsm_b200_ctx* ctx, std::vector<StageField>& fl, const ctsm_bounds_t& alloc, const ctsm_bounds_t& call,
                bool preserve_out) {
  if (ctx->window_open) return stage_begin_window(ctx, fl, alloc, call);
  size_t total = 0;
  for (auto& f : fl) {
    if (!f.host_ptr) return CTSM_ERR_BAD_ARG;
    const size_t ld = (size_t)(sub_end(alloc, f.sub, ctx->prm.npft_table) - sub_beg(alloc, f.sub) + 1);
    total += align256((size_t)f.elem_size * ld * f.nlev);
  }
  int rc = arena_reserve(ctx, ctx->arena_fields, total);
  if (rc) return rc;
  size_t off = 0;
  for (auto& f : fl) {
    const size_t ld = (size_t)(sub_end(alloc, f.sub, ctx->prm.npft_table) - sub_beg(alloc, f.sub) + 1);
    void* d = (char*)ctx->arena_fields.p + off;
    *f.dev_slot = d;
    off += align256((size_t)f.elem_size * ld * f.nlev);
    if ((f.intent & INTENT_IN) || preserve_out) {
      const size_t o = (size_t)(sub_beg(call, f.sub) - sub_beg(alloc, f.sub));
      const size_t n = (size_t)(sub_end(call, f.sub, ctx->prm.npft_table) - sub_beg(call, f.sub) + 1);
      rc = copy_range(d, f.host_ptr, f.elem_size, ld, f.nlev, o, n, cudaMemcpyHostToDevice, ctx->stream);
      if (rc) return rc;
    }
  }
  return CTSM_OK;
}

int stage_end(ctsm_b200_ctx* ctx, std::vector<StageField>& fl, const ctsm_bounds_t& alloc, const ctsm_bounds_t& call) {
  cudaStream_t down = ctx->stream;
  if (ctx->window_open) {
    cudaEvent_t ev;
    int rc = window_event(ctx, &ev);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ev, ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->s_d2h, ev, 0));
    down = ctx->s_d2h;
  }
  for (auto& f : fl) {
    if (!(f.intent & INTENT_OUT)) continue;
    const size_t ld = (size_t)(sub_end(alloc, f.sub, ctx->prm.npft_table) - sub_beg(alloc, f.sub) + 1);
    const size_t o = (size_t)(sub_beg(call, f.sub) - sub_beg(alloc, f.sub));
    const size_t n = (size_t)(sub_end(call, f.sub, ctx->prm.npft_table) - sub_beg(call, f.sub) + 1);
    int rc = copy_range(f.host_ptr, *f.dev_slot, f.elem_size, ld, f.nlev, o, n, cudaMemcpyDeviceToHost, down);
    if (rc) return rc;
    if (ctx->window_open) ctx->win_d2h_bytes += (uint64_t)f.elem_size * (uint64_t)n * (uint64_t)f.nlev;
  }
  return CTSM_OK;
}

// filters of the calls of a window live until the window ends (calls are asynchronous): bump allocation from chunks
static int stage_filter_window(ctsm_b200_ctx* ctx, const int32_t* host_filter, int numf, const int32_t** dev_filter) {
  const size_t bytes = align256(sizeof(int32_t) * (size_t)(numf > 0 ? numf : 1));
  while (ctx->win_chunk < ctx->win_chunks.size() && ctx->win_off + bytes > ctx->win_chunks[ctx->win_chunk].cap) { ctx->win_chunk++; ctx->win_off = 0; }
  if (ctx->win_chunk == ctx->win_chunks.size()) {
    ctsm_b200_ctx::Arena c;
    c.cap = bytes > ((size_t)64 << 20) ? bytes : ((size_t)64 << 20);
    CUDA_TRY(cudaMalloc(&c.p, c.cap));
    ctx->win_chunks.push_back(c);
    ctx->win_off = 0;
  }
  char* d = (char*)ctx->win_chunks[ctx->win_chunk].p + ctx->win_off;
  ctx->win_off += bytes;
  if (numf > 0) {
    CUDA_TRY(cudaMemcpyAsync(d, host_filter, sizeof(int32_t) * (size_t)numf, cudaMemcpyHostToDevice, ctx->s_h2d));
    ctx->win_h2d_bytes += sizeof(int32_t) * (uint64_t)numf;
    cudaEvent_t ev;
    int rc = window_event(ctx, &ev);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ev, ctx->s_h2d));
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ev, 0));
  }
  *dev_filter = (const int32_t*)d;
  return CTSM_OK;
}

int stage_filter(ctsm_b200_ctx* ctx, ctsm_b200_ctx::Arena& a, const int32_t* host_filter, int numf,
                 const int32_t** dev_filter) {
  if (ctx->window_open) return stage_filter_window(ctx, host_filter, numf, dev_filter);
  int rc = arena_reserve(ctx, a, sizeof(int32_t) * (size_t)(numf > 0 ? numf : 1));
  if (rc) return rc;
  if (numf > 0)
    CUDA_TRY(cudaMemcpyAsync(a.p, host_filter, sizeof(int32_t) * (size_t)numf, cudaMemcpyHostToDevice, ctx->stream));
  *dev_filter = (const int32_t*)a.p;
  return CTSM_OK;
}
