// common.cuh — shared declarations of the ctsm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <unordered_map>
#include <utility>
#include "../../include/ctsm_b200.h"

#define NLEVSNO CTSM_NLEVSNO
#define NLEVGRND CTSM_NLEVGRND
#define NLEVSOI CTSM_NLEVSOI
#define SNOSOI_LO (-NLEVSNO + 1)
#define SNOSOI0_LO (-NLEVSNO)

// Physical constants: components/cdeps/share/shr_const_mod.F90:16-53 and
// src/main/clm_varcon.F90:50-122.  Spelled exactly as the reference spells them
// so the double-precision values are identical.
namespace cst {
constexpr double tfrz = 273.15;
constexpr double denh2o = 1.000e3;
constexpr double denice = 0.917e3;
constexpr double cpliq = 4.188e3;
constexpr double cpice = 2.11727e3;
constexpr double hfus = 3.337e5;
constexpr double hvap = 2.501e6;
constexpr double sb = 5.67e-8;
constexpr double grav = 9.80616;
constexpr double tkair = 0.023;
constexpr double tkice = 2.290;
constexpr double tkwat = 0.57;
constexpr double capr = 0.34;
constexpr double cnfac = 0.5;
constexpr double thk_bedrock = 3.0;
constexpr double csol_bedrock = 2.0e6;
constexpr double thin_sfclayer = 1.0e-6;   // SoilTemperatureMod.F90:84
constexpr double m_to_mm = 1.e3;           // SoilWaterMovementMod.F90:65
}  // namespace cst

// Un-suffixed Fortran literals are REAL(4) constants promoted to double.
#define R4(x) ((double)(float)(x))

// Device-side "first failure" record (error convention).
// key = subgrid index (32 bits, filters are ascending so the lowest index is the
// first point the reference's loop would have reached) | code (8) | info (24).
struct DevStatus {
  unsigned long long key;
  int n_warnings;
  int pad;
};

__device__ __forceinline__ void report_failure(DevStatus* ds, int index, int code, int info) {
  const unsigned long long key = ((unsigned long long)(uint32_t)index << 32) |
                                 ((unsigned long long)(code & 0xff) << 24) | (unsigned long long)(info & 0xffffff);
  atomicMin(&ds->key, key);
}

struct ctsm_b200_ctx {
  ctsm_params_t prm;
  int device;
  cudaStream_t stream;
  DevStatus* d_status;        // device
  DevStatus* h_status;        // pinned host copy
  int64_t launches;
  // staging mirrors for CTSM_MEM_HOST calls: one grow-only device arena per purpose
  struct Arena { void* p = nullptr; size_t cap = 0; };
  Arena arena_fields, arena_filter0, arena_filter1, arena_scratch, arena_ints;
  int32_t* d_patchmask = nullptr; size_t patchmask_cap = 0;
  // CanopyFluxes: second stream for the tail kernel, per-round events, pinned survivor counts (canopy.cu)
  cudaStream_t stream2 = nullptr;
  std::vector<cudaEvent_t> ev_round, ev_tail;
  cudaEvent_t ev_join = nullptr;
  int* h_counts = nullptr; int h_counts_cap = 0;
  struct Tuning { int tail_max = 0, nt_budget = 0, tail_lanes = 1, nt_split = 1, soil_stream = 1, sw_warp = 1, sink_warp = 1; } tune;
  int* dbg_counts = nullptr; int* dbg_tail_end = nullptr; int dbg_npass = 0;
  // CTSM_MEM_HOST resident window (abi.cu: ctsm_b200_host_window_begin / _end): persistent device mirrors keyed by the
  // host array's base address, what is fresh on the device in the current window, copy streams, event pool
  struct Mirror { void* d = nullptr; size_t bytes = 0; std::vector<std::pair<int, int>> fresh; };
  std::unordered_map<const void*, Mirror> mirrors;
  bool window_open = false;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
  std::vector<Arena> win_chunks; size_t win_chunk = 0, win_off = 0;      // filters of the window's calls
  uint64_t win_h2d_bytes = 0, win_d2h_bytes = 0;
  struct BalPending { void* pinned; ctsm_balance_report_t* rep; int DAnstep; };
  std::vector<BalPending> bal_pending;
  std::vector<void*> bal_pinned, bal_dev; size_t bal_pinned_used = 0;
  void* bal_last_dev = nullptr;
  // perturbed-parameter ensembles: per-member values of the scalar parameters (device arrays, owned), ctsm_b200_set_member_params
  struct MemberPrm {
    int n = 0;
    double *e_ice = nullptr, *csoilc = nullptr, *cv = nullptr, *a_coef = nullptr, *z_dl = nullptr;
    int32_t* col_member = nullptr; int begc = 0, endc = -1;
  } member;
};

// records "name at file:line: text" for ctsm_b200_last_cuda_error() and maps the error to a CTSM_ERR_* code
int cuda_fail(cudaError_t e, const char* file, int line);

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) return cuda_fail(_e, __FILE__, __LINE__);                        \
  } while (0)

// --- staging of field tables (abi.cu) -----------------------------------------
enum SubLevel { SUB_GRC = 0, SUB_LUN = 1, SUB_COL = 2, SUB_PATCH = 3, SUB_PFT = 4 };
enum Intent { INTENT_IN = 1, INTENT_OUT = 2, INTENT_INOUT = 3 };
struct LevShape { int lo; int n; };
__host__ inline LevShape lev_shape(const char* lev) {
  if (!strcmp(lev, "L1")) return {1, 1};
  if (!strcmp(lev, "SNOSOI")) return {-NLEVSNO + 1, NLEVSNO + NLEVGRND};
  if (!strcmp(lev, "SNOSOI0")) return {-NLEVSNO, NLEVSNO + NLEVGRND + 1};
  if (!strcmp(lev, "GRND")) return {1, NLEVGRND};
  if (!strcmp(lev, "SOI")) return {1, NLEVSOI};
  if (!strcmp(lev, "SNO")) return {-NLEVSNO + 1, NLEVSNO};
  if (!strcmp(lev, "SNO1")) return {-NLEVSNO + 1, NLEVSNO + 1};
  if (!strcmp(lev, "VEGWCS")) return {1, CTSM_NVEGWCS};
  if (!strcmp(lev, "CAN")) return {1, CTSM_NLEVCAN};
  if (!strcmp(lev, "LAK")) return {1, CTSM_NLEVLAK};
  if (!strcmp(lev, "NUMRAD")) return {1, 2};
  if (!strcmp(lev, "AER")) return {1, 14};
  if (!strcmp(lev, "PHS2")) return {1, 2 * CTSM_NLEVCAN};
  return {1, 1};
}

struct StageField {
  void** dev_slot;      // where the device pointer goes (member of the device-side field struct)
  void* host_ptr;       // host array (element (alloc_beg, lev_lo))
  int elem_size;
  int sub;
  int nlev;
  int intent;
};

int arena_reserve(ctsm_b200_ctx* ctx, ctsm_b200_ctx::Arena& a, size_t bytes);
int ensure_round_events(ctsm_b200_ctx* ctx, int n);
int window_event(ctsm_b200_ctx* ctx, cudaEvent_t* ev);
// balance.cu: thresholds / abort decision of the BalanceCheck calls issued inside a resident window
int balance_finish_pending(ctsm_b200_ctx* ctx, ctsm_status_t* st);
int stage_begin(ctsm_b200_ctx* ctx, std::vector<StageField>& fl, const ctsm_bounds_t& alloc, const ctsm_bounds_t& call,
                bool preserve_out);
int stage_end(ctsm_b200_ctx* ctx, std::vector<StageField>& fl, const ctsm_bounds_t& alloc, const ctsm_bounds_t& call);
int stage_filter(ctsm_b200_ctx* ctx, ctsm_b200_ctx::Arena& a, const int32_t* host_filter, int numf, const int32_t** dev_filter);
int finish_call(ctsm_b200_ctx* ctx, int mem, ctsm_status_t* st);
void decode_status(const DevStatus& ds, ctsm_status_t* st);

static inline int sub_beg(const ctsm_bounds_t& b, int sub) {
  if (sub == SUB_PFT) return 0;
  return sub == SUB_GRC ? b.begg : sub == SUB_LUN ? b.begl : sub == SUB_COL ? b.begc : b.begp;
}
static inline int sub_end(const ctsm_bounds_t& b, int sub, int npft_table) {
  if (sub == SUB_PFT) return npft_table - 1;
  return sub == SUB_GRC ? b.endg : sub == SUB_LUN ? b.endl : sub == SUB_COL ? b.endc : b.endp;
}

// launch geometry: grids sized in whole waves of the 148 SMs where the work allows
static inline int grid_for(int n, int block) { return (n + block - 1) / block; }
