// filters.cu — setFiltersOneGroup on B200.
//
// Reference: src/main/filterMod.F90:303-592 (type clumpfilter :29-115).  The reference builds each of its 21 lists
// with its own sequential loop over the clump's columns / patches / landunits.  Here one kernel per subgrid level
// evaluates ALL membership predicates of an element once into a bit mask, and each list is then a stable
// (ascending-index) stream compaction of one bit: per-block counts -> exclusive scan -> scatter.  The lists are
// bit-exact by construction (a compaction has one right answer).  Rebuilt only when subgrid weights change, so this is
// plumbing, not a hot kernel: HBM-bound, 4 B per element per list.
#include "common.cuh"

namespace {
#define FB 256
#define FI 8
struct FilterIn {
  const int32_t *col_active, *col_landunit, *col_gridcell, *col_hyd;
  const int32_t *lun_active, *lun_lakpoi, *lun_urbpoi, *lun_itype;
  const int32_t *patch_active, *patch_landunit, *patch_itype;
  const int32_t* melt;
  int begc0, begl0, begp0, begg0;
  int include_inactive, use_cn, use_fates, use_fates_bgc, npcropmin, npcropmax;
};
// bit k of a column / patch / landunit mask = membership in the k-th list of its level (order of the CTSM_FLT_* enum)
__global__ void col_mask_kernel(FilterIn in, int begc, int n, uint32_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = begc + i, cc = c - in.begc0;
  uint32_t m = 0;
  if (in.col_active[cc] || in.include_inactive) {
    const int ll = in.col_landunit[cc] - in.begl0;
    const int lt = in.lun_itype[ll];
    const bool lak = in.lun_lakpoi[ll] != 0, urb = in.lun_urbpoi[ll] != 0;
    const bool soil = (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP);
    m |= 1u << 0;                                               // allc
    m |= lak ? (1u << 1) : (1u << 2);                           // lakec / nolakec
    if ((in.use_cn || in.use_fates_bgc) && soil) m |= 1u << 3;  // bgc_soilc
    if (soil) m |= 1u << 4;                                     // soilc
    if (in.col_hyd[cc]) m |= 1u << 5;                           // hydrologyc
    m |= urb ? (1u << 6) : (1u << 7);                           // urbanc / nourbanc
    if (lt == CTSM_ISTICE) m |= 1u << 8;                        // icec
    if (in.melt[in.col_gridcell[cc] - in.begg0] && (lt == CTSM_ISTICE || lt == CTSM_ISTSOIL)) m |= 1u << 9;   // do_smb_c
  }
  mask[i] = m;
}
__global__ void patch_mask_kernel(FilterIn in, int begp, int n, uint32_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = begp + i, pp = p - in.begp0;
  uint32_t m = 0;
  if (in.patch_active[pp] || in.include_inactive) {
    const int ll = in.patch_landunit[pp] - in.begl0;
    const int lt = in.lun_itype[ll];
    const bool lak = in.lun_lakpoi[ll] != 0, urb = in.lun_urbpoi[ll] != 0;
    const bool soil = (lt == CTSM_ISTSOIL || lt == CTSM_ISTCROP);
    if (lak) m |= 1u << 0;                                      // lakep
    else { m |= 1u << 1; if (!urb) m |= 1u << 2; }              // nolakep, nolakeurbanp
    if (in.use_cn && soil) m |= 1u << 3;                        // bgc_vegp
    if (soil) m |= 1u << 4;                                     // soilp
    if (!in.use_fates) {
      const int ivt = in.patch_itype[pp];
      if (ivt >= in.npcropmin && ivt <= in.npcropmax) m |= 1u << 5;   // pcropp
      else if (soil) m |= 1u << 6;                                    // soilnopcropp
    }
    m |= urb ? (1u << 7) : (1u << 8);                           // urbanp / nourbanp
  }
  mask[i] = m;
}
__global__ void lun_mask_kernel(FilterIn in, int begl, int n, uint32_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ll = begl + i - in.begl0;
  uint32_t m = 0;
  if (in.lun_active[ll] || in.include_inactive) m |= in.lun_urbpoi[ll] ? (1u << 0) : (1u << 1);
  mask[i] = m;
}

__global__ void __launch_bounds__(FB)
bit_count_kernel(int n, const uint32_t* __restrict__ mask, int bit, int* __restrict__ blockc) {
  __shared__ int sh[FB / 32];
  const int base = (blockIdx.x * FB + threadIdx.x) * FI;
  int c = 0;
  for (int k = 0; k < FI; ++k) if (base + k < n) c += (mask[base + k] >> bit) & 1u;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int k = 0; k < FB / 32; ++k) s += sh[k];
    blockc[blockIdx.x] = s;
  }
}
// single block: exclusive scan of the per-block counts; the total goes to *total
__global__ void block_scan_kernel(int nblocks, int* __restrict__ blockc, int* __restrict__ total) {
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
__global__ void __launch_bounds__(FB)
bit_scatter_kernel(int n, const uint32_t* __restrict__ mask, int bit, int first_index, const int* __restrict__ blockc,
                   int32_t* __restrict__ out) {
  __shared__ int sh[FB];
  const int base = (blockIdx.x * FB + threadIdx.x) * FI;
  int c = 0;
  bool y[FI];
  for (int k = 0; k < FI; ++k) {
    y[k] = (base + k < n) && ((mask[base + k] >> bit) & 1u);
    c += y[k] ? 1 : 0;
  }
  sh[threadIdx.x] = c;
  __syncthreads();
  for (int o = 1; o < FB; o <<= 1) {
    const int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  int pos = blockc[blockIdx.x] + sh[threadIdx.x] - c;
  for (int k = 0; k < FI; ++k)
    if (y[k]) out[pos++] = first_index + base + k;
}
}  // namespace

extern "C" int ctsm_b200_set_filters(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, const ctsm_filter_inputs_t* in,
                                     ctsm_filters_t* out, int mem) {
  if (!ctx || !bounds || !in || !out) return CTSM_ERR_BAD_ARG;
  const void* need[] = {in->col_active, in->col_landunit, in->col_gridcell, in->col_hydrologically_active, in->lun_active,
                        in->lun_lakpoi, in->lun_urbpoi, in->lun_itype, in->patch_active, in->patch_landunit, in->patch_itype,
                        in->melt_replaced_by_ice_grc};
  for (const void* q : need) if (!q) return CTSM_ERR_BAD_ARG;
  for (int k = 0; k < CTSM_FLT_COUNT; ++k) { if (!out->list[k]) return CTSM_ERR_BAD_ARG; out->num[k] = 0; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int nc = bounds->endc - bounds->begc + 1, np = bounds->endp - bounds->begp + 1, nl = bounds->endl - bounds->begl + 1;
  const int ncol_a = in->alloc.endc - in->alloc.begc + 1, nlun_a = in->alloc.endl - in->alloc.begl + 1;
  const int npat_a = in->alloc.endp - in->alloc.begp + 1, ngrc_a = in->alloc.endg - in->alloc.begg + 1;
  const int nmax = nc > np ? (nc > nl ? nc : nl) : (np > nl ? np : nl);
  if (nmax <= 0) return CTSM_OK;
  const int nblk_max = grid_for(nmax, FB * FI);
  // device scratch: mask[nmax], blockc[nblk_max], totals[CTSM_FLT_COUNT]; HOST mode adds input mirrors and one output list
  size_t ints = (size_t)nmax + (size_t)nblk_max + CTSM_FLT_COUNT + 64;
  if (mem != CTSM_MEM_DEVICE) ints += 4 * (size_t)ncol_a + 4 * (size_t)nlun_a + 3 * (size_t)npat_a + (size_t)ngrc_a + (size_t)nmax;
  int rc = arena_reserve(ctx, ctx->arena_ints, sizeof(int32_t) * ints);
  if (rc) return rc;
  int32_t* ip = (int32_t*)ctx->arena_ints.p;
  uint32_t* mask = (uint32_t*)ip; ip += nmax;
  int* blockc = ip; ip += nblk_max;
  int* totals = ip; ip += CTSM_FLT_COUNT + 64;
  FilterIn fi;
  fi.begc0 = in->alloc.begc; fi.begl0 = in->alloc.begl; fi.begp0 = in->alloc.begp; fi.begg0 = in->alloc.begg;
  fi.include_inactive = in->include_inactive; fi.use_cn = in->use_cn; fi.use_fates = in->use_fates;
  fi.use_fates_bgc = in->use_fates_bgc; fi.npcropmin = in->npcropmin; fi.npcropmax = in->npcropmax;
  auto stage = [&](const int32_t* h, int n) -> const int32_t* {
    if (mem == CTSM_MEM_DEVICE) return h;
    int32_t* d = ip; ip += n;
    cudaMemcpyAsync(d, h, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s);
    return d;
  };
  fi.col_active = stage(in->col_active, ncol_a); fi.col_landunit = stage(in->col_landunit, ncol_a);
  fi.col_gridcell = stage(in->col_gridcell, ncol_a); fi.col_hyd = stage(in->col_hydrologically_active, ncol_a);
  fi.lun_active = stage(in->lun_active, nlun_a); fi.lun_lakpoi = stage(in->lun_lakpoi, nlun_a);
  fi.lun_urbpoi = stage(in->lun_urbpoi, nlun_a); fi.lun_itype = stage(in->lun_itype, nlun_a);
  fi.patch_active = stage(in->patch_active, npat_a); fi.patch_landunit = stage(in->patch_landunit, npat_a);
  fi.patch_itype = stage(in->patch_itype, npat_a); fi.melt = stage(in->melt_replaced_by_ice_grc, ngrc_a);
  int32_t* dlist = (mem == CTSM_MEM_DEVICE) ? nullptr : ip;     // one list at a time in HOST mode
  CUDA_TRY(cudaMemsetAsync(totals, 0, sizeof(int) * CTSM_FLT_COUNT, s));

  struct Level { int first, count, n, beg; } levels[3] = {{CTSM_FLT_ALLC, 10, nc, bounds->begc},
                                                           {CTSM_FLT_LAKEP, 9, np, bounds->begp},
                                                           {CTSM_FLT_URBANL, 2, nl, bounds->begl}};
  int htot[CTSM_FLT_COUNT];
  for (int lv = 0; lv < 3; ++lv) {
    const Level& L = levels[lv];
    if (L.n <= 0) continue;
    if (lv == 0) col_mask_kernel<<<grid_for(L.n, 256), 256, 0, s>>>(fi, L.beg, L.n, mask);
    else if (lv == 1) patch_mask_kernel<<<grid_for(L.n, 256), 256, 0, s>>>(fi, L.beg, L.n, mask);
    else lun_mask_kernel<<<grid_for(L.n, 256), 256, 0, s>>>(fi, L.beg, L.n, mask);
    ctx->launches++;
    const int nblk = grid_for(L.n, FB * FI);
    for (int b = 0; b < L.count; ++b) {
      const int k = L.first + b;
      bit_count_kernel<<<nblk, FB, 0, s>>>(L.n, mask, b, blockc);
      block_scan_kernel<<<1, 1024, 0, s>>>(nblk, blockc, totals + k);
      int32_t* dst = (mem == CTSM_MEM_DEVICE) ? out->list[k] : dlist;
      bit_scatter_kernel<<<nblk, FB, 0, s>>>(L.n, mask, b, L.beg, blockc, dst);
      ctx->launches += 3;
      if (mem != CTSM_MEM_DEVICE) {
        CUDA_TRY(cudaMemcpyAsync(&htot[k], totals + k, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (htot[k] > 0)
          CUDA_TRY(cudaMemcpyAsync(out->list[k], dlist, sizeof(int32_t) * (size_t)htot[k], cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
      }
    }
  }
  if (mem == CTSM_MEM_DEVICE) {
    CUDA_TRY(cudaMemcpyAsync(htot, totals, sizeof(int) * CTSM_FLT_COUNT, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
  }
  for (int lv = 0; lv < 3; ++lv)
    for (int b = 0; b < levels[lv].count; ++b) {
      const int k = levels[lv].first + b;
      out->num[k] = levels[lv].n > 0 ? htot[k] : 0;
    }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { fprintf(stderr, "ctsm_b200_set_filters: %s\n", cudaGetErrorString(e)); return CTSM_ERR_NO_DEVICE; }
  return CTSM_OK;
}
