// solvers.cu — batched solver kernels (one column per thread, column index
// fastest => coalesced loads) and their C-ABI entry points.
#include "solvers.cuh"

#define MAXLEV 64   // upper bound on ubj-lbj+1 for the stand-alone solver entry points

// ---------------------------------------------------------------------------
// Tridiagonal: TridiagonalMod.F90:23-91
// HBM-bound: 4 reads + 1 write of 8 B per active level.
// gam(j) and u(j) stay in per-thread local arrays between the forward and the
// backward sweep so that u is written exactly once.
__global__ void __launch_bounds__(128)
tridiagonal_kernel(int begc, int ld, int lbj, int ubj, const int32_t* __restrict__ jtop, int numf,
                   const int32_t* __restrict__ filter, const double* __restrict__ a, const double* __restrict__ b,
                   const double* __restrict__ c, const double* __restrict__ r, double* __restrict__ u) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int ci = filter[fc] - begc;
  const int jt = jtop[ci];
  double gam[MAXLEV], ul[MAXLEV];
  const int j0 = max(jt, lbj);
  double bet = b[(size_t)(j0 - lbj) * ld + ci];
  double uprev = 0.0, cprev = 0.0;
  for (int j = j0; j <= ubj; ++j) {
    const size_t o = (size_t)(j - lbj) * ld + ci;
    const double rj = r[o];
    if (j == jt) {
      uprev = rj / bet;
    } else {
      const double aj = a[o];
      const double g = cprev / bet;
      gam[j - lbj] = g;
      bet = b[o] - aj * g;
      uprev = (rj - aj * uprev) / bet;
    }
    ul[j - lbj] = uprev;
    cprev = c[o];
  }
  double unext = ul[ubj - lbj];
  u[(size_t)(ubj - lbj) * ld + ci] = unext;
  for (int j = ubj - 1; j >= j0; --j) {
    unext = ul[j - lbj] - gam[j + 1 - lbj] * unext;
    u[(size_t)(j - lbj) * ld + ci] = unext;
  }
}

// ---------------------------------------------------------------------------
// BandDiagonal: BandDiagonalMod.F90:29-221 with dgbsv(kl=ku=2) semantics.
__global__ void __launch_bounds__(128)
banddiagonal_kernel(int begc, int ld, int lbj, const int32_t* __restrict__ jtop, const int32_t* __restrict__ jbot,
                    int numf, const int32_t* __restrict__ filter, const double* __restrict__ b,
                    const double* __restrict__ r, double* __restrict__ u, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int ci = filter[fc] - begc;
  const int jt = jtop[ci], jb = jbot[ci];
  const int n = jb - jt + 1;
  double U[MAXLEV][5], y[MAXLEV];
  auto row = [&](int i, double* e) -> double {
    const size_t lev = (size_t)(jt + i - lbj);
    const double* bp = b + lev * 5 * ld + ci;
    e[0] = bp[4 * (size_t)ld];   // band 5: A(i, i-2)
    e[1] = bp[3 * (size_t)ld];   // band 4: A(i, i-1)
    e[2] = bp[2 * (size_t)ld];   // band 3: diagonal
    e[3] = bp[1 * (size_t)ld];   // band 2: A(i, i+1)
    e[4] = bp[0];                // band 1: A(i, i+2)
    return r[lev * ld + ci];
  };
  const int info = band5_solve<MAXLEV>(n, row, U, y);
  if (info != 0) {
    // dgbsv leaves B = rhs when the factor is singular; BandDiagonal copies it to u and aborts (:198-213)
    for (int i = 0; i < n; ++i) u[(size_t)(jt + i - lbj) * ld + ci] = r[(size_t)(jt + i - lbj) * ld + ci];
    report_failure(ds, ci + begc, CTSM_ERR_DGBSV, info);
    return;
  }
  for (int i = 0; i < n; ++i) u[(size_t)(jt + i - lbj) * ld + ci] = y[i];
}

// ---------------------------------------------------------------------------
// dgtsv call site: SoilWaterMovementMod.F90:1279-1299
__global__ void __launch_bounds__(128)
dgtsv_batch_kernel(int begc, int ld, const int32_t* __restrict__ nlayers, int numf, const int32_t* __restrict__ filter,
                   const double* __restrict__ amx, const double* __restrict__ bmx, const double* __restrict__ cmx,
                   const double* __restrict__ rmx, double* __restrict__ x, DevStatus* ds) {
  const int fc = blockIdx.x * blockDim.x + threadIdx.x;
  if (fc >= numf) return;
  const int ci = filter[fc] - begc;
  const int n = nlayers[ci];
  double dl[MAXLEV], d[MAXLEV], du[MAXLEV], rhs[MAXLEV];
  for (int j = 0; j < n; ++j) {
    const size_t o = (size_t)j * ld + ci;
    d[j] = bmx[o];
    rhs[j] = rmx[o];
    if (j < n - 1) { dl[j] = amx[o + ld]; du[j] = cmx[o]; }
  }
  const int info = dgtsv_solve(n, dl, d, du, rhs);
  if (info != 0) { report_failure(ds, ci + begc, CTSM_ERR_DGTSV, info); return; }
  for (int j = 0; j < n; ++j) x[(size_t)j * ld + ci] = rhs[j];
}

// ---------------------------------------------------------------------------
// host side

struct HostStage {   // tiny helper for the array-argument entry points
  ctsm_b200_ctx* ctx;
  std::vector<std::pair<void*, std::pair<void*, size_t>>> outs;   // host dst, (dev src, bytes)
  size_t off = 0;
  int rc = 0;
  void* in(const void* host, size_t bytes, bool upload) {
    void* d = (char*)ctx->arena_fields.p + off;
    off += (bytes + 255) & ~(size_t)255;
    if (upload && cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = CTSM_ERR_NO_DEVICE;
    return d;
  }
};

extern "C" int ctsm_b200_tridiagonal(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int lbj, int ubj,
                                     const int32_t* jtop, int numf, const int32_t* filter, const double* a,
                                     const double* b, const double* c, const double* r, double* u, int mem) {
  if (!ctx || !bounds || !jtop || !filter || !a || !b || !c || !r || !u) return CTSM_ERR_BAD_ARG;
  const int nl = ubj - lbj + 1;
  if (nl < 1 || nl > MAXLEV || numf < 0) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  const int ld = bounds->endc - bounds->begc + 1;
  const size_t nb = sizeof(double) * (size_t)ld * nl;
  const int32_t *djtop = jtop, *dfilter = filter;
  const double *da = a, *db = b, *dc = c, *dr = r;
  double* du = u;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = arena_reserve(ctx, ctx->arena_fields, 5 * (nb + 256) + sizeof(int32_t) * (size_t)ld + 256);
    if (rc) return rc;
    HostStage hs{ctx};
    da = (double*)hs.in(a, nb, true); db = (double*)hs.in(b, nb, true); dc = (double*)hs.in(c, nb, true);
    dr = (double*)hs.in(r, nb, true); du = (double*)hs.in(u, nb, true);   // u is inout: untouched entries persist
    djtop = (int32_t*)hs.in(jtop, sizeof(int32_t) * (size_t)ld, true);
    if (hs.rc) return hs.rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter, numf, &dfilter);
    if (rc) return rc;
  }
  if (numf > 0) {
    tridiagonal_kernel<<<grid_for(numf, 128), 128, 0, ctx->stream>>>(bounds->begc, ld, lbj, ubj, djtop, numf, dfilter,
                                                                      da, db, dc, dr, du);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) CUDA_TRY(cudaMemcpyAsync(u, du, nb, cudaMemcpyDeviceToHost, ctx->stream));
  return finish_call(ctx, mem, nullptr);
}

extern "C" int ctsm_b200_banddiagonal(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int lbj, int ubj,
                                      const int32_t* jtop, const int32_t* jbot, int numf, const int32_t* filter,
                                      int nband, const double* b, const double* r, double* u, int mem,
                                      ctsm_status_t* st) {
  if (!ctx || !bounds || !jtop || !jbot || !filter || !b || !r || !u) return CTSM_ERR_BAD_ARG;
  const int nl = ubj - lbj + 1;
  if (nband != 5 || nl < 1 || nl > MAXLEV || numf < 0) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  const int ld = bounds->endc - bounds->begc + 1;
  const size_t nb = sizeof(double) * (size_t)ld * nl;
  const int32_t *djtop = jtop, *djbot = jbot, *dfilter = filter;
  const double *db = b, *dr = r;
  double* du = u;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = arena_reserve(ctx, ctx->arena_fields, 7 * (nb + 256) + 2 * (sizeof(int32_t) * (size_t)ld + 256));
    if (rc) return rc;
    HostStage hs{ctx};
    db = (double*)hs.in(b, 5 * nb, true); dr = (double*)hs.in(r, nb, true); du = (double*)hs.in(u, nb, true);
    djtop = (int32_t*)hs.in(jtop, sizeof(int32_t) * (size_t)ld, true);
    djbot = (int32_t*)hs.in(jbot, sizeof(int32_t) * (size_t)ld, true);
    if (hs.rc) return hs.rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter, numf, &dfilter);
    if (rc) return rc;
  }
  if (numf > 0) {
    banddiagonal_kernel<<<grid_for(numf, 128), 128, 0, ctx->stream>>>(bounds->begc, ld, lbj, djtop, djbot, numf,
                                                                       dfilter, db, dr, du, ctx->d_status);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) CUDA_TRY(cudaMemcpyAsync(u, du, nb, cudaMemcpyDeviceToHost, ctx->stream));
  return finish_call(ctx, mem, st);
}

extern "C" int ctsm_b200_dgtsv_batch(ctsm_b200_ctx* ctx, const ctsm_bounds_t* bounds, int nlev,
                                     const int32_t* nlayers, int numf, const int32_t* filter, const double* amx,
                                     const double* bmx, const double* cmx, const double* rmx, double* x, int mem,
                                     ctsm_status_t* st) {
  if (!ctx || !bounds || !nlayers || !filter || !amx || !bmx || !cmx || !rmx || !x) return CTSM_ERR_BAD_ARG;
  if (nlev < 1 || nlev > MAXLEV || numf < 0) return CTSM_ERR_BAD_ARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  const int ld = bounds->endc - bounds->begc + 1;
  const size_t nb = sizeof(double) * (size_t)ld * nlev;
  const int32_t *dn = nlayers, *dfilter = filter;
  const double *da = amx, *db = bmx, *dc = cmx, *dr = rmx;
  double* dx = x;
  if (mem != CTSM_MEM_DEVICE) {
    int rc = arena_reserve(ctx, ctx->arena_fields, 5 * (nb + 256) + sizeof(int32_t) * (size_t)ld + 256);
    if (rc) return rc;
    HostStage hs{ctx};
    da = (double*)hs.in(amx, nb, true); db = (double*)hs.in(bmx, nb, true); dc = (double*)hs.in(cmx, nb, true);
    dr = (double*)hs.in(rmx, nb, true); dx = (double*)hs.in(x, nb, true);
    dn = (int32_t*)hs.in(nlayers, sizeof(int32_t) * (size_t)ld, true);
    if (hs.rc) return hs.rc;
    rc = stage_filter(ctx, ctx->arena_filter0, filter, numf, &dfilter);
    if (rc) return rc;
  }
  if (numf > 0) {
    dgtsv_batch_kernel<<<grid_for(numf, 128), 128, 0, ctx->stream>>>(bounds->begc, ld, dn, numf, dfilter, da, db, dc,
                                                                      dr, dx, ctx->d_status);
    ctx->launches++;
  }
  if (mem != CTSM_MEM_DEVICE) CUDA_TRY(cudaMemcpyAsync(x, dx, nb, cudaMemcpyDeviceToHost, ctx->stream));
  return finish_call(ctx, mem, st);
}
