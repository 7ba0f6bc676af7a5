// CSR sparse matrix-vector product of the Krylov loop (dpcg.f90:105-110, iccg.f90:139-147,
// bicgstab.f90:142-147) and the initial residual (dpcg.f90:51-56).
//
// Layout: the reference's CSR (ioffset/ja/a), FP64 values + int32 columns, 12 B per
// non-zero.  Finite-volume rows are short (7 entries on a hex mesh, ~15 on polyhedra),
// so a warp per row would idle most lanes.  Instead every CTA owns ROWS consecutive rows:
//   phase 1  all threads stream the CTA's contiguous slice of a/ja with fully coalesced
//            loads, multiply by the gathered x and park the products in shared memory;
//   phase 2  one thread per row adds its products left to right -- the reference's own
//            summation order, so y is bit-identical to the Fortran loop.
// Algorithmic traffic: 12 B per non-zero + 20 B per row (ioffset 4, x 8, y 8); x is
// gathered through L1/L2, every x element is used by its ~7 neighbouring rows.
#include "fc_reduce.cuh"

namespace {

enum { MODE_SPMV = 0, MODE_DOT = 1, MODE_RESID = 2, MODE_DOT2 = 3 };  // DOT: w.y ; DOT2: w.y and y.y

struct strip_t {               // processor-boundary coupling kept outside the CSR (src-parallel `apr`)
  const int *off;              // [n+1] per-row range into idx, or nullptr on a single rank
  const int *idx;              // processor-face index i (0-based, ascending per row)
  const double *apr;           // [npro]
  int halo0;                   // x[halo0 + i] = value on the other rank
};

template <int ROWS, int CAP, int MODE, bool STRIP>
__global__ void __launch_bounds__(ROWS)
k_spmv(int n, const int *__restrict__ ioffset, const int *__restrict__ ja, const double *__restrict__ a,
       const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ su,
       const double *__restrict__ w, const int *__restrict__ diag, double *__restrict__ adiag, strip_t st, double *partials, fc_scalars *sc,
       int step, int local_step) {
  __shared__ double prod[CAP];
  __shared__ int s_off[ROWS + 1];
  __shared__ double s_red[64];
  if ((MODE == MODE_DOT || MODE == MODE_DOT2) && sc->done) return;
  const int tid = threadIdx.x;
  const int nchunks = (n + ROWS - 1) / ROWS;
  double acc = 0.0, acc2 = 0.0;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int r0 = chunk * ROWS;
    const int nr = min(ROWS, n - r0);
    if (tid < nr) s_off[tid] = ioffset[r0 + tid];
    if (tid == 0) s_off[nr] = ioffset[r0 + nr];
    __syncthreads();
    const int k0 = s_off[0], k1 = s_off[nr];
    const bool staged = (k1 - k0) <= CAP;
    if (staged) {
#pragma unroll 4
      for (int k = k0 + tid; k < k1; k += ROWS) prod[k - k0] = a[k] * x[ja[k]];
    }
    __syncthreads();
    if (tid < nr) {
      const int r = r0 + tid;
      const int s = s_off[tid], e = s_off[tid + 1];
      double v = (MODE == MODE_RESID) ? su[r] : 0.0;
      if (staged) {
        for (int k = s; k < e; ++k) v = (MODE == MODE_RESID) ? v - prod[k - k0] : v + prod[k - k0];
      } else {  // very long rows: straight from global memory
        for (int k = s; k < e; ++k) {
          double t = a[k] * x[ja[k]];
          v = (MODE == MODE_RESID) ? v - t : v + t;
        }
      }
      if (STRIP) {
        for (int q = st.off[r]; q < st.off[r + 1]; ++q) {
          const int i = st.idx[q];
          double t = st.apr[i] * x[st.halo0 + i];
          v = (MODE == MODE_RESID) ? v - t : v + t;
        }
      }
      y[r] = v;
      if (MODE == MODE_DOT || MODE == MODE_DOT2) acc += w[r] * v;
      if (MODE == MODE_DOT2) acc2 += v * v;
      if (MODE == MODE_RESID) {
        acc += fabs(v);
        adiag[r] = a[diag[r]];
      }
    }
    __syncthreads();
  }
  if (MODE == MODE_DOT2) {
    double v[2] = {acc, acc2};
    if (fc_grid_sum<2>(v, partials, &sc->ticket[0], s_red)) {
      sc->red[0] = v[0];
      sc->red[1] = v[1];
      if (local_step) fc_scalar_step(sc, step, nullptr);
    }
  } else if (MODE != MODE_SPMV) {
    double v[1] = {acc};
    if (fc_grid_sum<1>(v, partials, &sc->ticket[0], s_red)) {
      sc->red[0] = v[0];
      if (local_step) fc_scalar_step(sc, step, nullptr);
    }
  }
}

template <int MODE>
int launch(fc_context *ctx, const double *a, const double *x, double *y, const double *su, const double *w,
           double *adiag, int step) {
  const int n = ctx->n;
  strip_t st{ctx->strip_off, ctx->strip_idx, ctx->field[FC_APR], n};
  const int nchunks = (n + 255) / 256;
  int grid = nchunks < FC_SMS * 8 ? nchunks : FC_SMS * 8;
  if (grid < 1) grid = 1;
  const int local = ctx->nranks == 1;
  const bool strip = ctx->npro > 0;
  const bool small_rows = ctx->spmv_max_chunk <= 2304;
  if (!small_rows && grid > FC_SMS * 4) grid = FC_SMS * 4;
#define FC_SPMV_LAUNCH(CAP, STRIP)                                                                              \
  k_spmv<256, CAP, MODE, STRIP><<<grid, 256, 0, ctx->stream>>>(n, ctx->ioffset, ctx->ja, a, x, y, su, w, ctx->diag, \
                                                               adiag, st, ctx->partials, ctx->sc, step, local)
  if (small_rows) { if (strip) FC_SPMV_LAUNCH(2304, true); else FC_SPMV_LAUNCH(2304, false); }
  else            { if (strip) FC_SPMV_LAUNCH(5632, true); else FC_SPMV_LAUNCH(5632, false); }
#undef FC_SPMV_LAUNCH
  FC_LAUNCH_CHECK();
  return FC_OK;
}

}  // namespace

int fc_launch_spmv(fc_context *ctx, const double *a, const double *x, double *y) {
  return launch<MODE_SPMV>(ctx, a, x, y, nullptr, nullptr, nullptr, STEP_NONE);
}

// y = A x fused with red[0] = w.y (and red[1] = y.y when `two`)
int fc_launch_spmv_dots(fc_context *ctx, const double *a, const double *x, double *y, const double *w, int two,
                        int step) {
  const bool sample = (size_t)(2 * ctx->spmv_sampled + 1) < ctx->spmv_ev.size();
  if (sample) FC_CUDA(cudaEventRecord(ctx->spmv_ev[2 * ctx->spmv_sampled], ctx->stream));
  FC_CHECK(two ? launch<MODE_DOT2>(ctx, a, x, y, nullptr, w, nullptr, step)
               : launch<MODE_DOT>(ctx, a, x, y, nullptr, w, nullptr, step));
  if (sample) {
    FC_CUDA(cudaEventRecord(ctx->spmv_ev[2 * ctx->spmv_sampled + 1], ctx->stream));
    ctx->spmv_sampled++;
  }
  return FC_OK;
}

int fc_launch_residual(fc_context *ctx, const double *a, const double *su, const double *x, double *res,
                       double *adiag) {
  return launch<MODE_RESID>(ctx, a, x, res, su, nullptr, adiag, STEP_RES0);
}
