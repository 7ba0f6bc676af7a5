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
  // P2P mode: the neighbours store the halo of x themselves (fc_p2p.cu) and raise hflag[c] to hseq
  const unsigned long long *hflag;
  unsigned long long hseq;
  int nconn;
};

template <int ROWS, int CAP, int MODE, bool STRIP>
__global__ void __launch_bounds__(ROWS)
k_spmv(int n, const int *__restrict__ ioffset, const int *__restrict__ ja, const double *__restrict__ a,
       const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ su,
       const double *__restrict__ w, const int *__restrict__ diag, double *__restrict__ adiag, strip_t st,
       double *partials, fc_scalars *sc, int step, fc_sync sy) {
  __shared__ double prod[CAP];
  __shared__ int s_off[ROWS + 1];
  __shared__ double s_red[64];
  if (MODE == MODE_DOT || MODE == MODE_DOT2) {
    if (!fc_kernel_begin(sc, sy)) return;
  }
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
        const int q0 = st.off[r], q1 = st.off[r + 1];
        if (q1 > q0 && st.hseq) {   // rows with processor faces wait for the neighbours' stores; the rest overlap them
          for (int c = 0; c < st.nconn; ++c)
            while (fc_ld_acquire_sys(st.hflag + c) < st.hseq) {}
        }
        for (int q = q0; q < q1; ++q) {
          const int i = st.idx[q];
          double t = st.apr[i] * __ldcg(x + st.halo0 + i);
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
    if (fc_grid_sum<2>(v, partials, &sc->ticket[0], s_red)) fc_reduction_done<2>(sc, sy, v, step);
  } else if (MODE != MODE_SPMV) {
    double v[1] = {acc};
    if (fc_grid_sum<1>(v, partials, &sc->ticket[0], s_red)) fc_reduction_done<1>(sc, sy, v, step);
  }
}

template <int MODE>
int launch(fc_context *ctx, const double *a, const double *x, double *y, const double *su, const double *w,
           double *adiag, int step, const fc_sync &sy) {
  const int n = ctx->n;
  strip_t st{ctx->strip_off, ctx->strip_idx, ctx->field[FC_APR], n, nullptr, 0ull, 0};
  if (ctx->p2p && ctx->halo_wait && (x == ctx->pk || x == ctx->zk)) {
    st.hflag = (const unsigned long long *)((const char *)ctx->arena + ctx->arena_hflag_off);
    st.hseq = ctx->halo_wait;
    st.nconn = (int)ctx->nbr_rank.size();
    ctx->halo_wait = 0;
  }
  const int nchunks = (n + 255) / 256;
  int grid = nchunks < FC_SMS * 8 ? nchunks : FC_SMS * 8;
  if (grid < 1) grid = 1;
  const bool strip = ctx->npro > 0;
  const bool small_rows = ctx->spmv_max_chunk <= 2304;
  if (!small_rows && grid > FC_SMS * 4) grid = FC_SMS * 4;
#define FC_SPMV_LAUNCH(CAP, STRIP)                                                                              \
  k_spmv<256, CAP, MODE, STRIP><<<grid, 256, 0, ctx->stream>>>(n, ctx->ioffset, ctx->ja, a, x, y, su, w, ctx->diag, \
                                                               adiag, st, ctx->partials, ctx->sc, step, sy)
  if (small_rows) { if (strip) FC_SPMV_LAUNCH(2304, true); else FC_SPMV_LAUNCH(2304, false); }
  else            { if (strip) FC_SPMV_LAUNCH(5632, true); else FC_SPMV_LAUNCH(5632, false); }
#undef FC_SPMV_LAUNCH
  FC_LAUNCH_CHECK();
  return FC_OK;
}

}  // namespace

int fc_launch_spmv(fc_context *ctx, const double *a, const double *x, double *y) {
  fc_sync sy{};
  return launch<MODE_SPMV>(ctx, a, x, y, nullptr, nullptr, nullptr, STEP_NONE, sy);
}

// y = A x fused with red[0] = w.y (and red[1] = y.y when `two`)
int fc_launch_spmv_dots(fc_context *ctx, const double *a, const double *x, double *y, const double *w, int two,
                        int step, const fc_sync &sy) {
  const bool sample = (size_t)(2 * ctx->spmv_sampled + 1) < ctx->spmv_ev.size();
  if (sample) FC_CUDA(cudaEventRecord(ctx->spmv_ev[2 * ctx->spmv_sampled], ctx->stream));
  FC_CHECK(two ? launch<MODE_DOT2>(ctx, a, x, y, nullptr, w, nullptr, step, sy)
               : launch<MODE_DOT>(ctx, a, x, y, nullptr, w, nullptr, step, sy));
  if (sample) {
    FC_CUDA(cudaEventRecord(ctx->spmv_ev[2 * ctx->spmv_sampled + 1], ctx->stream));
    ctx->spmv_sampled++;
  }
  return FC_OK;
}

int fc_launch_residual(fc_context *ctx, const double *a, const double *su, const double *x, double *res,
                       double *adiag, const fc_sync &sy) {
  return launch<MODE_RESID>(ctx, a, x, res, su, nullptr, adiag, STEP_RES0, sy);
}
