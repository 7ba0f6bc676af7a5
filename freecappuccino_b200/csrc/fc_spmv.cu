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
#include "fc_spmv_pipe.cuh"

namespace {

enum { MODE_SPMV = FC_MODE_SPMV, MODE_DOT = FC_MODE_DOT, MODE_RESID = FC_MODE_RESID, MODE_DOT2 = FC_MODE_DOT2 };

using strip_t = fc_strip;

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
        if (st.any32[r >> 5]) {
          const int q0 = st.off[r], q1 = st.off[r + 1];
          if (q1 > q0 && st.hseq) {   // rows with processor faces wait for the neighbours' stores; the rest overlap them
            fc_spin_guard g;
            for (int c = 0; c < st.nconn; ++c)
              while (fc_ld_acquire_sys(st.hflag + c) < st.hseq) g.tick();
          }
          for (int q = q0; q < q1; ++q) {
            const int i = st.idx[q];
            double t = st.apr[i] * __ldcg(x + st.halo0 + i);
            v = (MODE == MODE_RESID) ? v - t : v + t;
          }
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

// The same product as a TMA-fed pipeline (fc_spmv_pipe.cuh): every CTA owns an equal-cost contiguous share of
// the rows, so the grid has no tail wave.
template <int T, int CAP, int S, int MODE, bool STRIP>
__global__ void __launch_bounds__(T)
k_spmv_tma(fc_spmv_mat M, fc_spmv_vec V, strip_t st, double *partials, fc_scalars *sc, int step, fc_sync sy) {
  extern __shared__ __align__(128) unsigned char fc_smem_raw[];
  __shared__ double s_red[64];
  if (MODE == MODE_DOT || MODE == MODE_DOT2) {
    if (!fc_kernel_begin(sc, sy)) return;
  }
  fc_spmv_pipe<T, CAP, S> pipe;
  int rbeg, rend;
  fc_row_range(M.n, STRIP ? st.off : nullptr, &rbeg, &rend);
  pipe.init(reinterpret_cast<fc_spmv_smem<T, CAP, S> *>(fc_smem_raw), M.ioffset, rbeg, rend, STRIP ? st.off : nullptr);
  pipe.prefetch(M);
  pipe.halo_pending = STRIP && pipe.cta_strip && st.hseq != 0ull;
  double acc = 0.0, acc2 = 0.0;
  pipe.template sweep<MODE, STRIP>(M, V, st, acc, acc2);
  if (MODE == MODE_DOT2) {
    double v[2] = {acc, acc2};
    if (fc_grid_sum<2>(v, partials, &sc->ticket[0], s_red)) fc_reduction_done<2>(sc, sy, v, step);
  } else if (MODE != MODE_SPMV) {
    double v[1] = {acc};
    if (fc_grid_sum<1>(v, partials, &sc->ticket[0], s_red)) fc_reduction_done<1>(sc, sy, v, step);
  }
}

template <int T, int CAP, int S, int MODE, bool STRIP>
int launch_tma(fc_context *ctx, const fc_spmv_mat &M, const fc_spmv_vec &V, const strip_t &st, int step,
               const fc_sync &sy, bool *ok) {
  auto kern = k_spmv_tma<T, CAP, S, MODE, STRIP>;
  const size_t smem = sizeof(fc_spmv_smem<T, CAP, S>);
  // per instantiation AND per device: the opt-in above 48 KB of dynamic shared memory is a per-device function
  // attribute, so a second context on another GPU of the same process has to set it again
  static int per_sm_dev[FC_MAX_DEVICES];
  static bool per_sm_set[FC_MAX_DEVICES];
  const int dev = ctx->device >= 0 && ctx->device < FC_MAX_DEVICES ? ctx->device : 0;
  if (!per_sm_set[dev] || dev != ctx->device) {
    int v = 0;
    FC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, T, smem));
    per_sm_dev[dev] = v < 1 ? 0 : v;
    per_sm_set[dev] = true;
  }
  const int per_sm = per_sm_dev[dev];
  if (per_sm == 0) { *ok = false; return FC_OK; }
  const long long groups = ((long long)M.n + 31) / 32;
  long long grid = (long long)ctx->sms * per_sm;
  if (grid > groups) grid = groups;
  if (grid < 1) grid = 1;
  // the chunk boundaries of a CTA must fit its shared-memory table (an unweighted share is the largest)
  const long long rows_per_cta = (((long long)M.n + ctx->npro + grid - 1) / grid / 32 + 2) * 32;
  if ((rows_per_cta + T - 1) / T > FC_KB_MAX) { *ok = false; return FC_OK; }
  kern<<<(int)grid, T, smem, ctx->stream>>>(M, V, st, ctx->partials, ctx->sc, step, sy);
  FC_LAUNCH_CHECK();
  *ok = true;
  return FC_OK;
}

template <int MODE>
int launch(fc_context *ctx, const double *a, const double *x, double *y, const double *su, const double *w,
           double *adiag, int step, const fc_sync &sy) {
  const int n = ctx->n;
  strip_t st{ctx->strip_off, ctx->strip_idx, ctx->field[FC_APR], ctx->strip_any32, n, nullptr, 0ull, 0};
  if (ctx->p2p && ctx->halo_wait && (x == ctx->pk || x == ctx->zk)) {
    st.hflag = (const unsigned long long *)((const char *)ctx->arena + ctx->arena_hflag_off);
    st.hseq = ctx->halo_wait;
    st.nconn = (int)ctx->nbr_rank.size();
    ctx->halo_wait = 0;
  }
  const bool strip = ctx->npro > 0;
  // stand-alone launches: the pipeline wins while a launch is short (its CTAs own equal row ranges: no tail
  // wave), the stream kernel on long launches (profiles/r01_variants.txt); inside the persistent DPCG kernel
  // the pipeline is always used
  if (ctx->tune_spmv == 1 || (ctx->tune_spmv == 2 && n < 4000000)) {
    fc_spmv_mat M{n, ctx->ioffset, ctx->ja, a};
    fc_spmv_vec V{x, y, su, w, ctx->diag, adiag};
    bool ok = false;
#define FC_TMA(T, CAP, S)                                                                         \
  do {                                                                                            \
    if (strip) FC_CHECK((launch_tma<T, CAP, S, MODE, true>(ctx, M, V, st, step, sy, &ok)));       \
    else       FC_CHECK((launch_tma<T, CAP, S, MODE, false>(ctx, M, V, st, step, sy, &ok)));      \
  } while (0)
    if (ctx->spmv_max_chunk <= 2000) {   // <= 8.75 non-zeros per row: hexahedra
      switch (ctx->tune_pipe) {
        case 1: FC_TMA(256, 2304, 2); break;
        case 2: FC_TMA(256, 2048, 2); break;
        case 3: FC_TMA(128, 1024, 2); break;
        default: FC_TMA(256, 2304, 3); break;
      }
    } else {                             // polyhedra (~15 per row); longer chunks fall back to global loads
      FC_TMA(256, 4096, 2);
    }
#undef FC_TMA
    if (ok) return FC_OK;
  }
  const int nchunks = (n + 255) / 256;
  int grid = nchunks < ctx->sms * 8 ? nchunks : ctx->sms * 8;
  if (grid < 1) grid = 1;
  const bool small_rows = ctx->spmv_max_chunk <= 2304;
  if (!small_rows && grid > ctx->sms * 4) grid = ctx->sms * 4;
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
