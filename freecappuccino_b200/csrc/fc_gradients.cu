// `grad(phi,dPhidxi)` as the reference dispatches it (src/gradients.f90:95-151): Gauss (fc_assemble.cu), the
// three least-squares variants and the optional slope limiter.  SURVEY.md 8(f) rank 3; bodies in fc_grad_body.cuh.
//
// Algorithmic bytes per call (n cells, F inner faces, B boundary faces; one thread per cell over the
// cell-to-face map: 3 ints per entry, 2F + B entries):
//   lstsq / lstsq_dm solve   map 12 (2F+B) + gathered centres and phi 32 per entry (L2-resident re-use) + dmat 72 n
//                            + gradient 24 n
//   lstsq_qr solve           map other 4 (2F+B) + phi 8 per entry + D 144 n + gradient 24 n  -- the cheapest: no geometry
//   limiter                  CSR row 4 nnz + neighbour phi and centres 32 nnz + gradient 48 n
// All HBM / L2-gather bound; FP64, no tensor cores.
#include "fc_grad_body.cuh"
#include "fc_body_views.cuh"
#include "fc_reduce.cuh"

int fc_grad_gauss_dev(fc_context *ctx, double *phi, double *grad, int nigrad);   // fc_assemble.cu
int fc_bpres_dev(fc_context *ctx, double *p, const double *dPdxi, int istage);       // fc_assemble.cu

namespace {


__global__ void __launch_bounds__(256) k_lsq_matrix(fcm_geom g, fcm_c2f m, int weighted, double *dmat) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < g.n) fcg_lsq_matrix_row(g, m, weighted, dmat, c);
}
__global__ void __launch_bounds__(256)
k_grad_lsq(fcm_geom g, fcm_c2f m, fcm_slots sl, int weighted, const double *dmat, const double *fi, double *out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < g.n) fcg_grad_lsq_row(g, m, sl, weighted, dmat, fi, out, c);
}
__global__ void __launch_bounds__(128) k_lsq_qr_matrix(fcm_geom g, fcm_c2f m, double *D, int *bad, int npro) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < g.n && fcg_lsq_qr_matrix_row(g, m, D, c, npro)) atomicAdd(bad, 1);
}
__global__ void __launch_bounds__(256)
k_grad_lsq_qr(fcm_geom g, fcm_c2f m, const double *D, const double *fi, double *out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < g.n) fcg_grad_lsq_qr_row(g, m, D, fi, out, c);
}

// u, v, w (and p) in one walk over the map (fcg_gaussn_row)
template <int NF, bool HAS_OLD>
__global__ void __launch_bounds__(256) k_grad_passn(fcm_geom g, fcm_c2f m, fcg_gaussn k) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < g.n) fcg_gaussn_row<NF, HAS_OLD>(g, m, k, c);
}

// glomin / glomax = minval / maxval(phi(1:numCells)): min and max do not depend on the order, so a plain
// two-stage tree gives the reference's values exactly
__global__ void __launch_bounds__(256) k_minmax_partial(int n, const double *phi, double *part) {
  __shared__ double s_lo[256], s_hi[256];
  double lo = phi[0], hi = phi[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = phi[i];
    lo = v < lo ? v : lo;
    hi = v > hi ? v : hi;
  }
  s_lo[threadIdx.x] = lo; s_hi[threadIdx.x] = hi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_lo[threadIdx.x] = s_lo[threadIdx.x + o] < s_lo[threadIdx.x] ? s_lo[threadIdx.x + o] : s_lo[threadIdx.x];
      s_hi[threadIdx.x] = s_hi[threadIdx.x + o] > s_hi[threadIdx.x] ? s_hi[threadIdx.x + o] : s_hi[threadIdx.x];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[blockIdx.x] = s_lo[0]; part[gridDim.x + blockIdx.x] = s_hi[0]; }
}
__global__ void k_minmax_final(int nb, const double *part, double *out) {
  double lo = part[0], hi = part[nb];
  for (int b = 1; b < nb; ++b) {
    lo = part[b] < lo ? part[b] : lo;
    hi = part[nb + b] > hi ? part[nb + b] : hi;
  }
  out[0] = lo; out[1] = hi;
}
__global__ void __launch_bounds__(256)
k_limiter(fcm_geom g, const int *ioffset, const int *ja, const int *diag, int which, const double *phi, double *grad,
          const double *minmax, double small) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < g.n) fcg_limiter_row(g, ioffset, ja, diag, which, phi, grad, minmax[0], minmax[1], small, c);
}
// several ranks: minmax = {-glomin, glomax} after the all-reduce (one ncclMax covers both)
__global__ void __launch_bounds__(256)
k_limiter_par(fcm_geom g, fcm_c2f m, int npro, const int *ioffset, const int *ja, const int *diag, int which,
              const double *phi, double *grad, const double *minmax, double small) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < g.n) fcg_limiter_row_par(g, m, npro, ioffset, ja, diag, which, phi, grad, -minmax[0], minmax[1], small, c);
}
__global__ void k_negate_first(double *v) { v[0] = -v[0]; }

}  // namespace

// gradients.f90:133-148: the limiter the `input` file selected, applied to a freshly computed gradient
int fc_limit_gradient_dev(fc_context *ctx, const double *phi, double *grad) {
  if (!ctx->grad_limiter) return FC_OK;
  const int nb = 512;
  k_minmax_partial<<<nb, 256, 0, ctx->stream>>>(ctx->n, phi, ctx->partials);   // partials holds >= 2 * 512 doubles
  FC_LAUNCH_CHECK();
  k_minmax_final<<<1, 1, 0, ctx->stream>>>(nb, ctx->partials, &ctx->sc->aux[0]);
  FC_LAUNCH_CHECK();
  if (ctx->nranks > 1) {   // src-parallel/gradients.f90: global_min / global_max, then the parallel build's limiters
    k_negate_first<<<1, 1, 0, ctx->stream>>>(&ctx->sc->aux[0]);
    FC_LAUNCH_CHECK();
    FC_CHECK(fc_allreduce_max(ctx, &ctx->sc->aux[0], 2));
    k_limiter_par<<<fc_blocks(ctx->n, 256), 256, 0, ctx->stream>>>(fcm_geom_of(ctx), fcm_c2f_of(ctx), ctx->npro, ctx->ioffset,
                                                                   ctx->ja, ctx->diag, ctx->grad_limiter, phi, grad,
                                                                   &ctx->sc->aux[0], ctx->grad_small);
    FC_LAUNCH_CHECK();
    return FC_OK;
  }
  k_limiter<<<fc_blocks(ctx->n, 256), 256, 0, ctx->stream>>>(fcm_geom_of(ctx), ctx->ioffset, ctx->ja, ctx->diag,
                                                             ctx->grad_limiter, phi, grad, &ctx->sc->aux[0],
                                                             ctx->grad_small);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

// grad(phi,dPhidxi) (gradients.f90:95-151; several ranks: src-parallel/gradients.f90:95-160 -- exchange(phi) first, the
// three gradient components exchanged after the limiter)
int fc_grad_dev(fc_context *ctx, double *phi, double *grad, int nigrad) {
  if (ctx->grad_method == 0) {
    FC_CHECK(fc_grad_gauss_dev(ctx, phi, grad, nigrad));   // exchanges phi before and the gradient after
  } else {
    if (!ctx->has_mesh || !ctx->c2f_off) FC_FAIL(FC_ERR_ARG, "fc_grad: call fc_set_mesh and fc_create_csr first");
    if (ctx->npro > 0) FC_CHECK(fc_halo_exchange(ctx, phi));
    const int B = 256, G = fc_blocks(ctx->n, B);
    if (ctx->grad_method == 2)
      k_grad_lsq_qr<<<G, B, 0, ctx->stream>>>(fcm_geom_of(ctx), fcm_c2f_of(ctx), ctx->dmatqr, phi, grad);
    else
      k_grad_lsq<<<G, B, 0, ctx->stream>>>(fcm_geom_of(ctx), fcm_c2f_of(ctx), fcm_slots_of(ctx), ctx->grad_method == 3 ? 1 : 0,
                                           ctx->dmat, phi, grad);
    FC_LAUNCH_CHECK();
  }
  FC_CHECK(fc_limit_gradient_dev(ctx, phi, grad));
  if (ctx->npro > 0 && (ctx->grad_method != 0 || ctx->grad_limiter != 0)) FC_CHECK(fc_halo_exchange3(ctx, grad));
  return FC_OK;
}

// grad(U), grad(V), grad(W) as calcuvw (:59-61) and calcp (:38-40) ask for them, optionally followed by the first stage
// of calcPressDiv (bpres(p,1); grad(p), fieldManipulation.f90:82-87) which does not depend on them.  With
// FC_TUNE_FUSED_GRAD and plain Gauss gradients the fields share one kernel per pass; otherwise one dispatcher call each.
int fc_grad_uvw_dev(fc_context *ctx, int nigrad, bool with_p_stage1) {
  double **fl = ctx->field;
  double *phi[FCG_MAXF] = {fl[FC_U], fl[FC_V], fl[FC_W], fl[FC_P]};
  double *grad[FCG_MAXF] = {fl[FC_DUDXI], fl[FC_DVDXI], fl[FC_DWDXI], fl[FC_DPDXI]};
  const int nf = with_p_stage1 ? 4 : 3;
  if (!ctx->tune_fused_grad || ctx->grad_method != 0 || ctx->grad_limiter != 0 || !ctx->has_mesh || !ctx->c2f_off) {
    for (int t = 0; t < 3; ++t) FC_CHECK(fc_grad_dev(ctx, phi[t], grad[t], nigrad));
    if (with_p_stage1) {
      FC_CHECK(fc_bpres_dev(ctx, phi[3], grad[3], 1));
      FC_CHECK(fc_grad_dev(ctx, phi[3], grad[3], nigrad));
    }
    return FC_OK;
  }
  if (nigrad < 1) FC_FAIL(FC_ERR_ARG, "fc_grad_gauss: nigrad < 1");
  // bpres(p,1) only copies owner values into the boundary slots of p: it commutes with the velocity gradients
  if (with_p_stage1) FC_CHECK(fc_bpres_dev(ctx, phi[3], grad[3], 1));
  const size_t g3 = 3 * (size_t)ctx->NP;
  if (nigrad > 1 && !ctx->gtmp3) FC_CHECK(fc_dev_alloc(ctx, &ctx->gtmp3, FCG_MAXF * g3));
  const int B = 256, G = fc_blocks(ctx->n, B);
  if (ctx->npro > 0)
    for (int t = 0; t < nf; ++t) FC_CHECK(fc_halo_exchange(ctx, phi[t]));
  fcg_gaussn k{};
  k.npro = ctx->npro;
  k.fpro = ctx->fpro;
  for (int t = 0; t < nf; ++t) { k.phi[t] = phi[t]; k.old[t] = nullptr; k.out[t] = grad[t]; }
  for (int lc = 1; lc <= nigrad; ++lc) {
    if (lc == 1) {
      if (nf == 4) k_grad_passn<4, false><<<G, B, 0, ctx->stream>>>(fcm_geom_of(ctx), fcm_c2f_of(ctx), k);
      else k_grad_passn<3, false><<<G, B, 0, ctx->stream>>>(fcm_geom_of(ctx), fcm_c2f_of(ctx), k);
    } else {
      for (int t = 0; t < nf; ++t) {
        if (ctx->npro > 0) FC_CHECK(fc_halo_exchange3(ctx, grad[t]));
        FC_CUDA(cudaMemcpyAsync(ctx->gtmp3 + t * g3, grad[t], sizeof(double) * g3, cudaMemcpyDeviceToDevice, ctx->stream));
        k.old[t] = ctx->gtmp3 + t * g3;
      }
      if (nf == 4) k_grad_passn<4, true><<<G, B, 0, ctx->stream>>>(fcm_geom_of(ctx), fcm_c2f_of(ctx), k);
      else k_grad_passn<3, true><<<G, B, 0, ctx->stream>>>(fcm_geom_of(ctx), fcm_c2f_of(ctx), k);
    }
    FC_LAUNCH_CHECK();
  }
  if (ctx->npro > 0)
    for (int t = 0; t < nf; ++t) FC_CHECK(fc_halo_exchange3(ctx, grad[t]));
  return FC_OK;
}

// lstsq / lstsq_qr / lstsq_dm / gauss flags + `limiter` of the input file; builds the geometric matrices
// (create_lsq_gradients_matrix, gradients.f90:65-90)
int fc_set_gradient_dev(fc_context *ctx, int method, int limiter, double small) {
  if (method < 0 || method > 3 || limiter < 0 || limiter > 3) FC_FAIL(FC_ERR_ARG, "fc_set_gradient: unknown method / limiter");
  if (method != 0 || limiter != 0) {
    if (!ctx->has_mesh || !ctx->has_csr || !ctx->c2f_off)
      FC_FAIL(FC_ERR_ARG, "fc_set_gradient: call fc_set_mesh and fc_create_csr first");
    if ((ctx->npro > 0 || ctx->nranks > 1) && (method == 1 || method == 3))
      FC_FAIL(FC_ERR_UNSUPPORTED, "fc_set_gradient: lstsq / lstsq_dm run on one rank in this version (on several ranks: "
                                  "gauss or lstsq_qr, with any limiter)");
  }
  const int B = 256, G = fc_blocks(ctx->n, B);
  if (method == 1 || method == 3) {
    FC_CHECK(fc_dev_alloc(ctx, &ctx->dmat, 9 * (size_t)ctx->n));
    k_lsq_matrix<<<G, B, 0, ctx->stream>>>(fcm_geom_of(ctx), fcm_c2f_of(ctx), method == 3 ? 1 : 0, ctx->dmat);
    FC_LAUNCH_CHECK();
  } else if (method == 2) {
    FC_CHECK(fc_dev_alloc(ctx, &ctx->dmatqr, 18 * (size_t)ctx->n));
    int *bad = nullptr, host_bad = 0;
    FC_CUDA(cudaMalloc((void **)&bad, sizeof(int)));
    FC_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
    k_lsq_qr_matrix<<<fc_blocks(ctx->n, 128), 128, 0, ctx->stream>>>(fcm_geom_of(ctx), fcm_c2f_of(ctx), ctx->dmatqr, bad,
                                                                     ctx->npro);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&host_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(bad);
    if (e != cudaSuccess) FC_FAIL(FC_ERR_CUDA, std::string("fc_set_gradient: ") + cudaGetErrorString(e));
    if (host_bad > 0)
      FC_FAIL(FC_ERR_UNSUPPORTED, "fc_set_gradient: lstsq_qr is defined for cells with exactly 6 neighbours "
                                  "(grad_lsq_qr.f90:31); " + std::to_string(host_bad) + " cells have another count");
  }
  ctx->grad_method = method;
  ctx->grad_limiter = limiter;
  ctx->grad_small = small;
  return FC_OK;
}
