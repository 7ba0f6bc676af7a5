// Krylov solvers of the pressure-correction path: dpcg (src/dpcg.f90), iccg (src/iccg.f90),
// bicgstab (src/bicgstab.f90) and their src-parallel twins.
//
// The recurrences' scalars (sk, s0, pkapk, alf, bet, res0, resl, ...) live in device memory
// (fc_scalars); every kernel reads what it needs from there, and the convergence test
// rsm = resl/(res0+small) < sor is evaluated on the device by whichever thread finishes the
// reduction.  Once `done` is set all later kernels of the batch return immediately, so the
// host only polls once per batch of iterations and the iteration count is exactly the
// reference's.  Vector updates are fused so that one DPCG iteration moves
// 12*nnz + 108*n bytes:  p-update 32n | SpMV + p.Ap 12nnz+20n | x,r update + |r|_1 + next r.z 56n.
#include "fc_reduce.cuh"


namespace {

#define GRID_STRIDE(i, n) for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)(n); i += (size_t)gridDim.x * blockDim.x)

// after convergence the remaining launches of a batch are no-ops, the scalar step included
__global__ void k_scalar_step(fc_scalars *sc, int step, double *hist) {
  if (sc->done && step != STEP_RES0 && step != STEP_RES0_SK) return;
  fc_scalar_step(sc, step, hist);
}

// P2P mode: fold a pending reduction into the scalars when no Krylov kernel follows that would do it
__global__ void k_apply(fc_scalars *sc, fc_sync sy) { fc_kernel_begin(sc, sy); }

__global__ void k_init_scalars(fc_scalars *sc, double sor, double small, int nsw) {
  sc->sor = sor; sc->small = small; sc->nsw = nsw;
  sc->done = 0; sc->iters = 0;
  for (int i = 0; i < 4; ++i) sc->ticket[i] = 0u;
}

// sk = sum res * (res / (a_ii [+small]))  -- first Jacobi inner product (dpcg.f90:85-90)
__global__ void __launch_bounds__(FC_RED_BLOCK)
k_jacobi_sk(int n, const double *__restrict__ res, const double *__restrict__ adiag, double padd, double *partials,
            fc_scalars *sc, fc_sync sy) {
  __shared__ double s_red[32];
  if (!fc_kernel_begin(sc, sy)) return;
  double acc = 0.0;
  GRID_STRIDE(i, n) {
    double r = res[i];
    acc += r * (r / (adiag[i] + padd));
  }
  double v[1] = {acc};
  if (fc_grid_sum<1>(v, partials, &sc->ticket[1], s_red)) fc_reduction_done<1>(sc, sy, v, STEP_SK);
}

// pk = res/(a_ii[+small]) + bet*pk   (dpcg.f90:85-87, 95-100)
__global__ void __launch_bounds__(FC_RED_BLOCK)
k_dpcg_pupdate(int n, const double *__restrict__ res, const double *__restrict__ adiag, double padd,
               double *__restrict__ pk, fc_scalars *sc, fc_sync sy) {
  if (!fc_kernel_begin(sc, sy)) return;
  const double bet = sc->sk / sc->s0;
  GRID_STRIDE(i, n) pk[i] = res[i] / (adiag[i] + padd) + bet * pk[i];
}

// fi += alf*pk ; res -= alf*zk ; resl = sum|res| ; next sk = sum res*res/(a_ii[+small])   (dpcg.f90:121-130)
__global__ void __launch_bounds__(FC_RED_BLOCK)
k_dpcg_update(int n, double *__restrict__ fi, const double *__restrict__ pk, double *__restrict__ res,
              const double *__restrict__ zk, const double *__restrict__ adiag, double padd, double *partials,
              fc_scalars *sc, fc_sync sy) {
  __shared__ double s_red[64];
  if (!fc_kernel_begin(sc, sy)) return;
  const double alf = sc->sk / sc->pkapk;
  double a0 = 0.0, a1 = 0.0;
  GRID_STRIDE(i, n) {
    fi[i] = fi[i] + alf * pk[i];
    double r = res[i] - alf * zk[i];
    res[i] = r;
    a0 += fabs(r);
    a1 += r * (r / (adiag[i] + padd));
  }
  double v[2] = {a0, a1};
  if (fc_grid_sum<2>(v, partials, &sc->ticket[1], s_red)) fc_reduction_done<2>(sc, sy, v, STEP_CG_UPDATE_SK);
}

// generic dot products: red[0] = sum x*y (and red[1] = sum x*z if z)
__global__ void __launch_bounds__(FC_RED_BLOCK)
k_dot(int n, const double *__restrict__ x, const double *__restrict__ y, double *partials, fc_scalars *sc, int step,
      fc_sync sy) {
  __shared__ double s_red[32];
  if (!fc_kernel_begin(sc, sy)) return;
  double acc = 0.0;
  GRID_STRIDE(i, n) acc += x[i] * y[i];
  double v[1] = {acc};
  if (fc_grid_sum<1>(v, partials, &sc->ticket[1], s_red)) fc_reduction_done<1>(sc, sy, v, step);
}

// pk = zk + bet*pk   (iccg.f90:121-126)
__global__ void __launch_bounds__(FC_RED_BLOCK)
k_cg_pupdate(int n, const double *__restrict__ zk, double *__restrict__ pk, fc_scalars *sc, fc_sync sy) {
  if (!fc_kernel_begin(sc, sy)) return;
  const double bet = sc->sk / sc->s0;
  GRID_STRIDE(i, n) pk[i] = zk[i] + bet * pk[i];
}

// fi += alf*pk ; res -= alf*zk ; resl = sum|res|   (iccg.f90:156-165)
__global__ void __launch_bounds__(FC_RED_BLOCK)
k_cg_update(int n, double *__restrict__ fi, const double *__restrict__ pk, double *__restrict__ res,
            const double *__restrict__ zk, double *partials, fc_scalars *sc, fc_sync sy) {
  __shared__ double s_red[32];
  if (!fc_kernel_begin(sc, sy)) return;
  const double alf = sc->sk / sc->pkapk;
  double a0 = 0.0;
  GRID_STRIDE(i, n) {
    fi[i] = fi[i] + alf * pk[i];
    double r = res[i] - alf * zk[i];
    res[i] = r;
    a0 += fabs(r);
  }
  double v[1] = {a0};
  if (fc_grid_sum<1>(v, partials, &sc->ticket[1], s_red)) fc_reduction_done<1>(sc, sy, v, STEP_CG_UPDATE);
}

// pk = res + om*(pk - alf*uk)   (bicgstab.f90:113-115)
__global__ void __launch_bounds__(FC_RED_BLOCK)
k_bi_pupdate(int n, const double *__restrict__ res, double *__restrict__ pk, const double *__restrict__ uk,
             fc_scalars *sc, fc_sync sy) {
  if (!fc_kernel_begin(sc, sy)) return;
  const double om = sc->om, alf = sc->alf;
  GRID_STRIDE(i, n) pk[i] = res[i] + om * (pk[i] - alf * uk[i]);
}

// fi += gam*zk ; res -= gam*uk   (bicgstab.f90:159-162)
__global__ void __launch_bounds__(FC_RED_BLOCK)
k_bi_half(int n, double *__restrict__ fi, const double *__restrict__ zk, double *__restrict__ res,
          const double *__restrict__ uk, fc_scalars *sc, fc_sync sy) {
  if (!fc_kernel_begin(sc, sy)) return;
  const double gam = sc->gam;
  GRID_STRIDE(i, n) {
    fi[i] = fi[i] + gam * zk[i];
    res[i] = res[i] - gam * uk[i];
  }
}

// fi += alf*zk ; res -= alf*vk ; resl = sum|res|   (bicgstab.f90:208-217)
__global__ void __launch_bounds__(FC_RED_BLOCK)
k_bi_update(int n, double *__restrict__ fi, const double *__restrict__ zk, double *__restrict__ res,
            const double *__restrict__ vk, double *partials, fc_scalars *sc, fc_sync sy) {
  __shared__ double s_red[32];
  if (!fc_kernel_begin(sc, sy)) return;
  const double alf = sc->alf;
  double a0 = 0.0;
  GRID_STRIDE(i, n) {
    fi[i] = fi[i] + alf * zk[i];
    double r = res[i] - alf * vk[i];
    res[i] = r;
    a0 += fabs(r);
  }
  double v[1] = {a0};
  if (fc_grid_sum<1>(v, partials, &sc->ticket[1], s_red)) fc_reduction_done<1>(sc, sy, v, STEP_BI_UPDATE);
}

inline int vec_grid(int n) {
  int g = fc_blocks((size_t)n, FC_RED_BLOCK);
  if (g > FC_RED_GRID) g = FC_RED_GRID;
  return g < 1 ? 1 : g;
}

// Host side of the reduction hand-over.  Three modes:
//   single rank  the finishing thread of the producing kernel runs the scalar step itself;
//   NCCL         producing kernel -> ncclAllReduce(red) -> k_scalar_step   (global_sum of src-parallel);
//   P2P          producing kernel posts to the peers' mailboxes, the NEXT kernel folds them in
//                (`pending`); flush() does it with a 1-CTA kernel when no Krylov kernel follows.
struct flow_t {
  fc_context *ctx;
  double *hist;

  // descriptor for the next launch; `step`/`count` describe the reduction it produces (0: none)
  fc_sync next(int step = 0, int count = 0) {
    fc_sync s{};
    s.hist = hist;
    s.local = ctx->nranks == 1;
    if (ctx->p2p) {
      s.p2p = ctx->p2p_dev;
      if (ctx->pending.seq) {
        s.wait_seq = ctx->pending.seq; s.wait_step = ctx->pending.step; s.wait_count = ctx->pending.count;
        ctx->pending = {0, 0, 0};
      }
      if (step) {
        s.post_seq = ++ctx->red_seq;
        ctx->pending = {s.post_seq, step, count};
      }
    }
    return s;
  }
  // after a producing launch
  int reduced(int step, int count) {
    if (ctx->nranks == 1 || ctx->p2p) return FC_OK;
    FC_CHECK(fc_allreduce_scalars(ctx, ctx->sc->red, count));
    k_scalar_step<<<1, 1, 0, ctx->stream>>>(ctx->sc, step, hist);
    FC_LAUNCH_CHECK();
    return FC_OK;
  }
  int flush() {
    if (!ctx->p2p || !ctx->pending.seq) return FC_OK;
    k_apply<<<1, 32, 0, ctx->stream>>>(ctx->sc, next());
    FC_LAUNCH_CHECK();
    return FC_OK;
  }
  // halo of a Krylov vector before the SpMV (call exchange(pk), src-parallel/dpcg.f90:114)
  int exchange(double *x) {
    if (ctx->npro == 0) return FC_OK;
    if (ctx->p2p) return fc_p2p_pack(ctx, x);
    return fc_halo_exchange(ctx, x);
  }
};

int poll(fc_context *ctx) {
  FC_CUDA(cudaMemcpyAsync(ctx->sc_host, ctx->sc, sizeof(fc_scalars), cudaMemcpyDeviceToHost, ctx->stream));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  return FC_OK;
}

}  // namespace

int fc_alloc_solver_scratch(fc_context *ctx) {
  const size_t np = (size_t)ctx->n + (size_t)ctx->npro;
  if (ctx->scratch_n >= np && ctx->pk) return FC_OK;
  FC_CHECK(fc_dev_alloc(ctx, &ctx->pk, np));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->zk, np));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->dd, np));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->reso, np));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->uk, np));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->vk, np));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->adiag, np));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->tt, np));
  ctx->scratch_n = np;
  return FC_OK;
}

// Solve A fi = su with the resident pattern, FC_A and FC_SU; `fi` has >= numCells(+npro) entries.
int fc_solve_device(fc_context *ctx, int solver, double *fi, const fc_solver_opts *o, fc_solver_report *rep,
                    double *hist_host) {
  if (!ctx->has_csr) FC_FAIL(FC_ERR_ARG, "fc_solve: no CSR pattern (call fc_create_csr or fc_solve_csr)");
  if (solver < FC_DPCG || solver > FC_BICGSTAB) FC_FAIL(FC_ERR_ARG, "fc_solve: unknown solver");
  FC_CHECK(fc_alloc_solver_scratch(ctx));
  const int n = ctx->n;
  const int g = vec_grid(n);
  const double *a = ctx->field[FC_A], *su = ctx->field[FC_SU];
  double *res = ctx->field[FC_RES];
  double *pk = ctx->pk, *zk = ctx->zk, *d = ctx->dd;
  const double padd = o->parallel ? o->small : 0.0;
  cudaStream_t st = ctx->stream;
  // residual history: a context-owned buffer (grown on demand, freed with the context), so that no early return
  // of this function can leak it
  double *hist = nullptr;
  if (hist_host && o->nsw > 0) {
    if (ctx->hist_cap < (size_t)o->nsw) {
      cudaFree(ctx->hist);
      ctx->hist = nullptr;
      ctx->hist_cap = 0;
      FC_CHECK(fc_dev_alloc(ctx, &ctx->hist, (size_t)o->nsw));
      ctx->hist_cap = (size_t)o->nsw;
    }
    hist = ctx->hist;
  }

  ctx->spmv_sampled = 0;
  FC_CUDA(cudaEventRecord(ctx->ev[0], st));
  k_init_scalars<<<1, 1, 0, st>>>(ctx->sc, o->sor, o->small, o->nsw);
  FC_LAUNCH_CHECK();
  flow_t fl{ctx, hist};
  ctx->pending = {0, 0, 0};
  ctx->halo_wait = 0;
  ctx->tm.persist_ms = ctx->tm.persist_pupdate_ms = ctx->tm.persist_spmv_ms = ctx->tm.persist_update_ms = 0.0;
  ctx->tm.persist_iters = ctx->tm.persist_grid = ctx->tm.persist_index_bytes = 0;
  ctx->tm.persist_mail_ms = 0.0;
  if (solver == FC_DPCG) {   // the whole solve as one persistent kernel (fc_dpcg_persist.cu)
    bool handled = false;
    FC_CHECK(fc_dpcg_persistent(ctx, fi, o, rep, hist, &handled));
    if (handled) {
      FC_CUDA(cudaEventRecord(ctx->ev[1], st));
      if (ctx->npro > 0 && !(o->tol >= 0.0 && rep->res0 < o->tol))
        FC_CHECK(fc_halo_exchange(ctx, fi));  // src-parallel/dpcg.f90:173
      FC_CUDA(cudaStreamSynchronize(st));
      if (hist && rep->iters > 0)
        FC_CUDA(cudaMemcpy(hist_host, hist, sizeof(double) * (size_t)rep->iters, cudaMemcpyDeviceToHost));
      float ms = 0.f;
      FC_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
      ctx->tm.solve_ms = ms;
      return FC_OK;
    }
  }
  // res = su - A fi, res0 = sum|res|   (dpcg.f90:51-64; the parallel twin uses fi's halo as it is)
  FC_CHECK(fc_launch_residual(ctx, a, su, fi, res, ctx->adiag, fl.next(STEP_RES0, 1)));
  FC_CHECK(fl.reduced(STEP_RES0, 1));
  FC_CHECK(fl.flush());
  FC_CHECK(poll(ctx));
  rep->res0 = ctx->sc_host->res0;
  rep->resl = ctx->sc_host->res0;
  rep->iters = 0;
  if (o->tol >= 0.0 && rep->res0 < o->tol) {  // dpcg.f90:66-70
    FC_CUDA(cudaEventRecord(ctx->ev[1], st));
    FC_CUDA(cudaEventSynchronize(ctx->ev[1]));
    float ms0 = 0.f;
    FC_CUDA(cudaEventElapsedTime(&ms0, ctx->ev[0], ctx->ev[1]));
    ctx->tm.solve_ms = ms0;
    ctx->tm.spmv_samples = 0;
    return FC_OK;
  }
  FC_CUDA(cudaMemsetAsync(pk, 0, sizeof(double) * ((size_t)n + ctx->npro), st));
  if (solver != FC_DPCG) {
    FC_CHECK(fc_levels_build(ctx));
    FC_CHECK(fc_levels_reset(ctx));
    FC_CHECK(fc_precond_factor(ctx, solver == FC_BICGSTAB ? 2 : (o->parallel ? 1 : 0), a, d, padd));
  }
  if (solver == FC_DPCG) {
    k_jacobi_sk<<<g, FC_RED_BLOCK, 0, st>>>(n, res, ctx->adiag, padd, ctx->partials, ctx->sc, fl.next(STEP_SK, 1));
    FC_LAUNCH_CHECK();
    FC_CHECK(fl.reduced(STEP_SK, 1));
  } else if (solver == FC_BICGSTAB) {
    FC_CUDA(cudaMemcpyAsync(ctx->reso, res, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    FC_CUDA(cudaMemsetAsync(ctx->uk, 0, sizeof(double) * (size_t)n, st));
  }

  // iterations enqueued before the host looks at `done`: a solve that ends after one or two iterations (the momentum
  // predictor's bicgstab calls) should not enqueue eight; start from what the last solve of this kind needed
  int batch = ctx->first_batch[solver];
  int launched = 0;
  while (launched < o->nsw) {
    const int todo = (o->nsw - launched) < batch ? (o->nsw - launched) : batch;
    for (int it = 0; it < todo; ++it) {
      if (solver == FC_DPCG) {
        k_dpcg_pupdate<<<g, FC_RED_BLOCK, 0, st>>>(n, res, ctx->adiag, padd, pk, ctx->sc, fl.next());
        FC_LAUNCH_CHECK();
        FC_CHECK(fl.exchange(pk));
        FC_CHECK(fc_launch_spmv_dots(ctx, a, pk, zk, pk, 0, STEP_PKAPK, fl.next(STEP_PKAPK, 1)));
        FC_CHECK(fl.reduced(STEP_PKAPK, 1));
        k_dpcg_update<<<g, FC_RED_BLOCK, 0, st>>>(n, fi, pk, res, zk, ctx->adiag, padd, ctx->partials, ctx->sc,
                                                  fl.next(STEP_CG_UPDATE_SK, 2));
        FC_LAUNCH_CHECK();
        FC_CHECK(fl.reduced(STEP_CG_UPDATE_SK, 2));
      } else if (solver == FC_ICCG) {
        FC_CHECK(fl.flush());   // the sweeps read `done`: the end-of-iteration reduction must be folded in first
        FC_CHECK(fc_precond_apply(ctx, a, d, res, ctx->tt, zk, o->small));
        k_dot<<<g, FC_RED_BLOCK, 0, st>>>(n, res, zk, ctx->partials, ctx->sc, STEP_SK, fl.next(STEP_SK, 1));
        FC_LAUNCH_CHECK();
        FC_CHECK(fl.reduced(STEP_SK, 1));
        k_cg_pupdate<<<g, FC_RED_BLOCK, 0, st>>>(n, zk, pk, ctx->sc, fl.next());
        FC_LAUNCH_CHECK();
        FC_CHECK(fl.exchange(pk));
        FC_CHECK(fc_launch_spmv_dots(ctx, a, pk, zk, pk, 0, STEP_PKAPK, fl.next(STEP_PKAPK, 1)));
        FC_CHECK(fl.reduced(STEP_PKAPK, 1));
        k_cg_update<<<g, FC_RED_BLOCK, 0, st>>>(n, fi, pk, res, zk, ctx->partials, ctx->sc, fl.next(STEP_CG_UPDATE, 1));
        FC_LAUNCH_CHECK();
        FC_CHECK(fl.reduced(STEP_CG_UPDATE, 1));
      } else {
        double *uk = ctx->uk, *vk = ctx->vk, *reso = ctx->reso, *t = ctx->tt;
        k_dot<<<g, FC_RED_BLOCK, 0, st>>>(n, res, reso, ctx->partials, ctx->sc, STEP_BET, fl.next(STEP_BET, 1));
        FC_LAUNCH_CHECK();
        FC_CHECK(fl.reduced(STEP_BET, 1));
        k_bi_pupdate<<<g, FC_RED_BLOCK, 0, st>>>(n, res, pk, uk, ctx->sc, fl.next());
        FC_LAUNCH_CHECK();
        FC_CHECK(fc_precond_apply(ctx, a, d, pk, t, zk, o->small));
        FC_CHECK(fl.exchange(zk));
        FC_CHECK(fc_launch_spmv_dots(ctx, a, zk, uk, reso, 0, STEP_UKRESO, fl.next(STEP_UKRESO, 1)));
        FC_CHECK(fl.reduced(STEP_UKRESO, 1));
        k_bi_half<<<g, FC_RED_BLOCK, 0, st>>>(n, fi, zk, res, uk, ctx->sc, fl.next());
        FC_LAUNCH_CHECK();
        FC_CHECK(fc_precond_apply(ctx, a, d, res, t, zk, o->small));
        FC_CHECK(fl.exchange(zk));
        FC_CHECK(fc_launch_spmv_dots(ctx, a, zk, vk, res, 1, STEP_VK, fl.next(STEP_VK, 2)));
        FC_CHECK(fl.reduced(STEP_VK, 2));
        k_bi_update<<<g, FC_RED_BLOCK, 0, st>>>(n, fi, zk, res, vk, ctx->partials, ctx->sc, fl.next(STEP_BI_UPDATE, 1));
        FC_LAUNCH_CHECK();
        FC_CHECK(fl.reduced(STEP_BI_UPDATE, 1));
      }
    }
    launched += todo;
    FC_CHECK(fl.flush());
    FC_CHECK(poll(ctx));
    if (ctx->sc_host->done) break;
    if (batch < 32) batch *= 2;
  }
  FC_CUDA(cudaEventRecord(ctx->ev[1], st));
  if (ctx->npro > 0) FC_CHECK(fc_halo_exchange(ctx, fi));  // src-parallel/dpcg.f90:173
  FC_CHECK(poll(ctx));
  rep->resl = ctx->sc_host->resl;
  rep->iters = ctx->sc_host->iters;
  ctx->first_batch[solver] = rep->iters + 1 < 2 ? 2 : (rep->iters + 1 > 8 ? 8 : rep->iters + 1);
  if (hist && rep->iters > 0)
    FC_CUDA(cudaMemcpy(hist_host, hist, sizeof(double) * (size_t)rep->iters, cudaMemcpyDeviceToHost));
  float ms = 0.f;
  FC_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  ctx->tm.solve_ms = ms;
  // SpMV launches that ran after convergence returned immediately: only the first `iters`
  // (x2 for bicgstab) samples are real work
  int real = rep->iters * (solver == FC_BICGSTAB ? 2 : 1);
  if (real > ctx->spmv_sampled) real = ctx->spmv_sampled;
  if (real > 0) {
    double sum = 0.0;
    for (int i = 0; i < real; ++i) {
      float t = 0.f;
      FC_CUDA(cudaEventElapsedTime(&t, ctx->spmv_ev[2 * i], ctx->spmv_ev[2 * i + 1]));
      sum += t;
    }
    ctx->tm.spmv_ms = sum / real;
    ctx->tm.spmv_samples = real;
  }
  return FC_OK;
}
