/*
 * fc_oracle_par.c -- TEST INFRASTRUCTURE ONLY (see fc_oracle.h).
 *
 * src-parallel semantics of the pressure-correction path: R ranks advance in lock step inside
 * ONE process; `exchange` (src-parallel/exchange.f90:3-92) is a memory copy between the ranks'
 * arrays and `global_sum` (global_sum_mpi.f90) adds the ranks' values in rank order.  It is the
 * checker for the multi-GPU (NCCL) path and is compiled as one translation unit with the
 * serial oracle so that both share the same face-flux / gradient helpers.
 */
#ifdef _OPENMP
#include <omp.h>
#endif
#include "fc_oracle.c"

/* The ranks of one lock-step phase are independent (they only read halo values copied in an earlier
 * phase), so a phase may run one rank per host thread: the src-parallel build as R threads instead of
 * R MPI processes.  Sums are still added in rank order, so the results are bit-identical to the
 * single-thread run.  Used by bench.py's reference arm; the tests run it with one thread. */
static int fco_threads = 1;
void fco_par_set_threads(int n) { fco_threads = n > 1 ? n : 1; }
/* `outer` threads over the ranks x `inner` threads inside each rank's row loops (fc_oracle.c, ROWS_MT) */
void fco_par_set_threads2(int outer, int inner) {
  fco_threads = outer > 1 ? outer : 1;
  fco_inner = inner > 1 ? inner : 1;
#ifdef _OPENMP
  omp_set_max_active_levels(fco_threads > 1 && fco_inner > 1 ? 2 : 1);
#endif
}
int fco_par_openmp(void) {
#ifdef _OPENMP
  return 1;
#else
  return 0;
#endif
}
#define PAR_RANKS _Pragma("omp parallel for schedule(static, 1) num_threads(fco_threads) if (fco_threads > 1)")

/* phi_r(iProcStart+i) <- phi_q(bufind_q(i')) for every connection r<->q; `stride` addresses one
 * component of an interleaved (3,numPCells) gradient (exchange(dPhidxi(1,:)) passes a strided
 * section; gfortran packs it into a contiguous temporary, the effect is the same). */
void fco_par_exchange(fco_rank *R, int nr, double **phi, int stride) {
  double **buf = (double **)malloc(sizeof(double *) * (size_t)nr);
  PAR_RANKS for (int r = 0; r < nr; ++r) {
    const fco_mesh *g = &R[r].g;
    buf[r] = (double *)malloc(sizeof(double) * (size_t)(g->npro > 0 ? g->npro : 1));
    for (int i = 1; i <= g->npro; ++i) /* buffer(i) = phi(bufind(i)), bufind(i) = owner(iProcFacesStart+i) */
      buf[r][i - 1] = phi[r][(size_t)(A1(g->owner, g->iProcFacesStart + i) - 1) * stride];
  }
  PAR_RANKS for (int r = 0; r < nr; ++r) {
    const fco_mesh *g = &R[r].g;
    for (int c = 0; c < R[r].numConnections; ++c) {
      const int q = R[r].neighbProcNo[c];
      const int s = R[r].neighbProcOffset[c], e = R[r].neighbProcOffset[c + 1]; /* 1-based [s, e) */
      int cq = -1;
      for (int k = 0; k < R[q].numConnections; ++k)
        if (R[q].neighbProcNo[k] == r) cq = k;
      const int sq = R[q].neighbProcOffset[cq];
      for (int i = s; i < e; ++i) phi[r][(size_t)(g->numCells + i - 1) * stride] = buf[q][sq + (i - s) - 1];
    }
  }
  for (int r = 0; r < nr; ++r) free(buf[r]);
  free(buf);
}

static double gsum(const double *v, int nr) {
  double s = 0.0;
  for (int r = 0; r < nr; ++r) s = s + v[r];
  return s;
}

/* grad(phi,dPhidxi) of src-parallel/gradients.f90:95-160: exchange(phi), grad_gauss with the
 * processor-face loop, exchange of the three components.  For nigrad > 1 the reference indexes
 * its numCells-sized old-gradient copy with halo cells (out of bounds); here the old gradient is
 * exchanged between passes, the only well-defined reading. */
void fco_par_grad_gauss(fco_rank *R, int nr, double **phi, int nigrad, double **grad) {
  fco_par_exchange(R, nr, phi, 1);
  double **dfo = (double **)malloc(sizeof(double *) * (size_t)nr);
  double **comp = (double **)malloc(sizeof(double *) * (size_t)nr);
  for (int r = 0; r < nr; ++r)
    dfo[r] = (double *)calloc(3 * (size_t)(R[r].g.numCells + R[r].g.npro), sizeof(double));
  for (int lc = 1; lc <= nigrad; ++lc) {
    for (int r = 0; r < nr; ++r) grad_pass(&R[r].g, phi[r], dfo[r], grad[r]);
    if (lc != nigrad) {
      for (int r = 0; r < nr; ++r)
        memcpy(dfo[r], grad[r], sizeof(double) * 3 * (size_t)(R[r].g.numCells + R[r].g.npro));
      for (int c = 0; c < 3; ++c) {
        for (int r = 0; r < nr; ++r) comp[r] = dfo[r] + c;
        fco_par_exchange(R, nr, comp, 3);
      }
    }
  }
  for (int c = 0; c < 3; ++c) {
    for (int r = 0; r < nr; ++r) comp[r] = grad[r] + c;
    fco_par_exchange(R, nr, comp, 3);
  }
  for (int r = 0; r < nr; ++r) free(dfo[r]);
  free(dfo);
  free(comp);
}

void fco_par_grad_gauss_corrected(fco_rank *R, int nr, double **phi, double **grad) {
  fco_par_exchange(R, nr, phi, 1);
  double **comp = (double **)malloc(sizeof(double *) * (size_t)nr);
  for (int r = 0; r < nr; ++r) {
    const size_t n3 = 3 * (size_t)(R[r].g.numCells + R[r].g.npro);
    double *dfo = (double *)malloc(sizeof(double) * n3);
    memcpy(dfo, grad[r], sizeof(double) * n3);
    grad_pass(&R[r].g, phi[r], dfo, grad[r]);
    free(dfo);
  }
  for (int c = 0; c < 3; ++c) {
    for (int r = 0; r < nr; ++r) comp[r] = grad[r] + c;
    fco_par_exchange(R, nr, comp, 3);
  }
  free(comp);
}

/* laplacian of src-parallel/fvm_laplacian.f90: exchange(mu), inner faces, processor faces -> apr, wall BC */
void fco_par_laplacian(fco_rank *R, int nr, double **mu, double **phi) {
  fco_par_exchange(R, nr, mu, 1);
  for (int r = 0; r < nr; ++r) {
    const fco_mesh *g = &R[r].g;
    fco_fields *f = &R[r].f;
    /* inner faces + wall part exactly as the serial routine, but the processor loop sits between them */
    for (int k = 0; k < R[r].m.nnz; ++k) f->a[k] = 0.0;
    for (int i = 1; i <= g->numInnerFaces; ++i) {
      int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
      double cap, can;
      facefluxlaplacian(g, ijp, ijn, A1(g->arx, i), A1(g->ary, i), A1(g->arz, i), A1(g->facint, i), mu[r], &cap, &can);
      A1(f->a, A1(R[r].m.icell_jcell, i)) = can;
      A1(f->a, A1(R[r].m.jcell_icell, i)) = cap;
      A1(f->a, A1(R[r].m.diag, ijp)) = A1(f->a, A1(R[r].m.diag, ijp)) - can;
      A1(f->a, A1(R[r].m.diag, ijn)) = A1(f->a, A1(R[r].m.diag, ijn)) - cap;
    }
    for (int i = 1; i <= g->npro; ++i) { /* :93-114 */
      int iface = g->iProcFacesStart + i, ijp = A1(g->owner, iface), ijn = g->numCells + i;
      double cap, can;
      facefluxlaplacian(g, ijp, ijn, A1(g->arx, iface), A1(g->ary, iface), A1(g->arz, iface), A1(g->fpro, i), mu[r],
                        &cap, &can);
      A1(R[r].apr, i) = can;
      A1(f->a, A1(R[r].m.diag, ijp)) = A1(f->a, A1(R[r].m.diag, ijp)) - can;
    }
    const int iWallStart = g->numCells + g->npro + g->ninl + g->nout + g->nsym;
    for (int i = 1; i <= g->nwal; ++i) {
      int iface = g->iWallFacesStart + i, ijp = A1(g->owner, iface), ijb = iWallStart + i;
      int k = A1(R[r].m.diag, ijp);
      double ax = A1(g->arx, iface), ay = A1(g->ary, iface), az = A1(g->arz, iface);
      double are = sqrt(ax * ax + ay * ay + az * az);
      double dx = A1(g->xc, ijp) - A1(g->xf, iface), dy = A1(g->yc, ijp) - A1(g->yf, iface),
             dz = A1(g->zc, ijp) - A1(g->zf, iface);
      double dpw = sqrt(dx * dx + dy * dy + dz * dz);
      A1(f->a, k) = A1(f->a, k) - A1(mu[r], ijp) * are / dpw;
      A1(f->su, ijp) = A1(f->su, ijp) + A1(f->a, k) * A1(phi[r], ijb);
    }
  }
}

/* `grad(phi,dPhidxi)` of src-parallel/gradients.f90:95-160 with the dispatcher's options: exchange(phi); the scheme
 * (0 gauss with `nigrad` passes, 2 lstsq_qr with the per-rank matrices Dqr[r] of fco_lsq_qr_matrix); the limiter with
 * glomin / glomax reduced over the ranks; exchange of the three gradient components.  (lstsq / lstsq_dm of the parallel
 * build are not restated: rc 2.) */
int fco_par_grad(fco_rank *R, int nr, int method, int limiter, double small, double **Dqr, double **phi, int nigrad,
                 double **grad) {
  if (method != 0 && method != 2) return 2;
  if (method == 0) {
    fco_par_grad_gauss(R, nr, phi, nigrad, grad);   /* exchanges phi before and the gradient after */
  } else {
    fco_par_exchange(R, nr, phi, 1);
    PAR_RANKS for (int r = 0; r < nr; ++r) {
      memset(grad[r], 0, sizeof(double) * 3 * (size_t)(R[r].g.numCells + R[r].g.npro));
      fco_grad_lsq_qr(&R[r].g, Dqr[r], phi[r], grad[r]);
    }
  }
  if (limiter) {
    double glomin = 0.0, glomax = 0.0;
    for (int r = 0; r < nr; ++r) {   /* minval / maxval(phi(1:numCells)), then global_min / global_max */
      const int n = R[r].g.numCells;
      double lo = phi[r][0], hi = phi[r][0];
      for (int i = 1; i < n; ++i) { if (phi[r][i] < lo) lo = phi[r][i]; if (phi[r][i] > hi) hi = phi[r][i]; }
      if (r == 0 || lo < glomin) glomin = lo;
      if (r == 0 || hi > glomax) glomax = hi;
    }
    PAR_RANKS for (int r = 0; r < nr; ++r)
      fco_slope_limiter_par(&R[r].g, &R[r].m, limiter, phi[r], grad[r], small, glomin, glomax);
  }
  if (method != 0 || limiter) {
    double **comp = (double **)malloc(sizeof(double *) * (size_t)nr);
    for (int c = 0; c < 3; ++c) {
      for (int r = 0; r < nr; ++r) comp[r] = grad[r] + c;
      fco_par_exchange(R, nr, comp, 3);
    }
    free(comp);
  }
  return 0;
}

/* ---- Krylov solvers in lock step: src-parallel/dpcg.f90, iccg.f90, bicgstab.f90 ---- */
typedef struct { double *pk, *zk, *d, *reso, *uk, *vk; fco_strips st; int *pown; } par_scratch;

static void rank_strips(const fco_rank *R, par_scratch *s) {
  const fco_mesh *g = &R->g;
  s->pown = (int *)malloc(sizeof(int) * (size_t)(g->npro > 0 ? g->npro : 1));
  for (int i = 1; i <= g->npro; ++i) s->pown[i - 1] = A1(g->owner, g->iProcFacesStart + i);
  fco_strips st = {0, 0, 0, 0, 0, g->npro, s->pown, R->apr, g->numCells};
  s->st = st;
}

int fco_par_solve(fco_rank *R, int nr, int solver, double **fi, const fco_solver_opts *o, fco_report *rep,
                  double *hist) {
  par_scratch *S = (par_scratch *)calloc((size_t)nr, sizeof(par_scratch));
  double *part = (double *)calloc((size_t)nr, sizeof(double)), *part2 = (double *)calloc((size_t)nr, sizeof(double));
  double **vec = (double **)malloc(sizeof(double *) * (size_t)nr);
  PAR_RANKS for (int r = 0; r < nr; ++r) {
    const size_t np = (size_t)(R[r].g.numCells + R[r].g.npro);
    S[r].pk = (double *)calloc(np, sizeof(double)); S[r].zk = (double *)calloc(np, sizeof(double));
    S[r].d = (double *)calloc(np, sizeof(double)); S[r].reso = (double *)calloc(np, sizeof(double));
    S[r].uk = (double *)calloc(np, sizeof(double)); S[r].vk = (double *)calloc(np, sizeof(double));
    rank_strips(&R[r], &S[r]);
  }
  PAR_RANKS for (int r = 0; r < nr; ++r)
    part[r] = initial_residual(&R[r].m, R[r].f.a, R[r].f.su, fi[r], R[r].f.res, &S[r].st);
  const double res0 = gsum(part, nr);
  double resl = res0;
  int used = 0;
  rep->res0 = res0; rep->resl = res0; rep->iters = 0;
  if (o->tol >= 0.0 && res0 < o->tol) goto done;
  if (solver != 0)
    PAR_RANKS for (int r = 0; r < nr; ++r) { /* rank-local DIC / DILU with +small (src-parallel/iccg.f90:94-100, bicgstab.f90:80-91) */
      const fco_csr *m = &R[r].m;
      const double *a = R[r].f.a;
      double *d = S[r].d;
      for (int i = 1; i <= m->n; ++i) {
        double di = A1(a, A1(m->diag, i));
        for (int k = A1(m->ioffset, i); k <= A1(m->diag, i) - 1; ++k) {
          int jc = A1(m->ja, k);
          if (solver == 1) di = di - A1(a, k) * A1(d, jc) * A1(a, k);
          else {
            int l;
            for (l = A1(m->diag, jc); l <= A1(m->ioffset, jc + 1) - 1; ++l)
              if (A1(m->ja, l) == i) break;
            di = di - A1(a, k) * A1(d, jc) * A1(a, l);
          }
        }
        A1(d, i) = 1.0 / (di + o->small);
      }
    }
  if (solver == 2)
    PAR_RANKS for (int r = 0; r < nr; ++r) memcpy(S[r].reso, R[r].f.res, sizeof(double) * (size_t)R[r].m.n);
  {
    double s0 = (double)1.e20f, alf = 1.0, beto = 1.0, gam = 1.0;
    for (int l = 1; l <= o->nsw; ++l) {
      if (solver != 2) {
        PAR_RANKS for (int r = 0; r < nr; ++r) {
          const fco_csr *m = &R[r].m;
          const int n = m->n;
          double *res = R[r].f.res, *zk = S[r].zk;
          if (solver == 0) {
            ROWS_MT for (int i = 1; i <= n; ++i) A1(zk, i) = A1(res, i) / (A1(R[r].f.a, A1(m->diag, i)) + o->small);
          }
          else
            precond_sweeps(m, R[r].f.a, S[r].d, res, zk, o->small);
          double sk = 0.0;
          for (int i = 1; i <= n; ++i) sk = sk + A1(res, i) * A1(zk, i);
          part[r] = sk;
        }
        const double sk = gsum(part, nr);
        const double bet = sk / s0;
        PAR_RANKS for (int r = 0; r < nr; ++r) {
          const int n = R[r].m.n;
          ROWS_MT for (int i = 1; i <= n; ++i) A1(S[r].pk, i) = A1(S[r].zk, i) + bet * A1(S[r].pk, i);
          vec[r] = S[r].pk;
        }
        fco_par_exchange(R, nr, vec, 1);
        PAR_RANKS for (int r = 0; r < nr; ++r) {
          const int n = R[r].m.n;
          matvec(&R[r].m, R[r].f.a, S[r].pk, S[r].zk, &S[r].st, 1);
          double pkapk = 0.0;
          for (int i = 1; i <= n; ++i) pkapk = pkapk + A1(S[r].pk, i) * A1(S[r].zk, i);
          part[r] = pkapk;
        }
        const double pkapk = gsum(part, nr);
        alf = sk / pkapk;
        PAR_RANKS for (int r = 0; r < nr; ++r) {
          const int n = R[r].m.n;
          double *res = R[r].f.res;
          ROWS_MT for (int i = 1; i <= n; ++i) A1(fi[r], i) = A1(fi[r], i) + alf * A1(S[r].pk, i);
          ROWS_MT for (int i = 1; i <= n; ++i) A1(res, i) = A1(res, i) - alf * A1(S[r].zk, i);
          double rl = 0.0;
          for (int i = 1; i <= n; ++i) rl = rl + fabs(A1(res, i));
          part[r] = rl;
        }
        resl = gsum(part, nr);
        s0 = sk;
      } else {
        PAR_RANKS for (int r = 0; r < nr; ++r) {
          const int n = R[r].m.n;
          double b = 0.0;
          for (int i = 1; i <= n; ++i) b = b + A1(R[r].f.res, i) * A1(S[r].reso, i);
          part[r] = b;
        }
        const double bet = gsum(part, nr);
        const double om = bet * gam / (alf * beto + o->small);
        beto = bet;
        PAR_RANKS for (int r = 0; r < nr; ++r) {
          const int n = R[r].m.n;
          for (int i = 1; i <= n; ++i)
            A1(S[r].pk, i) = A1(R[r].f.res, i) + om * (A1(S[r].pk, i) - alf * A1(S[r].uk, i));
          precond_sweeps(&R[r].m, R[r].f.a, S[r].d, S[r].pk, S[r].zk, o->small);
          vec[r] = S[r].zk;
        }
        fco_par_exchange(R, nr, vec, 1);
        PAR_RANKS for (int r = 0; r < nr; ++r) {
          const int n = R[r].m.n;
          matvec(&R[r].m, R[r].f.a, S[r].zk, S[r].uk, &S[r].st, 0);
          double t = 0.0;
          for (int i = 1; i <= n; ++i) t = t + A1(S[r].uk, i) * A1(S[r].reso, i);
          part[r] = t;
        }
        const double ukreso = gsum(part, nr);
        gam = bet / ukreso;
        PAR_RANKS for (int r = 0; r < nr; ++r) {
          const int n = R[r].m.n;
          double *res = R[r].f.res;
          for (int i = 1; i <= n; ++i) A1(fi[r], i) = A1(fi[r], i) + gam * A1(S[r].zk, i);
          for (int i = 1; i <= n; ++i) A1(res, i) = A1(res, i) - gam * A1(S[r].uk, i);
          precond_sweeps(&R[r].m, R[r].f.a, S[r].d, res, S[r].zk, o->small);
          vec[r] = S[r].zk;
        }
        fco_par_exchange(R, nr, vec, 1);
        PAR_RANKS for (int r = 0; r < nr; ++r) {
          const int n = R[r].m.n;
          matvec(&R[r].m, R[r].f.a, S[r].zk, S[r].vk, &S[r].st, 0);
          double t = 0.0, t2 = 0.0;
          for (int i = 1; i <= n; ++i) t = t + A1(S[r].vk, i) * A1(R[r].f.res, i);
          for (int i = 1; i <= n; ++i) t2 = t2 + A1(S[r].vk, i) * A1(S[r].vk, i);
          part[r] = t; part2[r] = t2;
        }
        const double svkres = gsum(part, nr), svkvk = gsum(part2, nr);
        alf = svkres / (svkvk + o->small);
        PAR_RANKS for (int r = 0; r < nr; ++r) {
          const int n = R[r].m.n;
          double *res = R[r].f.res;
          for (int i = 1; i <= n; ++i) A1(fi[r], i) = A1(fi[r], i) + alf * A1(S[r].zk, i);
          for (int i = 1; i <= n; ++i) A1(res, i) = A1(res, i) - alf * A1(S[r].vk, i);
          double rl = 0.0;
          for (int i = 1; i <= n; ++i) rl = rl + fabs(A1(res, i));
          part[r] = rl;
        }
        resl = gsum(part, nr);
      }
      ++used;
      if (hist) hist[l - 1] = resl;
      if (resl / (res0 + o->small) < o->sor) break;
    }
  }
  fco_par_exchange(R, nr, fi, 1); /* call exchange(fi) at exit */
done:
  rep->resl = resl; rep->iters = used;
  for (int r = 0; r < nr; ++r) {
    free(S[r].pk); free(S[r].zk); free(S[r].d); free(S[r].reso); free(S[r].uk); free(S[r].vk); free(S[r].pown);
  }
  free(S); free(part); free(part2); free(vec);
  return 0;
}

/* ---- calcp of src-parallel/calcp-multiple_correction_SIMPLE.f90 ---- */
#define FOR_RANKS for (int r = 0; r < nr; ++r)

static void par_outlet(fco_rank *R, int nr, double flomas, double small, int add_to_su, int global) {
  double *part = (double *)calloc((size_t)nr, sizeof(double));
  FOR_RANKS {
    const fco_mesh *g = &R[r].g;
    fco_fields *f = &R[r].f;
    const int iOutletStart = g->numCells + g->npro + g->ninl;
    double flowo = 0.0;
    for (int i = 1; i <= g->nout; ++i) {
      int iface = g->iOutletFacesStart + i, ijp = A1(g->owner, iface), ijb = iOutletStart + i;
      A1(f->u, ijb) = A1(f->u, ijp); A1(f->v, ijb) = A1(f->v, ijp); A1(f->w, ijb) = A1(f->w, ijp);
      A1(f->fmo, i) = A1(f->den, ijp) * (A1(f->u, ijb) * A1(g->arx, iface) + A1(f->v, ijb) * A1(g->ary, iface) +
                                          A1(f->w, ijb) * A1(g->arz, iface));
      flowo = flowo + A1(f->fmo, i);
    }
    part[r] = flowo;
  }
  const double tot = gsum(part, nr);
  FOR_RANKS {
    const fco_mesh *g = &R[r].g;
    fco_fields *f = &R[r].f;
    const int iOutletStart = g->numCells + g->npro + g->ninl;
    /* adjustMassFlow sums flowo over ranks (:49); correctBoundaryConditionsVelocity does not */
    const double fac = flomas / ((global ? tot : part[r]) + small);
    for (int i = 1; i <= g->nout; ++i) {
      int iface = g->iOutletFacesStart + i, ijp = A1(g->owner, iface), ijb = iOutletStart + i;
      A1(f->fmo, i) = A1(f->fmo, i) * fac;
      A1(f->u, ijb) = A1(f->u, ijb) * fac; A1(f->v, ijb) = A1(f->v, ijb) * fac; A1(f->w, ijb) = A1(f->w, ijb) * fac;
      if (add_to_su) A1(f->su, ijp) = A1(f->su, ijp) - A1(f->fmo, i);
    }
  }
  free(part);
}

/* the `grad` options of the input file for the lock-step routines (fco_par_set_gradient): method 0 gauss / 2 lstsq_qr,
 * limiter 0..3; the QR matrices are built per rank and kept until the next call */
static struct { int method, limiter, nr; double small; double **Dqr; } g_par_grad = {0, 0, 0, 0.0, 0};
int fco_par_grad(fco_rank *R, int nr, int method, int limiter, double small, double **Dqr, double **phi, int nigrad,
                 double **grad);
int fco_par_set_gradient(fco_rank *R, int nr, int method, int limiter, double small) {
  if (g_par_grad.Dqr) {
    for (int r = 0; r < g_par_grad.nr; ++r) free(g_par_grad.Dqr[r]);
    free(g_par_grad.Dqr);
    g_par_grad.Dqr = 0;
  }
  g_par_grad.method = method; g_par_grad.limiter = limiter; g_par_grad.small = small; g_par_grad.nr = nr;
  if (method != 0 && method != 2) return 2;
  if (method == 2) {
    g_par_grad.Dqr = (double **)malloc(sizeof(double *) * (size_t)nr);
    int bad = 0;
    for (int r = 0; r < nr; ++r) {
      g_par_grad.Dqr[r] = (double *)malloc(sizeof(double) * 18 * (size_t)R[r].g.numCells);
      bad += fco_lsq_qr_matrix(&R[r].g, g_par_grad.Dqr[r]);
    }
    if (bad) return 3;
  }
  return 0;
}

static void par_grad_field(fco_rank *R, int nr, int which_phi, int which_grad, int nigrad) {
  double **phi = (double **)malloc(sizeof(double *) * (size_t)nr), **gr = (double **)malloc(sizeof(double *) * (size_t)nr);
  FOR_RANKS {
    fco_fields *f = &R[r].f;
    phi[r] = which_phi == 0 ? f->u : which_phi == 1 ? f->v : which_phi == 2 ? f->w : which_phi == 4 ? f->p : f->pp;
    gr[r] = which_grad == 0 ? f->dUdxi : which_grad == 1 ? f->dVdxi : which_grad == 2 ? f->dWdxi : f->dPdxi;
  }
  if (nigrad > 0 && (g_par_grad.method != 0 || g_par_grad.limiter != 0) && g_par_grad.nr == nr)
    fco_par_grad(R, nr, g_par_grad.method, g_par_grad.limiter, g_par_grad.small, g_par_grad.Dqr, phi, nigrad, gr);
  else if (nigrad > 0) fco_par_grad_gauss(R, nr, phi, nigrad, gr);
  else fco_par_grad_gauss_corrected(R, nr, phi, gr);
  free(phi); free(gr);
}

/* `proc_variant`: flux routine of the processor faces -- 1 facefluxmass2 (calcp :113), 2 facefluxmass_piso (PISO :187) */
static void par_calcp_assemble_v(fco_rank *R, int nr, const fco_calcp_opts *o, int proc_variant) {
  FOR_RANKS {
    for (int k = 0; k < R[r].m.nnz; ++k) R[r].f.a[k] = 0.0;
    for (int i = 0; i < R[r].g.npro; ++i) R[r].apr[i] = 0.0;
    for (int i = 0; i < R[r].g.numCells; ++i) R[r].f.su[i] = 0.0;
  }
  par_grad_field(R, nr, 0, 0, o->nigrad);
  par_grad_field(R, nr, 1, 1, o->nigrad);
  par_grad_field(R, nr, 2, 2, o->nigrad);
  FOR_RANKS {
    const fco_mesh *g = &R[r].g;
    const fco_csr *m = &R[r].m;
    fco_fields *f = &R[r].f;
    for (int i = 1; i <= g->numInnerFaces; ++i) {
      int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
      double cap, can;
      fco_facefluxmass(g, f, o->flux_variant, ijp, ijn, A1(g->xf, i), A1(g->yf, i), A1(g->zf, i), A1(g->arx, i),
                       A1(g->ary, i), A1(g->arz, i), A1(g->facint, i), &cap, &can, &A1(f->flmass, i));
      A1(f->a, A1(m->icell_jcell, i)) = can;
      A1(f->a, A1(m->jcell_icell, i)) = cap;
      A1(f->a, A1(m->diag, ijp)) = A1(f->a, A1(m->diag, ijp)) - can;
      A1(f->a, A1(m->diag, ijn)) = A1(f->a, A1(m->diag, ijn)) - cap;
      A1(f->su, ijp) = A1(f->su, ijp) - A1(f->flmass, i);
      A1(f->su, ijn) = A1(f->su, ijn) + A1(f->flmass, i);
    }
    for (int i = 1; i <= g->npro; ++i) { /* :107-128: facefluxmass2 on processor faces */
      int iface = g->iProcFacesStart + i, ijp = A1(g->owner, iface), ijn = g->numCells + i;
      double cap, can;
      fco_facefluxmass(g, f, proc_variant, ijp, ijn, A1(g->xf, iface), A1(g->yf, iface), A1(g->zf, iface), A1(g->arx, iface),
                       A1(g->ary, iface), A1(g->arz, iface), A1(g->fpro, i), &cap, &can, &A1(R[r].fmpro, i));
      A1(R[r].apr, i) = can;
      A1(f->a, A1(m->diag, ijp)) = A1(f->a, A1(m->diag, ijp)) - can;
      A1(f->su, ijp) = A1(f->su, ijp) - A1(R[r].fmpro, i);
    }
  }
  if (!o->const_mflux) {
    FOR_RANKS {
      const fco_mesh *g = &R[r].g;
      for (int i = 1; i <= g->ninl; ++i) {
        int ijp = A1(g->owner, g->iInletFacesStart + i);
        A1(R[r].f.su, ijp) = A1(R[r].f.su, ijp) - A1(R[r].f.fmi, i);
      }
    }
    par_outlet(R, nr, o->flomas, o->sol.small, 1, 1);
  }
}

void fco_par_calcp_assemble(fco_rank *R, int nr, const fco_calcp_opts *o) { par_calcp_assemble_v(R, nr, o, 1); }

/* correctBoundaryConditionsVelocity of src-parallel: outlet extrapolation scaled with the LOCAL outflow, symmetry */
static void par_correct_bc_velocity(fco_rank *R, int nr, double flomas, double small) {
  par_outlet(R, nr, flomas, small, 0, 0);
  FOR_RANKS {
    const fco_mesh *g = &R[r].g;
    fco_fields *f = &R[r].f;
    const int iSymmetryStart = g->numCells + g->npro + g->ninl + g->nout;
    for (int i = 1; i <= g->nsym; ++i) {
      int iface = g->iSymmetryFacesStart + i, ijp = A1(g->owner, iface), ijb = iSymmetryStart + i;
      double Unmag = A1(f->u, ijp) * A1(g->arx, iface) + A1(f->v, ijp) * A1(g->ary, iface) + A1(f->w, ijp) * A1(g->arz, iface);
      A1(f->u, ijb) = A1(f->u, ijp) - Unmag * A1(g->arx, iface);
      A1(f->v, ijb) = A1(f->v, ijp) - Unmag * A1(g->ary, iface);
      A1(f->w, ijb) = A1(f->w, ijp) - Unmag * A1(g->arz, iface);
    }
  }
}

int fco_par_calcp(fco_rank *R, int nr, const fco_calcp_opts *o, fco_calcp_report *rep) {
  fco_par_calcp_assemble(R, nr, o);
  double **pp = (double **)malloc(sizeof(double *) * (size_t)nr);
  double *part = (double *)calloc((size_t)nr, sizeof(double)), *part2 = (double *)calloc((size_t)nr, sizeof(double));
  FOR_RANKS pp[r] = R[r].f.pp;
  int gloCells = 0;
  FOR_RANKS gloCells += R[r].g.numCells;
  for (int ipcorr = 1; ipcorr <= o->npcor; ++ipcorr) {
    FOR_RANKS for (int i = 0; i < R[r].g.numTotal; ++i) R[r].f.pp[i] = 0.0;
    fco_par_solve(R, nr, o->solver, pp, &o->sol, &rep->rep[ipcorr - 1 < 8 ? ipcorr - 1 : 7], 0);
    for (int istage = 1; istage <= o->nipgrad; ++istage) {
      FOR_RANKS fco_bpres(&R[r].g, R[r].f.pp, R[r].f.dPdxi, istage);
      par_grad_field(R, nr, 3, 3, o->nigrad);
    }
    if (o->lsq_flag) {
      FOR_RANKS memset(R[r].f.dPdxi, 0, sizeof(double) * 3 * (size_t)(R[r].g.numCells + R[r].g.npro));
      par_grad_field(R, nr, 3, 3, 0);
    }
    FOR_RANKS { /* :175-177 ppref = global mean */
      double s = 0.0;
      for (int i = 1; i <= R[r].g.numCells; ++i) s = s + A1(R[r].f.pp, i);
      part[r] = s;
    }
    const double ppref = gsum(part, nr) / gloCells;
    FOR_RANKS {
      const fco_mesh *g = &R[r].g;
      const fco_csr *m = &R[r].m;
      fco_fields *f = &R[r].f;
      for (int iface = 1; iface <= g->numInnerFaces; ++iface) {
        int ijp = A1(g->owner, iface), ijn = A1(g->neighbour, iface);
        A1(f->flmass, iface) = A1(f->flmass, iface) + A1(f->a, A1(m->icell_jcell, iface)) * (A1(f->pp, ijn) - A1(f->pp, ijp));
      }
      for (int i = 1; i <= g->npro; ++i) {
        int ijp = A1(g->owner, g->iProcFacesStart + i);
        A1(R[r].fmpro, i) = A1(R[r].fmpro, i) + A1(R[r].apr, i) * (A1(f->pp, g->numCells + i) - A1(f->pp, ijp));
      }
      for (int inp = 1; inp <= g->numCells; ++inp) {
        A1(f->u, inp) = A1(f->u, inp) - G3(f->dPdxi, 0, inp) * A1(g->vol, inp) * A1(f->apu, inp);
        A1(f->v, inp) = A1(f->v, inp) - G3(f->dPdxi, 1, inp) * A1(g->vol, inp) * A1(f->apv, inp);
        A1(f->w, inp) = A1(f->w, inp) - G3(f->dPdxi, 2, inp) * A1(g->vol, inp) * A1(f->apw, inp);
        A1(f->p, inp) = A1(f->p, inp) + o->urf_p * (A1(f->pp, inp) - ppref);
      }
    }
    par_outlet(R, nr, o->flomas, o->sol.small, 0, 0); /* correctBoundaryConditionsVelocity: local flowo */
    FOR_RANKS {
      const fco_mesh *g = &R[r].g;
      fco_fields *f = &R[r].f;
      const int iSymmetryStart = g->numCells + g->npro + g->ninl + g->nout;
      for (int i = 1; i <= g->nsym; ++i) {
        int iface = g->iSymmetryFacesStart + i, ijp = A1(g->owner, iface), ijb = iSymmetryStart + i;
        double Unmag = A1(f->u, ijp) * A1(g->arx, iface) + A1(f->v, ijp) * A1(g->ary, iface) + A1(f->w, ijp) * A1(g->arz, iface);
        A1(f->u, ijb) = A1(f->u, ijp) - Unmag * A1(g->arx, iface);
        A1(f->v, ijb) = A1(f->v, ijp) - Unmag * A1(g->ary, iface);
        A1(f->w, ijb) = A1(f->w, ijp) - Unmag * A1(g->arz, iface);
      }
    }
    if (ipcorr != o->npcor) {
      FOR_RANKS {
        const fco_mesh *g = &R[r].g;
        fco_fields *f = &R[r].f;
        for (int i = 0; i < g->numCells; ++i) f->su[i] = 0.0;
        for (int i = 1; i <= g->numInnerFaces; ++i) {
          int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
          double fmcor;
          fco_fluxmc(g, f, ijp, ijn, A1(g->xf, i), A1(g->yf, i), A1(g->zf, i), A1(g->arx, i), A1(g->ary, i),
                     A1(g->arz, i), A1(g->facint, i), &fmcor);
          A1(f->flmass, i) = A1(f->flmass, i) + fmcor;
          A1(f->su, ijp) = A1(f->su, ijp) - fmcor;
          A1(f->su, ijn) = A1(f->su, ijn) + fmcor;
        }
        for (int i = 1; i <= g->npro; ++i) {
          int iface = g->iProcFacesStart + i, ijp = A1(g->owner, iface), ijn = g->numCells + i;
          double fmcor;
          fco_fluxmc(g, f, ijp, ijn, A1(g->xf, iface), A1(g->yf, iface), A1(g->zf, iface), A1(g->arx, iface),
                     A1(g->ary, iface), A1(g->arz, iface), A1(g->fpro, i), &fmcor);
          A1(R[r].fmpro, i) = A1(R[r].fmpro, i) + fmcor;
          A1(f->su, ijp) = A1(f->su, ijp) - fmcor;
        }
      }
    }
  }
  { /* :289-292 */
    double **v = (double **)malloc(sizeof(double *) * (size_t)nr);
    for (int c = 0; c < 4; ++c) {
      FOR_RANKS v[r] = c == 0 ? R[r].f.u : c == 1 ? R[r].f.v : c == 2 ? R[r].f.w : R[r].f.p;
      fco_par_exchange(R, nr, v, 1);
    }
    free(v);
  }
  FOR_RANKS { /* continuityErrors.h of src-parallel */
    const fco_mesh *g = &R[r].g;
    fco_fields *f = &R[r].f;
    for (int i = 0; i < g->numCells; ++i) f->res[i] = 0.0;
    for (int i = 1; i <= g->numInnerFaces; ++i) {
      int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
      A1(f->res, ijp) = A1(f->res, ijp) - A1(f->flmass, ijp);
      A1(f->res, ijn) = A1(f->res, ijn) + A1(f->flmass, ijp);
    }
    for (int i = 1; i <= g->npro; ++i) {
      int ijp = A1(g->owner, g->iProcFacesStart + i);
      A1(f->res, ijp) = A1(f->res, ijp) - A1(R[r].fmpro, i);
    }
    for (int i = 1; i <= g->ninl; ++i) {
      int ijp = A1(g->owner, g->iInletFacesStart + i);
      A1(f->res, ijp) = A1(f->res, ijp) - A1(f->fmi, i);
    }
    for (int i = 1; i <= g->nout; ++i) {
      int ijp = A1(g->owner, g->iOutletFacesStart + i);
      A1(f->res, ijp) = A1(f->res, ijp) - A1(f->fmo, i);
    }
    double sl = 0.0, gl = 0.0;
    for (int i = 1; i <= g->numCells; ++i) sl = sl + fabs(A1(f->res, i));
    for (int i = 1; i <= g->numCells; ++i) gl = gl + A1(f->res, i);
    part[r] = sl; part2[r] = gl;
  }
  rep->sumLocalContErr = gsum(part, nr);
  rep->globalContErr = gsum(part2, nr);
  free(pp); free(part); free(part2);
  return 0;
}

#include "fc_oracle_par_uvw.c"
#include "fc_oracle_par_piso.c"
