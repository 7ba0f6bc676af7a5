/*
 * fc_oracle.c -- TEST INFRASTRUCTURE ONLY (see fc_oracle.h).
 *
 * Sequential CPU restatement of freeCappuccino's pressure-correction path.
 * Build: gcc -O2 -ffp-contract=off (mirrors F90FLAGS=-Wall -O2, src/Makefile:7;
 * baseline x86-64 gfortran emits no FMA).  Every loop keeps the reference's
 * left-to-right summation order so results are bit-comparable with it.
 */
#include "fc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define A1(p, i) ((p)[(i) - 1]) /* 1-based access, Fortran style */

/* ------------------------------------------------------------------------- */
/* CSR pattern: src/sparse_matrix.f90:42-172, src/utils.f90:76-170.           */
/* COO = (i,i) ++ (owner,neigh) ++ (neigh,owner), lexicographic sort, diag =  */
/* positions with ia==ja, ioffset = first occurrence of each row, face maps   */
/* by linear search in the row (csr_to_k).  The sort result is unique, so     */
/* qsort replaces the reference's heap sort (utils.f90:424-770).              */
/* ------------------------------------------------------------------------- */
typedef struct { int i, j; } fco_pair;
static int pair_cmp(const void *a, const void *b) {
  const fco_pair *x = (const fco_pair *)a, *y = (const fco_pair *)b;
  if (x->i != y->i) return x->i < y->i ? -1 : 1;
  if (x->j != y->j) return x->j < y->j ? -1 : 1;
  return 0;
}

static int csr_to_k(int icell, int jcell, const int *ioffset, const int *ja) {
  int k = 0;
  for (int l = A1(ioffset, icell); l <= A1(ioffset, icell + 1) - 1; ++l)
    if (A1(ja, l) == jcell) { k = l; break; }
  return k;
}

int fco_create_csr(int numCells, int numInnerFaces, const int *owner, const int *neighbour,
                   int *ioffset, int *ja, int *diag, int *icell_jcell, int *jcell_icell) {
  const int nnz = numCells + 2 * numInnerFaces;
  fco_pair *coo = (fco_pair *)malloc(sizeof(fco_pair) * (size_t)nnz);
  if (!coo) return 1;
  for (int c = 1; c <= numCells; ++c) { coo[c - 1].i = c; coo[c - 1].j = c; }
  for (int i = 1; i <= numInnerFaces; ++i) {
    coo[numCells + i - 1].i = A1(owner, i);
    coo[numCells + i - 1].j = A1(neighbour, i);
  }
  for (int i = 1; i <= numInnerFaces; ++i) {
    coo[numCells + numInnerFaces + i - 1].i = A1(neighbour, i);
    coo[numCells + numInnerFaces + i - 1].j = A1(owner, i);
  }
  qsort(coo, (size_t)nnz, sizeof(fco_pair), pair_cmp);
  int id = 1;
  for (int k = 1; k <= nnz; ++k) {
    A1(ja, k) = coo[k - 1].j;
    if (coo[k - 1].i == coo[k - 1].j) { A1(diag, id) = k; ++id; }
  }
  int pos = 1;
  for (int c = 1; c <= numCells; ++c) {
    while (pos <= nnz && coo[pos - 1].i != c) ++pos; /* find_index_position: first appearance */
    A1(ioffset, c) = pos;
    ++pos;
  }
  A1(ioffset, numCells + 1) = nnz + 1;
  free(coo);
  if (icell_jcell && jcell_icell)
    for (int i = 1; i <= numInnerFaces; ++i) {
      int ijp = A1(owner, i), ijn = A1(neighbour, i);
      A1(icell_jcell, i) = csr_to_k(ijp, ijn, ioffset, ja);
      A1(jcell_icell, i) = csr_to_k(ijn, ijp, ioffset, ja);
    }
  return 0;
}

/* Host threads INSIDE one rank's row loops (bench.py's reference arm and its full-size parity solve).  Only loops
 * whose iterations are independent (one row / one cell each) are split, so every result is bit-identical to the
 * one-thread run; the inner products stay sequential left-to-right sums. */
static int fco_inner = 1;
#define ROWS_MT _Pragma("omp parallel for schedule(static) num_threads(fco_inner) if (fco_inner > 1)")

/* y = A x, row loop of src/dpcg.f90:105-110 */
void fco_spmv(const fco_csr *m, const double *a, const double *x, double *y) {
  ROWS_MT for (int i = 1; i <= m->n; ++i) {
    double s = 0.0;
    for (int k = A1(m->ioffset, i); k <= A1(m->ioffset, i + 1) - 1; ++k)
      s = s + A1(a, k) * A1(x, A1(m->ja, k));
    A1(y, i) = s;
  }
}

/* ------------------------------------------------------------------------- */
/* Shared pieces of the three solvers                                         */
/* ------------------------------------------------------------------------- */
static const fco_strips no_strips = {0, 0, 0, 0, 0, 0, 0, 0, 0};

/* res = su - A fi (+ O-C strips, + processor strips): dpcg.f90:51-61,
 * src-parallel/dpcg.f90:65-70.  Returns the L1 norm (dpcg.f90:64). */
static double initial_residual(const fco_csr *m, const double *a, const double *su, const double *fi,
                               double *res, const fco_strips *s) {
  const int n = m->n;
  ROWS_MT for (int i = 1; i <= n; ++i) {
    double r = A1(su, i);
    for (int k = A1(m->ioffset, i); k <= A1(m->ioffset, i + 1) - 1; ++k)
      r = r - A1(a, k) * A1(fi, A1(m->ja, k));
    A1(res, i) = r;
  }
  for (int i = 1; i <= s->noc; ++i) {
    A1(res, A1(s->ijl, i)) = A1(res, A1(s->ijl, i)) - A1(s->ar, i) * A1(fi, A1(s->ijr, i));
    A1(res, A1(s->ijr, i)) = A1(res, A1(s->ijr, i)) - A1(s->al, i) * A1(fi, A1(s->ijl, i));
  }
  for (int i = 1; i <= s->npro; ++i) {
    int k = A1(s->pown, i);
    A1(res, k) = A1(res, k) - A1(s->apr, i) * A1(fi, s->iProcStart + i);
  }
  double r0 = 0.0;
  for (int i = 1; i <= n; ++i) r0 = r0 + fabs(A1(res, i));
  return r0;
}

/* y = A x + strips: dpcg.f90:105-116, src-parallel/dpcg.f90:132-136.
 * `with_oc` = 0 reproduces the serial bicgstab, which applies the O-C strips
 * only to the initial residual (bicgstab.f90:142-147,190-195). */
static void matvec(const fco_csr *m, const double *a, const double *x, double *y, const fco_strips *s,
                   int with_oc) {
  fco_spmv(m, a, x, y);
  if (with_oc)
    for (int i = 1; i <= s->noc; ++i) {
      A1(y, A1(s->ijl, i)) = A1(y, A1(s->ijl, i)) + A1(s->ar, i) * A1(x, A1(s->ijr, i));
      A1(y, A1(s->ijr, i)) = A1(y, A1(s->ijr, i)) + A1(s->al, i) * A1(x, A1(s->ijl, i));
    }
  for (int i = 1; i <= s->npro; ++i) {
    int k = A1(s->pown, i);
    A1(y, k) = A1(y, k) + A1(s->apr, i) * A1(x, s->iProcStart + i);
  }
}

/* z = M^-1 r with M = (D+L) D^-1 (D+U): iccg.f90:94-111 (same in bicgstab.f90:117-136). */
static void precond_sweeps(const fco_csr *m, const double *a, const double *d, const double *r, double *zk,
                           double small) {
  const int n = m->n;
  for (int i = 1; i <= n; ++i) {
    double z = A1(r, i);
    for (int k = A1(m->ioffset, i); k <= A1(m->diag, i) - 1; ++k) z = z - A1(a, k) * A1(zk, A1(m->ja, k));
    A1(zk, i) = z * A1(d, i);
  }
  for (int i = 1; i <= n; ++i) A1(zk, i) = A1(zk, i) / (A1(d, i) + small);
  for (int i = n; i >= 1; --i) {
    double z = A1(zk, i);
    for (int k = A1(m->diag, i) + 1; k <= A1(m->ioffset, i + 1) - 1; ++k) z = z - A1(a, k) * A1(zk, A1(m->ja, k));
    A1(zk, i) = z * A1(d, i);
  }
}

/* ------------------------------------------------------------------------- */
/* dpcg: src/dpcg.f90:3-154; src-parallel/dpcg.f90 (o->parallel)              */
/* ------------------------------------------------------------------------- */
int fco_dpcg(const fco_csr *m, const double *a, const double *su, double *fi, double *res, const fco_strips *s,
             const fco_solver_opts *o, fco_report *rep, double *hist) {
  const int n = m->n;
  if (!s) s = &no_strips;
  const int npk = n + s->npro; /* pk(numPCells) in the parallel build */
  double *pk = (double *)calloc((size_t)(s->npro ? s->iProcStart + s->npro : npk), sizeof(double));
  double *zk = (double *)calloc((size_t)n, sizeof(double));
  double res0 = initial_residual(m, a, su, fi, res, s), resl = res0;
  rep->res0 = res0; rep->resl = res0; rep->iters = 0;
  if (o->tol >= 0.0 && res0 < o->tol) { free(pk); free(zk); return 0; }
  double s0 = (double)1.e20f;
  int used = 0;
  for (int l = 1; l <= o->nsw; ++l) {
    if (o->parallel) {
      ROWS_MT for (int i = 1; i <= n; ++i) A1(zk, i) = A1(res, i) / (A1(a, A1(m->diag, i)) + o->small);
    } else {
      ROWS_MT for (int i = 1; i <= n; ++i) A1(zk, i) = A1(res, i) / A1(a, A1(m->diag, i));
    }
    double sk = 0.0;
    for (int i = 1; i <= n; ++i) sk = sk + A1(res, i) * A1(zk, i);
    double bet = sk / s0;
    ROWS_MT for (int i = 1; i <= n; ++i) A1(pk, i) = A1(zk, i) + bet * A1(pk, i);
    /* single-rank "exchange": a lone rank has npro = 0; multi-rank runs go through fco_par_* */
    matvec(m, a, pk, zk, s, 1);
    double pkapk = 0.0;
    for (int i = 1; i <= n; ++i) pkapk = pkapk + A1(pk, i) * A1(zk, i);
    double alf = sk / pkapk;
    ROWS_MT for (int i = 1; i <= n; ++i) A1(fi, i) = A1(fi, i) + alf * A1(pk, i);
    ROWS_MT for (int i = 1; i <= n; ++i) A1(res, i) = A1(res, i) - alf * A1(zk, i);
    resl = 0.0;
    for (int i = 1; i <= n; ++i) resl = resl + fabs(A1(res, i));
    s0 = sk;
    ++used;
    double rsm = resl / (res0 + o->small);
    if (hist) hist[l - 1] = resl;
    if (rsm < o->sor) break;
  }
  rep->resl = resl; rep->iters = used;
  free(pk); free(zk);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* iccg: src/iccg.f90:3-177; src-parallel/iccg.f90                            */
/* ------------------------------------------------------------------------- */
int fco_iccg(const fco_csr *m, const double *a, const double *su, double *fi, double *res, const fco_strips *s,
             const fco_solver_opts *o, fco_report *rep, double *hist) {
  const int n = m->n;
  if (!s) s = &no_strips;
  double *pk = (double *)calloc((size_t)(s->npro ? s->iProcStart + s->npro : n), sizeof(double));
  double *zk = (double *)calloc((size_t)n, sizeof(double));
  double *d = (double *)calloc((size_t)n, sizeof(double));
  double res0 = initial_residual(m, a, su, fi, res, s), resl = res0;
  rep->res0 = res0; rep->resl = res0; rep->iters = 0;
  if (o->tol >= 0.0 && res0 < o->tol) { free(pk); free(zk); free(d); return 0; }
  for (int i = 1; i <= n; ++i) { /* iccg.f90:77-83 ; parallel :94-100 */
    double di = A1(a, A1(m->diag, i));
    for (int k = A1(m->ioffset, i); k <= A1(m->diag, i) - 1; ++k) {
      if (o->parallel) di = di - A1(a, k) * A1(d, A1(m->ja, k)) * A1(a, k);
      else             di = di - (A1(a, k) * A1(a, k)) * A1(d, A1(m->ja, k));
    }
    A1(d, i) = o->parallel ? 1.0 / (di + o->small) : 1.0 / di;
  }
  double s0 = (double)1.e20f;
  int used = 0;
  for (int l = 1; l <= o->nsw; ++l) {
    precond_sweeps(m, a, d, res, zk, o->small);
    double sk = 0.0;
    for (int i = 1; i <= n; ++i) sk = sk + A1(res, i) * A1(zk, i);
    double bet = sk / s0;
    for (int i = 1; i <= n; ++i) A1(pk, i) = A1(zk, i) + bet * A1(pk, i);
    matvec(m, a, pk, zk, s, 1);
    double pkapk = 0.0;
    for (int i = 1; i <= n; ++i) pkapk = pkapk + A1(pk, i) * A1(zk, i);
    double alf = sk / pkapk;
    for (int i = 1; i <= n; ++i) A1(fi, i) = A1(fi, i) + alf * A1(pk, i);
    for (int i = 1; i <= n; ++i) A1(res, i) = A1(res, i) - alf * A1(zk, i);
    resl = 0.0;
    for (int i = 1; i <= n; ++i) resl = resl + fabs(A1(res, i));
    s0 = sk;
    ++used;
    double rsm = resl / (res0 + o->small);
    if (hist) hist[l - 1] = resl;
    if (rsm < o->sor) break;
  }
  rep->resl = resl; rep->iters = used;
  free(pk); free(zk); free(d);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* bicgstab: src/bicgstab.f90:1-230; src-parallel/bicgstab.f90                */
/* ------------------------------------------------------------------------- */
int fco_bicgstab(const fco_csr *m, const double *a, const double *su, double *fi, double *res, const fco_strips *s,
                 const fco_solver_opts *o, fco_report *rep, double *hist) {
  const int n = m->n;
  if (!s) s = &no_strips;
  const size_t nz = (size_t)(s->npro ? s->iProcStart + s->npro : n);
  double *reso = (double *)calloc((size_t)n, sizeof(double));
  double *pk = (double *)calloc((size_t)n, sizeof(double));
  double *uk = (double *)calloc((size_t)n, sizeof(double));
  double *zk = (double *)calloc(nz, sizeof(double));
  double *vk = (double *)calloc((size_t)n, sizeof(double));
  double *d = (double *)calloc((size_t)n, sizeof(double));
  double res0 = initial_residual(m, a, su, fi, res, s), resl = res0;
  rep->res0 = res0; rep->resl = res0; rep->iters = 0;
  int used = 0;
  if (o->tol >= 0.0 && res0 < o->tol) goto done;
  for (int i = 1; i <= n; ++i) { /* DILU diagonal, bicgstab.f90:68-79 */
    double di = A1(a, A1(m->diag, i));
    for (int k = A1(m->ioffset, i); k <= A1(m->diag, i) - 1; ++k) {
      int jc = A1(m->ja, k), l;
      for (l = A1(m->diag, jc); l <= A1(m->ioffset, jc + 1) - 1; ++l)
        if (A1(m->ja, l) == i) break;
      di = di - A1(a, k) * A1(d, jc) * A1(a, l);
    }
    A1(d, i) = o->parallel ? 1.0 / (di + o->small) : 1.0 / di;
  }
  memcpy(reso, res, sizeof(double) * (size_t)n);
  double alf = 1.0, beto = 1.0, gam = 1.0;
  for (int l = 1; l <= o->nsw; ++l) {
    double bet = 0.0;
    for (int i = 1; i <= n; ++i) bet = bet + A1(res, i) * A1(reso, i);
    double om = bet * gam / (alf * beto + o->small);
    beto = bet;
    for (int i = 1; i <= n; ++i) A1(pk, i) = A1(res, i) + om * (A1(pk, i) - alf * A1(uk, i));
    precond_sweeps(m, a, d, pk, zk, o->small);
    matvec(m, a, zk, uk, s, 0);
    double ukreso = 0.0;
    for (int i = 1; i <= n; ++i) ukreso = ukreso + A1(uk, i) * A1(reso, i);
    gam = bet / ukreso;
    for (int i = 1; i <= n; ++i) A1(fi, i) = A1(fi, i) + gam * A1(zk, i);
    for (int i = 1; i <= n; ++i) A1(res, i) = A1(res, i) - gam * A1(uk, i);
    precond_sweeps(m, a, d, res, zk, o->small);
    matvec(m, a, zk, vk, s, 0);
    double svkres = 0.0, svkvk = 0.0;
    for (int i = 1; i <= n; ++i) svkres = svkres + A1(vk, i) * A1(res, i);
    for (int i = 1; i <= n; ++i) svkvk = svkvk + A1(vk, i) * A1(vk, i);
    alf = svkres / (svkvk + o->small);
    for (int i = 1; i <= n; ++i) A1(fi, i) = A1(fi, i) + alf * A1(zk, i);
    for (int i = 1; i <= n; ++i) A1(res, i) = A1(res, i) - alf * A1(vk, i);
    resl = 0.0;
    for (int i = 1; i <= n; ++i) resl = resl + fabs(A1(res, i));
    ++used;
    double rsm = resl / (res0 + o->small);
    if (hist) hist[l - 1] = resl;
    if (rsm < o->sor) break;
  }
done:
  rep->resl = resl; rep->iters = used;
  free(reso); free(pk); free(uk); free(zk); free(vk); free(d);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* laplacian: src/fvm_laplacian.f90:1-163, facefluxlaplacian :171-221          */
/* ------------------------------------------------------------------------- */
static void facefluxlaplacian(const fco_mesh *g, int ijp, int ijn, double arx, double ary, double arz,
                              double lambda, const double *mu, double *cap, double *can) {
  double fxn = lambda, fxp = 1.0 - lambda;
  double xpn = A1(g->xc, ijn) - A1(g->xc, ijp);
  double ypn = A1(g->yc, ijn) - A1(g->yc, ijp);
  double zpn = A1(g->zc, ijn) - A1(g->zc, ijp);
  double smdpn = (arx * arx + ary * ary + arz * arz) / (arx * xpn + ary * ypn + arz * zpn);
  *cap = (fxp * A1(mu, ijp) + fxn * A1(mu, ijn)) * smdpn;
  *can = *cap;
}

void fco_laplacian(const fco_mesh *g, const fco_csr *m, const double *mu, const double *phi, double *a,
                   double *su, double *al, double *ar) {
  for (int k = 0; k < m->nnz; ++k) a[k] = 0.0;
  for (int i = 1; i <= g->numInnerFaces; ++i) {
    int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
    double cap, can;
    facefluxlaplacian(g, ijp, ijn, A1(g->arx, i), A1(g->ary, i), A1(g->arz, i), A1(g->facint, i), mu, &cap, &can);
    A1(a, A1(m->icell_jcell, i)) = can;
    A1(a, A1(m->jcell_icell, i)) = cap;
    A1(a, A1(m->diag, ijp)) = A1(a, A1(m->diag, ijp)) - can;
    A1(a, A1(m->diag, ijn)) = A1(a, A1(m->diag, ijn)) - cap;
  }
  for (int i = 1; i <= g->noc; ++i) {
    int iface = g->iOCFacesStart + i, ijp = A1(g->ijl, i), ijn = A1(g->ijr, i);
    facefluxlaplacian(g, ijp, ijn, A1(g->arx, iface), A1(g->ary, iface), A1(g->arz, iface), A1(g->foc, i), mu,
                      &A1(al, i), &A1(ar, i));
    A1(a, A1(m->diag, ijp)) = A1(a, A1(m->diag, ijp)) - A1(ar, i);
    A1(a, A1(m->diag, ijn)) = A1(a, A1(m->diag, ijn)) - A1(al, i);
  }
  const int iWallStart = g->numCells + g->npro + g->ninl + g->nout + g->nsym;
  for (int i = 1; i <= g->nwal; ++i) { /* :140-154 */
    int iface = g->iWallFacesStart + i, ijp = A1(g->owner, iface), ijb = iWallStart + i;
    int k = A1(m->diag, ijp);
    double ax = A1(g->arx, iface), ay = A1(g->ary, iface), az = A1(g->arz, iface);
    double are = sqrt(ax * ax + ay * ay + az * az);
    double dx = A1(g->xc, ijp) - A1(g->xf, iface), dy = A1(g->yc, ijp) - A1(g->yf, iface),
           dz = A1(g->zc, ijp) - A1(g->zf, iface);
    double dpw = sqrt(dx * dx + dy * dy + dz * dz);
    A1(a, k) = A1(a, k) - A1(mu, ijp) * are / dpw;
    A1(su, ijp) = A1(su, ijp) + A1(a, k) * A1(phi, ijb);
  }
}

/* ------------------------------------------------------------------------- */
/* Gauss gradients: src/grad_gauss.f90:1-211, src/grad_gauss_corrected.f90     */
/* dudxi is the Fortran dPhidxi(3,numCells): xyz interleaved per cell.         */
/* ------------------------------------------------------------------------- */
#define G3(p, c, i) ((p)[3 * ((size_t)(i) - 1) + (c)])

static void gradco(const fco_mesh *g, int ijp, int ijn, double xfc, double yfc, double zfc, double sx, double sy,
                   double sz, double fif, const double *fi, const double *dfo, double *df) {
  double fxn = fif, fxp = 1.0 - fxn;
  double xi = A1(g->xc, ijp) * fxp + A1(g->xc, ijn) * fxn;
  double yi = A1(g->yc, ijp) * fxp + A1(g->yc, ijn) * fxn;
  double zi = A1(g->zc, ijp) * fxp + A1(g->zc, ijn) * fxn;
  double dfxi = G3(dfo, 0, ijp) * fxp + G3(dfo, 0, ijn) * fxn;
  double dfyi = G3(dfo, 1, ijp) * fxp + G3(dfo, 1, ijn) * fxn;
  double dfzi = G3(dfo, 2, ijp) * fxp + G3(dfo, 2, ijn) * fxn;
  double fie = A1(fi, ijp) * fxp + A1(fi, ijn) * fxn + dfxi * (xfc - xi) + dfyi * (yfc - yi) + dfzi * (zfc - zi);
  double dfxe = fie * sx, dfye = fie * sy, dfze = fie * sz;
  G3(df, 0, ijp) = G3(df, 0, ijp) + dfxe;
  G3(df, 1, ijp) = G3(df, 1, ijp) + dfye;
  G3(df, 2, ijp) = G3(df, 2, ijp) + dfze;
  G3(df, 0, ijn) = G3(df, 0, ijn) - dfxe;
  G3(df, 1, ijn) = G3(df, 1, ijn) - dfye;
  G3(df, 2, ijn) = G3(df, 2, ijn) - dfze;
}

static void grad_pass(const fco_mesh *g, const double *u, const double *dfo, double *df) {
  const int n = g->numCells;
  memset(df, 0, sizeof(double) * 3 * (size_t)(n + g->npro)); /* dudx(numPCells) in src-parallel */
  for (int i = 1; i <= g->numInnerFaces; ++i)
    gradco(g, A1(g->owner, i), A1(g->neighbour, i), A1(g->xf, i), A1(g->yf, i), A1(g->zf, i), A1(g->arx, i),
           A1(g->ary, i), A1(g->arz, i), A1(g->facint, i), u, dfo, df);
  for (int i = 1; i <= g->noc; ++i) {
    int iface = A1(g->ijlFace, i);
    gradco(g, A1(g->ijl, i), A1(g->ijr, i), A1(g->xf, iface), A1(g->yf, iface), A1(g->zf, iface),
           A1(g->arx, iface), A1(g->ary, iface), A1(g->arz, iface), A1(g->foc, i), u, dfo, df);
  }
  /* processor boundaries: src-parallel/grad_gauss.f90:68-75 (halo cell = iProcStart + i) */
  for (int i = 1; i <= g->npro; ++i) {
    int iface = g->iProcFacesStart + i;
    gradco(g, A1(g->owner, iface), g->numCells + i, A1(g->xf, iface), A1(g->yf, iface), A1(g->zf, iface),
           A1(g->arx, iface), A1(g->ary, iface), A1(g->arz, iface), A1(g->fpro, i), u, dfo, df);
  }
  /* boundary faces in the order inlet, outlet, symmetry, wall, prOutlet (:70-103) */
  const int cnt[5] = {g->ninl, g->nout, g->nsym, g->nwal, g->npru};
  const int fst[5] = {g->iInletFacesStart, g->iOutletFacesStart, g->iSymmetryFacesStart, g->iWallFacesStart,
                      g->iPressOutletFacesStart};
  int slot = g->numCells + g->npro;
  for (int b = 0; b < 5; ++b) {
    for (int i = 1; i <= cnt[b]; ++i) {
      int iface = fst[b] + i, ijp = A1(g->owner, iface), ijb = slot + i;
      double fi = A1(u, ijb);
      G3(df, 0, ijp) = G3(df, 0, ijp) + fi * A1(g->arx, iface);
      G3(df, 1, ijp) = G3(df, 1, ijp) + fi * A1(g->ary, iface);
      G3(df, 2, ijp) = G3(df, 2, ijp) + fi * A1(g->arz, iface);
    }
    slot += cnt[b];
  }
  for (int ijp = 1; ijp <= n; ++ijp) {
    double volr = 1.0 / A1(g->vol, ijp);
    G3(df, 0, ijp) = G3(df, 0, ijp) * volr;
    G3(df, 1, ijp) = G3(df, 1, ijp) * volr;
    G3(df, 2, ijp) = G3(df, 2, ijp) * volr;
  }
}

void fco_grad_gauss(const fco_mesh *g, const double *u, int nigrad, double *dudxi) {
  const size_t n3 = 3 * (size_t)(g->numCells + g->npro);
  double *dfo = (double *)calloc(n3, sizeof(double));
  for (int lc = 1; lc <= nigrad; ++lc) {
    grad_pass(g, u, dfo, dudxi);
    if (lc != nigrad) memcpy(dfo, dudxi, sizeof(double) * n3);
  }
  free(dfo);
}

void fco_grad_gauss_corrected(const fco_mesh *g, const double *u, double *dudxi) {
  const size_t n3 = 3 * (size_t)(g->numCells + g->npro);
  double *dfo = (double *)malloc(sizeof(double) * n3);
  memcpy(dfo, dudxi, sizeof(double) * n3);
  grad_pass(g, u, dfo, dudxi);
  free(dfo);
}

/* ------------------------------------------------------------------------- */
/* bpres: src/bpres.f90:37-150                                                 */
/* ------------------------------------------------------------------------- */
void fco_bpres(const fco_mesh *g, double *p, const double *dPdxi, int istage) {
  const int iInletStart = g->numCells + g->npro;
  const int iOutletStart = iInletStart + g->ninl, iSymmetryStart = iOutletStart + g->nout;
  const int iWallStart = iSymmetryStart + g->nsym, iPressOutletStart = iWallStart + g->nwal;
  const int iOCStart = iPressOutletStart + g->npru;
  if (istage == 1) {
    const int cnt[6] = {g->ninl, g->nout, g->nsym, g->nwal, g->npru, g->noc};
    const int fst[6] = {g->iInletFacesStart, g->iOutletFacesStart, g->iSymmetryFacesStart, g->iWallFacesStart,
                        g->iPressOutletFacesStart, g->iOCFacesStart};
    const int sst[6] = {iInletStart, iOutletStart, iSymmetryStart, iWallStart, iPressOutletStart, iOCStart};
    for (int b = 0; b < 6; ++b)
      for (int i = 1; i <= cnt[b]; ++i) A1(p, sst[b] + i) = A1(p, A1(g->owner, fst[b] + i));
  } else {
    const int cnt[2] = {g->nwal, g->npru};
    const int fst[2] = {g->iWallFacesStart, g->iPressOutletFacesStart};
    const int sst[2] = {iWallStart, iPressOutletStart};
    for (int b = 0; b < 2; ++b)
      for (int i = 1; i <= cnt[b]; ++i) {
        int iface = fst[b] + i, ijp = A1(g->owner, iface), ijb = sst[b] + i;
        double xpb = A1(g->xf, iface) - A1(g->xc, ijp);
        double ypb = A1(g->yf, iface) - A1(g->yc, ijp);
        double zpb = A1(g->zf, iface) - A1(g->zc, ijp);
        A1(p, ijb) = A1(p, ijp) + G3(dPdxi, 0, ijp) * xpb + G3(dPdxi, 1, ijp) * ypb + G3(dPdxi, 2, ijp) * zpb;
      }
    for (int i = 1; i <= g->noc; ++i) { /* :134-148 -- note xf(i), not xf(iface): kept as is */
      int iface = g->iOCFacesStart + i, ijp = A1(g->owner, iface), ijb = iOCStart + i;
      double xpb = A1(g->xf, i) - A1(g->xc, ijp);
      double ypb = A1(g->yf, i) - A1(g->yc, ijp);
      double zpb = A1(g->zf, i) - A1(g->zc, ijp);
      A1(p, ijb) = A1(p, ijp) + G3(dPdxi, 0, ijp) * xpb + G3(dPdxi, 1, ijp) * ypb + G3(dPdxi, 2, ijp) * zpb;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* face_value_central: src/interpolation.f90:145-199                          */
/* ------------------------------------------------------------------------- */
static double face_value_central(const fco_mesh *g, int inp, int inn, double xf, double yf, double zf,
                                 const double *fi, const double *gradfi) {
  double gradfidr = G3(gradfi, 0, inp) * (xf - A1(g->xc, inp)) + G3(gradfi, 1, inp) * (yf - A1(g->yc, inp)) +
                    G3(gradfi, 2, inp) * (zf - A1(g->zc, inp)) + G3(gradfi, 0, inn) * (xf - A1(g->xc, inn)) +
                    G3(gradfi, 1, inn) * (yf - A1(g->yc, inn)) + G3(gradfi, 2, inn) * (zf - A1(g->zc, inn));
  return 0.5 * (A1(fi, inp) + A1(fi, inn) + gradfidr);
}

/* ------------------------------------------------------------------------- */
/* facefluxmass :34-199 (variant 0), facefluxmass2 :204-288 (1),               */
/* facefluxmass_piso :294-358 (2) of src/facefluxmass.f90                     */
/* ------------------------------------------------------------------------- */
void fco_facefluxmass(const fco_mesh *g, const fco_fields *f, int variant, int ijp, int ijn, double xf, double yf,
                      double zf, double arx, double ary, double arz, double lambda, double *cap, double *can,
                      double *fluxmass) {
  double fxn = lambda, fxp = 1.0 - lambda;
  double xpn = A1(g->xc, ijn) - A1(g->xc, ijp);
  double ypn = A1(g->yc, ijn) - A1(g->yc, ijp);
  double zpn = A1(g->zc, ijn) - A1(g->zc, ijp);
  double dpn = sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
  double are = sqrt(arx * arx + ary * ary + arz * arz);
  double dene = A1(f->den, ijp) * fxp + A1(f->den, ijn) * fxn;
  double ui = face_value_central(g, ijp, ijn, xf, yf, zf, f->u, f->dUdxi);
  double vi = face_value_central(g, ijp, ijn, xf, yf, zf, f->v, f->dVdxi);
  double wi = face_value_central(g, ijp, ijn, xf, yf, zf, f->w, f->dWdxi);
  const double *dP = f->dPdxi;
  if (variant == 1) {
    double Kj = A1(g->vol, ijp) * A1(f->apu, ijp) * fxp + A1(g->vol, ijn) * A1(f->apu, ijn) * fxn;
    *cap = -dene * Kj * are / dpn;
    *can = *cap;
    double dpxi = (G3(dP, 0, ijn) * fxp + G3(dP, 0, ijp) * fxn) * xpn;
    double dpyi = (G3(dP, 1, ijn) * fxp + G3(dP, 1, ijp) * fxn) * ypn;
    double dpzi = (G3(dP, 2, ijn) * fxp + G3(dP, 2, ijp) * fxn) * zpn;
    *fluxmass = dene * (ui * arx + vi * ary + wi * arz) +
                *cap * (A1(f->p, ijn) - A1(f->p, ijp) - dpxi - dpyi - dpzi);
    return;
  }
  if (variant == 2) {
    *cap = -dene * (fxp * A1(g->vol, ijp) * A1(f->apu, ijp) + fxn * A1(g->vol, ijn) * A1(f->apu, ijn)) * are / dpn;
    *can = *cap;
    *fluxmass = dene * (ui * arx + vi * ary + wi * arz);
    return;
  }
  double nxx = arx / are, nyy = ary / are, nzz = arz / are;
  double Dpu = (fxn * A1(g->vol, ijn) * A1(f->apu, ijn) + fxp * A1(g->vol, ijp) * A1(f->apu, ijp));
  double Dpv = (fxn * A1(g->vol, ijn) * A1(f->apv, ijn) + fxp * A1(g->vol, ijp) * A1(f->apv, ijp));
  double Dpw = (fxn * A1(g->vol, ijn) * A1(f->apw, ijn) + fxp * A1(g->vol, ijp) * A1(f->apw, ijp));
  double sfdpnr = 1.0 / (arx * xpn * nxx + ary * ypn * nyy + arz * zpn * nzz);
  double smdpn = (arx * arx + ary * ary + arz * arz) * sfdpnr;
  *cap = -dene * Dpu * smdpn;
  *can = *cap;
  double dpxi = Dpu * (fxn * G3(dP, 0, ijn) + fxp * G3(dP, 0, ijp)) * xpn * nxx;
  double dpyi = Dpv * (fxn * G3(dP, 1, ijn) + fxp * G3(dP, 1, ijp)) * ypn * nyy;
  double dpzi = Dpw * (fxn * G3(dP, 2, ijn) + fxp * G3(dP, 2, ijp)) * zpn * nzz;
  double xpp = xf - (xf - A1(g->xc, ijp)) * nxx;
  double ypp = yf - (yf - A1(g->yc, ijp)) * nyy;
  double zpp = zf - (zf - A1(g->zc, ijp)) * nzz;
  double xep = xf - (xf - A1(g->xc, ijn)) * nxx;
  double yep = yf - (yf - A1(g->yc, ijn)) * nyy;
  double zep = zf - (zf - A1(g->zc, ijn)) * nzz;
  xpp = xpp - A1(g->xc, ijp); ypp = ypp - A1(g->yc, ijp); zpp = zpp - A1(g->zc, ijp);
  xep = xep - A1(g->xc, ijn); yep = yep - A1(g->yc, ijn); zep = zep - A1(g->zc, ijn);
  double dpe = (A1(f->p, ijn) - A1(f->p, ijp));
  /* operator precedence of :170-171 kept: only the first ijp term is subtracted */
  double dpecorr = (G3(dP, 0, ijn) * xep + G3(dP, 1, ijn) * yep + G3(dP, 2, ijn) * zep - G3(dP, 0, ijp) * xpp +
                    G3(dP, 1, ijp) * ypp + G3(dP, 2, ijp) * zpp);
  dpe = dpe + dpecorr;
  double dpex = Dpu * dpe * sfdpnr * arx;
  double dpey = Dpv * dpe * sfdpnr * ary;
  double dpez = Dpw * dpe * sfdpnr * arz;
  double ue = ui - dpex + dpxi;
  double ve = vi - dpey + dpyi;
  double we = wi - dpez + dpzi;
  *fluxmass = dene * (ue * arx + ve * ary + we * arz);
}

/* fluxmc: src/facefluxmass.f90:520-607 */
void fco_fluxmc(const fco_mesh *g, const fco_fields *f, int ijp, int ijn, double xf, double yf, double zf,
                double arx, double ary, double arz, double lambda, double *fmcor) {
  double fxn = lambda, fxp = 1.0 - lambda;
  double xpn = A1(g->xc, ijn) - A1(g->xc, ijp);
  double ypn = A1(g->yc, ijn) - A1(g->yc, ijp);
  double zpn = A1(g->zc, ijn) - A1(g->zc, ijp);
  double are = sqrt(arx * arx + ary * ary + arz * arz);
  double nxx = arx / are, nyy = ary / are, nzz = arz / are;
  double dppnnr = 1.0 / ((xpn * nxx) + (ypn * nyy) + (zpn * nzz));
  double xpp = xf - (xf - A1(g->xc, ijp)) * nxx;
  double ypp = yf - (yf - A1(g->yc, ijp)) * nyy;
  double zpp = zf - (zf - A1(g->zc, ijp)) * nzz;
  double xep = xf - (xf - A1(g->xc, ijn)) * nxx;
  double yep = yf - (yf - A1(g->yc, ijn)) * nyy;
  double zep = zf - (zf - A1(g->zc, ijn)) * nzz;
  xpp = xpp - A1(g->xc, ijp); ypp = ypp - A1(g->yc, ijp); zpp = zpp - A1(g->zc, ijp);
  xep = xep - A1(g->xc, ijn); yep = yep - A1(g->yc, ijn); zep = zep - A1(g->zc, ijn);
  double rapr = (A1(f->apu, ijp) * A1(f->den, ijp) * A1(g->vol, ijp) * fxp +
                 A1(f->apu, ijn) * A1(f->den, ijn) * A1(g->vol, ijn) * fxn);
  const double *dP = f->dPdxi;
  *fmcor = rapr * are *
           ((G3(dP, 0, ijn) * xep - G3(dP, 0, ijp) * xpp) + (G3(dP, 1, ijn) * yep - G3(dP, 1, ijp) * ypp) +
            (G3(dP, 2, ijn) * zep - G3(dP, 2, ijp) * zpp)) *
           dppnnr;
}

/* ------------------------------------------------------------------------- */
/* adjustMassFlow (src/adjustMassFlow.f90), correctBoundaryConditionsVelocity  */
/* ------------------------------------------------------------------------- */
static void outlet_extrapolate_and_scale(const fco_mesh *g, fco_fields *f, double flomas, double small,
                                         int add_to_su) {
  const int iOutletStart = g->numCells + g->npro + g->ninl;
  double flowo = 0.0;
  for (int i = 1; i <= g->nout; ++i) {
    int iface = g->iOutletFacesStart + i, ijp = A1(g->owner, iface), ijb = iOutletStart + i;
    A1(f->u, ijb) = A1(f->u, ijp);
    A1(f->v, ijb) = A1(f->v, ijp);
    A1(f->w, ijb) = A1(f->w, ijp);
    A1(f->fmo, i) = A1(f->den, ijp) * (A1(f->u, ijb) * A1(g->arx, iface) + A1(f->v, ijb) * A1(g->ary, iface) +
                                        A1(f->w, ijb) * A1(g->arz, iface));
    flowo = flowo + A1(f->fmo, i);
  }
  double fac = flomas / (flowo + small);
  for (int i = 1; i <= g->nout; ++i) {
    int iface = g->iOutletFacesStart + i, ijp = A1(g->owner, iface), ijb = iOutletStart + i;
    A1(f->fmo, i) = A1(f->fmo, i) * fac;
    A1(f->u, ijb) = A1(f->u, ijb) * fac;
    A1(f->v, ijb) = A1(f->v, ijb) * fac;
    A1(f->w, ijb) = A1(f->w, ijb) * fac;
    if (add_to_su) A1(f->su, ijp) = A1(f->su, ijp) - A1(f->fmo, i);
  }
}

static void adjustMassFlow(const fco_mesh *g, fco_fields *f, double flomas, double small) {
  for (int i = 1; i <= g->ninl; ++i) {
    int ijp = A1(g->owner, g->iInletFacesStart + i);
    A1(f->su, ijp) = A1(f->su, ijp) - A1(f->fmi, i);
  }
  outlet_extrapolate_and_scale(g, f, flomas, small, 1);
}

static void correctBoundaryConditionsVelocity(const fco_mesh *g, fco_fields *f, double flomas, double small) {
  outlet_extrapolate_and_scale(g, f, flomas, small, 0);
  const int iSymmetryStart = g->numCells + g->npro + g->ninl + g->nout;
  for (int i = 1; i <= g->nsym; ++i) {
    int iface = g->iSymmetryFacesStart + i, ijp = A1(g->owner, iface), ijb = iSymmetryStart + i;
    double Unmag = A1(f->u, ijp) * A1(g->arx, iface) + A1(f->v, ijp) * A1(g->ary, iface) +
                   A1(f->w, ijp) * A1(g->arz, iface);
    A1(f->u, ijb) = A1(f->u, ijp) - Unmag * A1(g->arx, iface);
    A1(f->v, ijb) = A1(f->v, ijp) - Unmag * A1(g->ary, iface);
    A1(f->w, ijb) = A1(f->w, ijp) - Unmag * A1(g->arz, iface);
  }
}

/* continuityErrors.h -- including its flmass(ijp) (not flmass(i)) indexing, :18-19 */
static void continuity_errors(const fco_mesh *g, fco_fields *f, double *sumLocal, double *global) {
  const int n = g->numCells;
  for (int i = 0; i < n; ++i) f->res[i] = 0.0;
  for (int i = 1; i <= g->numInnerFaces; ++i) {
    int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
    A1(f->res, ijp) = A1(f->res, ijp) - A1(f->flmass, ijp);
    A1(f->res, ijn) = A1(f->res, ijn) + A1(f->flmass, ijp);
  }
  for (int i = 1; i <= g->noc; ++i) {
    A1(f->res, A1(g->ijl, i)) = A1(f->res, A1(g->ijl, i)) - A1(f->fmoc, i);
    A1(f->res, A1(g->ijr, i)) = A1(f->res, A1(g->ijr, i)) + A1(f->fmoc, i);
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
  for (int i = 1; i <= n; ++i) sl = sl + fabs(A1(f->res, i));
  for (int i = 1; i <= n; ++i) gl = gl + A1(f->res, i);
  *sumLocal = sl;
  *global = gl;
}

/* ------------------------------------------------------------------------- */
/* calcp: src/calcp-multiple_correction_SIMPLE.f90                            */
/* ------------------------------------------------------------------------- */
void fco_calcp_assemble(const fco_mesh *g, const fco_csr *m, fco_fields *f, const fco_calcp_opts *o) {
  const int n = g->numCells;
  for (int k = 0; k < m->nnz; ++k) f->a[k] = 0.0;  /* :34-35 */
  for (int i = 0; i < n; ++i) f->su[i] = 0.0;
  /* grad(U), grad(V), grad(W) :38-40; grad_scalar_field zeroes, then grad_gauss (gradients.f90:104,128) */
  fco_grad(g, m, f->u, o->nigrad, f->dUdxi);
  fco_grad(g, m, f->v, o->nigrad, f->dVdxi);
  fco_grad(g, m, f->w, o->nigrad, f->dWdxi);
  for (int i = 1; i <= g->numInnerFaces; ++i) { /* :45-77 */
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
  for (int i = 1; i <= g->noc; ++i) { /* :81-104 */
    int iface = A1(g->ijlFace, i), ijp = A1(g->ijl, i), ijn = A1(g->ijr, i);
    fco_facefluxmass(g, f, o->flux_variant, ijp, ijn, A1(g->xf, iface), A1(g->yf, iface), A1(g->zf, iface),
                     A1(g->arx, iface), A1(g->ary, iface), A1(g->arz, iface), A1(g->foc, i), &A1(f->al, i),
                     &A1(f->ar, i), &A1(f->fmoc, i));
    A1(f->a, A1(m->diag, ijp)) = A1(f->a, A1(m->diag, ijp)) - A1(f->ar, i);
    A1(f->a, A1(m->diag, ijn)) = A1(f->a, A1(m->diag, ijn)) - A1(f->al, i);
    A1(f->su, ijp) = A1(f->su, ijp) - A1(f->fmoc, i);
    A1(f->su, ijn) = A1(f->su, ijn) + A1(f->fmoc, i);
  }
  if (!o->const_mflux) adjustMassFlow(g, f, o->flomas, o->sol.small); /* :107 */
}

int fco_calcp(const fco_mesh *g, const fco_csr *m, fco_fields *f, const fco_calcp_opts *o, fco_calcp_report *rep) {
  const int n = g->numCells;
  fco_calcp_assemble(g, m, f, o);
  fco_strips st = {g->noc, g->ijl, g->ijr, f->al, f->ar, 0, 0, 0, 0};
  for (int ipcorr = 1; ipcorr <= o->npcor; ++ipcorr) { /* :112 */
    for (int i = 0; i < g->numTotal; ++i) f->pp[i] = 0.0;
    fco_report *r = &rep->rep[ipcorr - 1 < 8 ? ipcorr - 1 : 7];
    if (o->solver == 0) fco_dpcg(m, f->a, f->su, f->pp, f->res, &st, &o->sol, r, 0);
    else if (o->solver == 1) fco_iccg(m, f->a, f->su, f->pp, f->res, &st, &o->sol, r, 0);
    else fco_bicgstab(m, f->a, f->su, f->pp, f->res, &st, &o->sol, r, 0);
    for (int istage = 1; istage <= o->nipgrad; ++istage) { /* :132-140 */
      fco_bpres(g, f->pp, f->dPdxi, istage);
      fco_grad(g, m, f->pp, o->nigrad, f->dPdxi);
    }
    if (o->lsq_flag) { /* :143 -- the option wrapper zeroes dPdxi first (gradients.f90:222) */
      memset(f->dPdxi, 0, sizeof(double) * 3 * (size_t)n);
      fco_grad_gauss_corrected(g, f->pp, f->dPdxi);
      fco_limit_configured(g, m, f->pp, f->dPdxi); /* grad_scalar_field_w_option ends with the limiter (:240-255) */
    }
    double ppref = A1(f->pp, o->pRefCell); /* :146 */
    for (int iface = 1; iface <= g->numInnerFaces; ++iface) { /* :154-164 */
      int ijp = A1(g->owner, iface), ijn = A1(g->neighbour, iface);
      int k = A1(m->icell_jcell, iface);
      A1(f->flmass, iface) = A1(f->flmass, iface) + A1(f->a, k) * (A1(f->pp, ijn) - A1(f->pp, ijp));
    }
    for (int i = 1; i <= g->noc; ++i)
      A1(f->fmoc, i) = A1(f->fmoc, i) + A1(f->ar, i) * (A1(f->pp, A1(g->ijr, i)) - A1(f->pp, A1(g->ijl, i)));
    for (int inp = 1; inp <= n; ++inp) { /* :176-181 */
      A1(f->u, inp) = A1(f->u, inp) - G3(f->dPdxi, 0, inp) * A1(g->vol, inp) * A1(f->apu, inp);
      A1(f->v, inp) = A1(f->v, inp) - G3(f->dPdxi, 1, inp) * A1(g->vol, inp) * A1(f->apv, inp);
      A1(f->w, inp) = A1(f->w, inp) - G3(f->dPdxi, 2, inp) * A1(g->vol, inp) * A1(f->apw, inp);
      A1(f->p, inp) = A1(f->p, inp) + o->urf_p * (A1(f->pp, inp) - ppref);
    }
    correctBoundaryConditionsVelocity(g, f, o->flomas, o->sol.small); /* :184 */
    if (ipcorr != o->npcor) { /* :187-223 */
      for (int i = 0; i < n; ++i) f->su[i] = 0.0;
      for (int i = 1; i <= g->numInnerFaces; ++i) {
        int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
        double fmcor;
        fco_fluxmc(g, f, ijp, ijn, A1(g->xf, i), A1(g->yf, i), A1(g->zf, i), A1(g->arx, i), A1(g->ary, i),
                   A1(g->arz, i), A1(g->facint, i), &fmcor);
        A1(f->flmass, i) = A1(f->flmass, i) + fmcor;
        A1(f->su, ijp) = A1(f->su, ijp) - fmcor;
        A1(f->su, ijn) = A1(f->su, ijn) + fmcor;
      }
      for (int i = 1; i <= g->noc; ++i) {
        int iface = A1(g->ijlFace, i), ijp = A1(g->ijl, i), ijn = A1(g->ijr, i);
        double fmcor;
        fco_fluxmc(g, f, ijp, ijn, A1(g->xf, iface), A1(g->yf, iface), A1(g->zf, iface), A1(g->arx, iface),
                   A1(g->ary, iface), A1(g->arz, iface), A1(g->foc, i), &fmcor);
        A1(f->fmoc, i) = A1(f->fmoc, i) + fmcor;
        A1(f->su, ijp) = A1(f->su, ijp) - fmcor;
        A1(f->su, ijn) = A1(f->su, ijn) + fmcor;
      }
    }
  }
  double sl, gl;
  continuity_errors(g, f, &sl, &gl);
  rep->sumLocalContErr = sl;
  rep->globalContErr = gl;
  return 0;
}

#include "fc_oracle_uvw.c"
#include "fc_oracle_grad.c"
#include "fc_oracle_piso.c"
