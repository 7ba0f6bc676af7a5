// TEST INFRASTRUCTURE ONLY -- not part of the product, never loaded by freecappuccino_b200.
//
// Compiles the per-index bodies of the momentum-predictor kernels
// (freecappuccino_b200/csrc/fc_momentum_body.cuh) with g++ and runs them in plain loops, so that
// tests/test_momentum_bodies.py can check their index logic and arithmetic order against the oracle
// on a machine without a GPU.  On the GPU the same bodies are called by the thin __global__ wrappers
// of fc_momentum.cu (tests/test_gpu_zz1_momentum.py compares those with the oracle).  The library has
// no host path: this file is built only by the test that uses it.
#include "../../freecappuccino_b200/csrc/fc_piso_body.cuh"   // includes fc_momentum_body.cuh
#include "../../freecappuccino_b200/csrc/fc_grad_body.cuh"

extern "C" {

void fcm_host_assemble(const fcm_geom *g, const fcm_c2f *m, const fcm_slots *sl, const fcm_flow *f, const fcm_opts *o,
                       const fcm_faces *fa, const fcm_proc *P, const fcm_rows *r, int nnz) {
  for (int i = 0; i < g->F; ++i) fcm_face(*g, *f, *o, *fa, i);
  for (int i = 0; i < P->npro; ++i) fcm_proc_face(*g, *f, *o, *P, i);
  for (int c = 0; c < g->n; ++c) fcm_row(*g, *m, *sl, *f, *o, *fa, *P, *r, c);
  if (o->cn) {
    for (int k = 0; k < nnz; ++k) r->a[k] = 0.5 * r->a[k];
    for (int i = 0; i < P->npro; ++i) P->apr[i] = 0.5 * P->apr[i];
  }
}

void fcm_host_component(const fcm_geom *g, const fcm_c2f *m, const fcm_comp *k) {
  for (int c = 0; c < g->n; ++c) fcm_component(*g, *m, *k, c);
}

// get_rAU_x_UEqnH: all rows first (they read the old u, v, w of their neighbours), then u = apu * su ...
void fcp_host_hbya(const fcm_geom *g, const fcm_c2f *m, const fcp_hbya *k, const double *apu, const double *apv,
                   const double *apw, double *u, double *v, double *w) {
  for (int c = 0; c < g->n; ++c) fcp_hbya_row(*g, *m, *k, c);
  for (int c = 0; c < g->n; ++c) fcp_hbya_scale(c, apu, apv, apw, k->su, k->sv, k->sw, u, v, w);
}

// the small kernels of the PISO driver: pin the reference row, flux correction from the matrix, velocity
// correction, PIMPLE pressure relaxation
void fcp_host_tail(const fcm_geom *g, const int *ioffset, const int *diag, const int *icj, double *a, double *su,
                   const double *src, int pref, const double *pp, double *flmass, const double *apu, const double *apv,
                   const double *apw, const double *dP, double *u, double *v, double *w, double urf, double *p) {
  fcp_pin_row(ioffset, diag, a, su, src, pref);
  for (int i = 0; i < g->F; ++i) fcp_flux_correct(*g, icj, a, pp, flmass, i);
  for (int c = 0; c < g->n; ++c) fcp_velocity_correct(*g, apu, apv, apw, dP, u, v, w, c);
  for (int c = 0; c < g->n; ++c) fcp_relax_p(urf, pp, p, c);
}

// least-squares gradients and limiters (fc_grad_body.cuh)
void fcg_host_lsq(const fcm_geom *g, const fcm_c2f *m, const fcm_slots *sl, int weighted, double *dmat, const double *fi,
                  double *out) {
  for (int c = 0; c < g->n; ++c) fcg_lsq_matrix_row(*g, *m, weighted, dmat, c);
  for (int c = 0; c < g->n; ++c) fcg_grad_lsq_row(*g, *m, *sl, weighted, dmat, fi, out, c);
}
int fcg_host_lsq_qr(const fcm_geom *g, const fcm_c2f *m, double *D, const double *fi, double *out) {
  int bad = 0;
  for (int c = 0; c < g->n; ++c) bad += fcg_lsq_qr_matrix_row(*g, *m, D, c);
  for (int c = 0; c < g->n; ++c) fcg_grad_lsq_qr_row(*g, *m, D, fi, out, c);
  return bad;
}
void fcg_host_limiter(const fcm_geom *g, const int *ioffset, const int *ja, const int *diag, int which,
                      const double *phi, double *grad, double glomin, double glomax, double small) {
  for (int c = 0; c < g->n; ++c) fcg_limiter_row(*g, ioffset, ja, diag, which, phi, grad, glomin, glomax, small, c);
}

// nf = 3 or 4 Gauss gradients in one walk (fcg_gaussn_row): nigrad passes, the later ones seeded with the previous result
void fcg_host_gaussn(const fcm_geom *g, const fcm_c2f *m, int npro, const double *fpro, int nf, const double **phi,
                     double **out, double **old, int nigrad) {
  fcg_gaussn k{};
  k.npro = npro; k.fpro = fpro;
  for (int t = 0; t < nf; ++t) { k.phi[t] = phi[t]; k.old[t] = old[t]; k.out[t] = out[t]; }
  for (int lc = 1; lc <= nigrad; ++lc) {
    if (lc > 1)
      for (int t = 0; t < nf; ++t)
        for (size_t i = 0; i < 3 * (size_t)g->n; ++i) old[t][i] = out[t][i];
    for (int c = 0; c < g->n; ++c) {
      if (nf == 3) { if (lc == 1) fcg_gaussn_row<3, false>(*g, *m, k, c); else fcg_gaussn_row<3, true>(*g, *m, k, c); }
      else         { if (lc == 1) fcg_gaussn_row<4, false>(*g, *m, k, c); else fcg_gaussn_row<4, true>(*g, *m, k, c); }
    }
  }
}

int fcm_host_sizes(int which) {
  switch (which) {
    case 0: return (int)sizeof(fcm_geom);
    case 1: return (int)sizeof(fcm_c2f);
    case 2: return (int)sizeof(fcm_slots);
    case 3: return (int)sizeof(fcm_flow);
    case 4: return (int)sizeof(fcm_opts);
    case 5: return (int)sizeof(fcm_faces);
    case 6: return (int)sizeof(fcm_rows);
    case 7: return (int)sizeof(fcm_comp);
    case 8: return (int)sizeof(fcp_hbya);
    case 9: return (int)sizeof(fcm_proc);
  }
  return -1;
}
}
