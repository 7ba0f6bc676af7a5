/*
 * fc_oracle_grad.c -- TEST INFRASTRUCTURE ONLY (see fc_oracle.h).  Included by fc_oracle.c.
 *
 * CPU restatement of the least-squares gradients and the slope limiters behind the reference's `grad`
 * dispatcher:
 *   grad_scalar_field              src/gradients.f90:95-151   (dPhidxi = 0, method, limiter)
 *   grad_lsq                       src/grad_lsq.f90            unweighted normal equations
 *   grad_lsq_dm                    src/grad_lsq_dm.f90         inverse-distance weighted normal equations
 *   grad_lsq_qr                    src/grad_lsq_qr.f90         R^-1 Q^T of the per-cell 6 x 3 system
 *   slope_limiter_* (3)            src/gradients.f90:263-522
 *
 * Parity status: UNPINNED.  grad_lsq_qr calls LAPACK DGEQRF (`-llapack`, no version pinned, absent from
 * /root/reference): its published algorithm for 3 columns (DGEQR2: DLARFG reflectors + DLARF updates) is
 * restated here; the result, R^-1 Q^T, is the unique pseudo-inverse, so a different LAPACK changes the last
 * bits only.  Everything else is in-tree arithmetic restated literally.
 *
 * Quirks kept on purpose:
 *  - grad_lsq / grad_lsq_dm return dFidxi(2) = b1*Dmat(4) - b2*Dmat(5) - b3*Dmat(6) (grad_lsq.f90:303): with
 *    the cofactors stored in Dmat this swaps the roles of b1 and b2 (the y-gradient is not the least-squares
 *    one).  Restated as written.
 *  - grad_lsq_dm's boundary weights in the solve stage use xf(i), yf(i), zf(i) with i the running index of
 *    the boundary kind, i.e. the coordinates of INNER face i, not of the boundary face (grad_lsq_dm.f90:285).
 *  - the matrix stage of grad_lsq / grad_lsq_dm visits the boundary faces in FACE order (numInnerFaces+1 ..),
 *    the solve stage in KIND order (inlet, outlet, symmetry, wall, prOutlet).
 *  - grad_lsq_qr is written for exactly m = 6 neighbours per cell (D(3,6,numCells), eye(l) assigned to a
 *    6 x 6 array): any other count is outside what the reference defines -> error here.
 *  - all three limiters compute phi_min = min(phi_max, phi(ja(k))) (gradients.f90:301): the running MAX is
 *    used on the right-hand side.
 */

/* neighbour slots of the boundary kinds in the reference's order */
static void boundary_tables(const fco_mesh *g, int cnt[5], int fst[5], int sst[5]) {
  const int c[5] = {g->ninl, g->nout, g->nsym, g->nwal, g->npru};
  const int f[5] = {g->iInletFacesStart, g->iOutletFacesStart, g->iSymmetryFacesStart, g->iWallFacesStart,
                    g->iPressOutletFacesStart};
  int slot = g->numCells + g->npro;
  for (int b = 0; b < 5; ++b) { cnt[b] = c[b]; fst[b] = f[b]; sst[b] = slot; slot += c[b]; }
}

/* ---- grad_lsq (weighted = 0) / grad_lsq_dm (weighted = 1), stage 1: dmat(9,numCells) ---- */
void fco_lsq_matrix(const fco_mesh *g, int weighted, double *dmat) {
  const int n = g->numCells;
#define DM(k, c) dmat[9 * ((size_t)(c) - 1) + (k) - 1]
  memset(dmat, 0, sizeof(double) * 9 * (size_t)n);
  for (int i = 1; i <= g->numInnerFaces; ++i) {
    int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
    double Dx = A1(g->xc, ijn) - A1(g->xc, ijp), Dy = A1(g->yc, ijn) - A1(g->yc, ijp), Dz = A1(g->zc, ijn) - A1(g->zc, ijp);
    if (weighted) {
      double w = 1.0 / (Dx * Dx + Dy * Dy + Dz * Dz);
      DM(1, ijp) = DM(1, ijp) + w * Dx * Dx; DM(1, ijn) = DM(1, ijn) + w * Dx * Dx;
      DM(4, ijp) = DM(4, ijp) + w * Dy * Dy; DM(4, ijn) = DM(4, ijn) + w * Dy * Dy;
      DM(6, ijp) = DM(6, ijp) + w * Dz * Dz; DM(6, ijn) = DM(6, ijn) + w * Dz * Dz;
      DM(2, ijp) = DM(2, ijp) + w * Dx * Dy; DM(2, ijn) = DM(2, ijn) + w * Dx * Dy;
      DM(3, ijp) = DM(3, ijp) + w * Dx * Dz; DM(3, ijn) = DM(3, ijn) + w * Dx * Dz;
      DM(5, ijp) = DM(5, ijp) + w * Dy * Dz; DM(5, ijn) = DM(5, ijn) + w * Dy * Dz;
    } else {
      DM(1, ijp) = DM(1, ijp) + Dx * Dx; DM(1, ijn) = DM(1, ijn) + Dx * Dx;
      DM(4, ijp) = DM(4, ijp) + Dy * Dy; DM(4, ijn) = DM(4, ijn) + Dy * Dy;
      DM(6, ijp) = DM(6, ijp) + Dz * Dz; DM(6, ijn) = DM(6, ijn) + Dz * Dz;
      DM(2, ijp) = DM(2, ijp) + Dx * Dy; DM(2, ijn) = DM(2, ijn) + Dx * Dy;
      DM(3, ijp) = DM(3, ijp) + Dx * Dz; DM(3, ijn) = DM(3, ijn) + Dx * Dz;
      DM(5, ijp) = DM(5, ijp) + Dy * Dz; DM(5, ijn) = DM(5, ijn) + Dy * Dz;
    }
  }
  for (int iface = g->numInnerFaces + 1; iface <= g->numFaces; ++iface) { /* face order */
    int ijp = A1(g->owner, iface);
    double Dx = A1(g->xf, iface) - A1(g->xc, ijp), Dy = A1(g->yf, iface) - A1(g->yc, ijp), Dz = A1(g->zf, iface) - A1(g->zc, ijp);
    if (weighted) {
      double w = 1.0 / (Dx * Dx + Dy * Dy + Dz * Dz);
      DM(1, ijp) = DM(1, ijp) + w * Dx * Dx; DM(4, ijp) = DM(4, ijp) + w * Dy * Dy; DM(6, ijp) = DM(6, ijp) + w * Dz * Dz;
      DM(2, ijp) = DM(2, ijp) + w * Dx * Dy; DM(3, ijp) = DM(3, ijp) + w * Dx * Dz; DM(5, ijp) = DM(5, ijp) + w * Dy * Dz;
    } else {
      DM(1, ijp) = DM(1, ijp) + Dx * Dx; DM(4, ijp) = DM(4, ijp) + Dy * Dy; DM(6, ijp) = DM(6, ijp) + Dz * Dz;
      DM(2, ijp) = DM(2, ijp) + Dx * Dy; DM(3, ijp) = DM(3, ijp) + Dx * Dz; DM(5, ijp) = DM(5, ijp) + Dy * Dz;
    }
  }
  for (int inp = 1; inp <= n; ++inp) { /* grad_lsq.f90:134-165 */
    double d11 = DM(1, inp), d12 = DM(2, inp), d13 = DM(3, inp), d22 = DM(4, inp), d23 = DM(5, inp), d33 = DM(6, inp);
    double d21 = d12, d31 = d13, d32 = d23;
    double tmp = 1.0 / (d11 * d22 * d33 - d11 * d23 * d32 - d12 * d21 * d33 + d12 * d23 * d31 + d13 * d21 * d32 - d13 * d22 * d31);
    DM(1, inp) = (d22 * d33 - d23 * d32) * tmp;
    DM(2, inp) = (d21 * d33 - d23 * d31) * tmp;
    DM(3, inp) = (d21 * d32 - d22 * d31) * tmp;
    DM(4, inp) = (d11 * d33 - d13 * d31) * tmp;
    DM(5, inp) = (d12 * d33 - d13 * d32) * tmp;
    DM(6, inp) = (d11 * d32 - d12 * d31) * tmp;
    DM(7, inp) = (d12 * d23 - d13 * d22) * tmp;
    DM(8, inp) = (d11 * d23 - d13 * d21) * tmp;
    DM(9, inp) = (d11 * d22 - d12 * d21) * tmp;
  }
}

/* stage 2 (grad_lsq.f90:168-306, grad_lsq_dm.f90:240-445) */
void fco_grad_lsq(const fco_mesh *g, int weighted, const double *dmat, const double *fi, double *dFidxi) {
  const int n = g->numCells;
  double *b = (double *)calloc(3 * (size_t)n, sizeof(double));
#define BB(k, c) b[3 * ((size_t)(c) - 1) + (k)]
  for (int i = 1; i <= g->numInnerFaces; ++i) {
    int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
    double Dx, Dy, Dz;
    if (weighted) {
      double dx = A1(g->xc, ijn) - A1(g->xc, ijp), dy = A1(g->yc, ijn) - A1(g->yc, ijp), dz = A1(g->zc, ijn) - A1(g->zc, ijp);
      double w = 1.0 / (dx * dx + dy * dy + dz * dz);
      Dx = w * (A1(g->xc, ijn) - A1(g->xc, ijp)) * (A1(fi, ijn) - A1(fi, ijp));
      Dy = w * (A1(g->yc, ijn) - A1(g->yc, ijp)) * (A1(fi, ijn) - A1(fi, ijp));
      Dz = w * (A1(g->zc, ijn) - A1(g->zc, ijp)) * (A1(fi, ijn) - A1(fi, ijp));
    } else {
      Dx = (A1(g->xc, ijn) - A1(g->xc, ijp)) * (A1(fi, ijn) - A1(fi, ijp));
      Dy = (A1(g->yc, ijn) - A1(g->yc, ijp)) * (A1(fi, ijn) - A1(fi, ijp));
      Dz = (A1(g->zc, ijn) - A1(g->zc, ijp)) * (A1(fi, ijn) - A1(fi, ijp));
    }
    BB(0, ijp) = BB(0, ijp) + Dx; BB(0, ijn) = BB(0, ijn) + Dx;
    BB(1, ijp) = BB(1, ijp) + Dy; BB(1, ijn) = BB(1, ijn) + Dy;
    BB(2, ijp) = BB(2, ijp) + Dz; BB(2, ijn) = BB(2, ijn) + Dz;
  }
  int cnt[5], fst[5], sst[5];
  boundary_tables(g, cnt, fst, sst);
  for (int k = 0; k < 5; ++k)
    for (int i = 1; i <= cnt[k]; ++i) {
      int iface = fst[k] + i, ijp = A1(g->owner, iface), ijn = sst[k] + i;
      double Dx, Dy, Dz;
      if (weighted) { /* xf(i): the reference's index (grad_lsq_dm.f90:285) */
        double ex = A1(g->xf, i) - A1(g->xc, ijp), ey = A1(g->yf, i) - A1(g->yc, ijp), ez = A1(g->zf, i) - A1(g->zc, ijp);
        double w = 1.0 / (ex * ex + ey * ey + ez * ez);
        Dx = w * (A1(fi, ijn) - A1(fi, ijp)) * (A1(g->xf, iface) - A1(g->xc, ijp));
        Dy = w * (A1(fi, ijn) - A1(fi, ijp)) * (A1(g->yf, iface) - A1(g->yc, ijp));
        Dz = w * (A1(fi, ijn) - A1(fi, ijp)) * (A1(g->zf, iface) - A1(g->zc, ijp));
      } else {
        Dx = (A1(fi, ijn) - A1(fi, ijp)) * (A1(g->xf, iface) - A1(g->xc, ijp));
        Dy = (A1(fi, ijn) - A1(fi, ijp)) * (A1(g->yf, iface) - A1(g->yc, ijp));
        Dz = (A1(fi, ijn) - A1(fi, ijp)) * (A1(g->zf, iface) - A1(g->zc, ijp));
      }
      BB(0, ijp) = BB(0, ijp) + Dx; BB(1, ijp) = BB(1, ijp) + Dy; BB(2, ijp) = BB(2, ijp) + Dz;
    }
  for (int inp = 1; inp <= n; ++inp) {
    double b1 = BB(0, inp), b2 = BB(1, inp), b3 = BB(2, inp);
    G3(dFidxi, 0, inp) = b1 * DM(1, inp) - b2 * DM(2, inp) + b3 * DM(3, inp);
    G3(dFidxi, 1, inp) = b1 * DM(4, inp) - b2 * DM(5, inp) - b3 * DM(6, inp);
    G3(dFidxi, 2, inp) = b1 * DM(7, inp) - b2 * DM(8, inp) + b3 * DM(9, inp);
  }
  free(b);
#undef BB
#undef DM
}

/* ---- grad_lsq_qr ---- */
/* DGEQR2 on an l x 3 column-major matrix with leading dimension 6 (LAPACK's unblocked Householder QR; DGEQRF
 * reduces to it for 3 columns): on exit R in the upper triangle, the reflector vectors below the diagonal. */
static double lapy2(double x, double y) {
  double xa = fabs(x), ya = fabs(y), w = xa > ya ? xa : ya, z = xa > ya ? ya : xa;
  if (z == 0.0) return w;
  return w * sqrt(1.0 + (z / w) * (z / w));
}
static void geqr2_l3(int l, double *A /* [3][6] column-major: A[j*6+i] */, double tau[3]) {
  for (int i = 0; i < 3; ++i) {
    double alpha = A[i * 6 + i], ss = 0.0;
    for (int r = i + 1; r < l; ++r) ss = ss + A[i * 6 + r] * A[i * 6 + r];
    double xnorm = sqrt(ss);
    if (xnorm == 0.0) { tau[i] = 0.0; continue; }
    double beta = -copysign(lapy2(alpha, xnorm), alpha);
    tau[i] = (beta - alpha) / beta;
    double sc = 1.0 / (alpha - beta);
    for (int r = i + 1; r < l; ++r) A[i * 6 + r] = A[i * 6 + r] * sc;
    A[i * 6 + i] = beta;
    for (int j = i + 1; j < 3; ++j) { /* DLARF: A(i:l,j) -= tau * v * (v^T A(i:l,j)), v(i) = 1 */
      double w = A[j * 6 + i];
      for (int r = i + 1; r < l; ++r) w = w + A[i * 6 + r] * A[j * 6 + r];
      A[j * 6 + i] = A[j * 6 + i] - tau[i] * w;
      for (int r = i + 1; r < l; ++r) A[j * 6 + r] = A[j * 6 + r] - tau[i] * w * A[i * 6 + r];
    }
  }
}

/* stage 1: D(3,6,numCells) = R1^-1 Q1^T (grad_lsq_qr.f90:62-247).  Returns the number of cells whose neighbour
 * count is not 6 (the routine is undefined for them); D of those cells is zero. */
int fco_lsq_qr_matrix(const fco_mesh *g, double *D) {
  const int n = g->numCells;
#define DD(i, l, c) D[18 * ((size_t)(c) - 1) + 3 * ((l) - 1) + (i) - 1]
  memset(D, 0, sizeof(double) * 18 * (size_t)n);
  int *nb = (int *)calloc((size_t)n + 1, sizeof(int));
  int bad = 0;
#define PUT(c, dx, dy, dz) do { int l_ = ++nb[c]; if (l_ <= 6) { DD(1, l_, c) = (dx); DD(2, l_, c) = (dy); DD(3, l_, c) = (dz); } } while (0)
  for (int i = 1; i <= g->numInnerFaces; ++i) {
    int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
    PUT(ijp, A1(g->xc, ijn) - A1(g->xc, ijp), A1(g->yc, ijn) - A1(g->yc, ijp), A1(g->zc, ijn) - A1(g->zc, ijp));
    PUT(ijn, A1(g->xc, ijp) - A1(g->xc, ijn), A1(g->yc, ijp) - A1(g->yc, ijn), A1(g->zc, ijp) - A1(g->zc, ijn));
  }
  for (int i = 1; i <= g->npro; ++i) { /* src-parallel/grad_lsq_qr.f90: the halo cell as neighbour, owner side only */
    int ijp = A1(g->owner, g->iProcFacesStart + i), ijn = g->numCells + i;
    PUT(ijp, A1(g->xc, ijn) - A1(g->xc, ijp), A1(g->yc, ijn) - A1(g->yc, ijp), A1(g->zc, ijn) - A1(g->zc, ijp));
  }
  int cnt[5], fst[5], sst[5];
  boundary_tables(g, cnt, fst, sst);
  for (int k = 0; k < 5; ++k)
    for (int i = 1; i <= cnt[k]; ++i) {
      int iface = fst[k] + i, ijp = A1(g->owner, iface);
      PUT(ijp, A1(g->xf, iface) - A1(g->xc, ijp), A1(g->yf, iface) - A1(g->yc, ijp), A1(g->zf, iface) - A1(g->zc, ijp));
    }
#undef PUT
  for (int inp = 1; inp <= n; ++inp) {
    const int l = nb[inp];
    if (l != 6) { ++bad; for (int k = 0; k < 18; ++k) D[18 * ((size_t)inp - 1) + k] = 0.0; continue; }
    double A[18], tau[3];
    for (int j = 0; j < 3; ++j)
      for (int r = 0; r < 6; ++r) A[j * 6 + r] = DD(j + 1, r + 1, inp); /* Dtmp = transpose(D(:,:,inp)) */
    geqr2_l3(l, A, tau);
    double r11 = A[0], r12 = A[6], r13 = A[12], r22 = A[7], r23 = A[13], r33 = A[14];
    /* H_i = I - tau_i v_i v_i^T; Q = H1 H2 H3 (only its first three columns are used) */
    double v[3][6], H[3][6][6], Q12[6][6], Q[6][6];
    for (int i = 0; i < 3; ++i)
      for (int r = 0; r < 6; ++r) v[i][r] = r < i ? 0.0 : (r == i ? 1.0 : A[i * 6 + r]);
    for (int i = 0; i < 3; ++i)
      for (int c = 0; c < 6; ++c)
        for (int r = 0; r < 6; ++r) H[i][r][c] = (r == c ? 1.0 : 0.0) + (-tau[i]) * v[i][r] * v[i][c];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s = s + H[0][r][k] * H[1][k][c];
        Q12[r][c] = s;
      }
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s = s + Q12[r][k] * H[2][k][c];
        Q[r][c] = s;
      }
    for (int k = 1; k <= 6; ++k) { /* :236-240 */
      double q1 = Q[k - 1][0], q2 = Q[k - 1][1], q3 = Q[k - 1][2];
      DD(1, k, inp) = q1 / r11 - (r12 * q2) / (r11 * r22) + (q3 * (r12 * r23 - r13 * r22)) / (r11 * r22 * r33);
      DD(2, k, inp) = q2 / r22 - (r23 * q3) / (r22 * r33);
      DD(3, k, inp) = q3 / r33;
    }
  }
  free(nb);
  return bad;
}

/* stage 2 (grad_lsq_qr.f90:250-330) */
void fco_grad_lsq_qr(const fco_mesh *g, const double *D, const double *fi, double *dFidxi) {
  const int n = g->numCells;
  double *b = (double *)calloc(6 * (size_t)n, sizeof(double));
  int *nb = (int *)calloc((size_t)n + 1, sizeof(int));
#define PUTB(c, val) do { int l_ = ++nb[c]; if (l_ <= 6) b[6 * ((size_t)(c) - 1) + l_ - 1] = (val); } while (0)
  for (int i = 1; i <= g->numInnerFaces; ++i) {
    int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
    PUTB(ijp, A1(fi, ijn) - A1(fi, ijp));
    PUTB(ijn, A1(fi, ijp) - A1(fi, ijn));
  }
  for (int i = 1; i <= g->npro; ++i) {
    int ijp = A1(g->owner, g->iProcFacesStart + i), ijn = g->numCells + i;
    PUTB(ijp, A1(fi, ijn) - A1(fi, ijp));
  }
  int cnt[5], fst[5], sst[5];
  boundary_tables(g, cnt, fst, sst);
  for (int k = 0; k < 5; ++k)
    for (int i = 1; i <= cnt[k]; ++i) {
      int ijp = A1(g->owner, fst[k] + i), ijn = sst[k] + i;
      PUTB(ijp, A1(fi, ijn) - A1(fi, ijp));
    }
#undef PUTB
  for (int inp = 1; inp <= n; ++inp) {
    int l = nb[inp] > 6 ? 6 : nb[inp];
    for (int c = 1; c <= 3; ++c) {
      double s = 0.0;
      for (int k = 1; k <= l; ++k) s = s + DD(c, k, inp) * b[6 * ((size_t)inp - 1) + k - 1];
      G3(dFidxi, c - 1, inp) = s;
    }
  }
  free(b); free(nb);
#undef DD
}

/* ---- slope limiters (gradients.f90:263-522): 1 Barth-Jespersen, 2 Venkatakrishnan, 3 mVenkatakrishnan ---- */
void fco_slope_limiter(const fco_mesh *g, const fco_csr *m, int which, const double *phi, double *dPhidxi, double small) {
  const int n = g->numCells;
  double glomin = A1(phi, 1), glomax = A1(phi, 1);
  for (int i = 2; i <= n; ++i) { if (A1(phi, i) < glomin) glomin = A1(phi, i); if (A1(phi, i) > glomax) glomax = A1(phi, i); }
  const double epsprim = 0.05;
  for (int inp = 1; inp <= n; ++inp) {
    double phi_p = A1(phi, inp);
    double phi_max = A1(phi, A1(m->ja, A1(m->ioffset, inp))), phi_min = phi_max;
    for (int k = A1(m->ioffset, inp) + 1; k <= A1(m->ioffset, inp + 1) - 1; ++k) {
      double pv = A1(phi, A1(m->ja, k));
      phi_max = FCO_MAX2(phi_max, pv);
      phi_min = FCO_MIN2(phi_max, pv); /* sic */
    }
    double deltamax = glomax - A1(phi, inp), deltamin = glomin - A1(phi, inp);
    double slopelimit = 1.0;
    for (int k = A1(m->ioffset, inp); k <= A1(m->ioffset, inp + 1) - 1; ++k) {
      if (k == A1(m->diag, inp)) continue;
      int ijn = A1(m->ja, k);
      double gradfiXdr = G3(dPhidxi, 0, inp) * (A1(g->xc, ijn) - A1(g->xc, inp)) +
                         G3(dPhidxi, 1, inp) * (A1(g->yc, ijn) - A1(g->yc, inp)) +
                         G3(dPhidxi, 2, inp) * (A1(g->zc, ijn) - A1(g->zc, inp));
      if (which == 3) {
        double cell_neighbour_value = phi_p + gradfiXdr;
        double deltam = cell_neighbour_value - phi_p, deltap;
        if (deltam > 0.0) deltap = phi_max - phi_p; else deltap = phi_min - phi_p;
        double epsi = epsprim * (glomax - glomin);
        double val = 1.0 / (deltam + small) * ((deltap * deltap + epsi * epsi) * deltam + 2 * (deltam * deltam) * deltap) /
                     (deltap * deltap + 2 * (deltam * deltam) + deltap * deltam + epsi * epsi + small);
        slopelimit = FCO_MAX2(FCO_MIN2(slopelimit, val), 0.0);
      } else {
        double r;
        if (fabs(gradfiXdr) < (double)1.e-6f) r = 1.0;
        else if (gradfiXdr > 0.0) r = deltamax / gradfiXdr;
        else r = deltamin / gradfiXdr;
        if (which == 1) slopelimit = FCO_MIN2(slopelimit, r);
        else slopelimit = FCO_MIN2(slopelimit, (r * r + 2.0 * r) / (r * r + r + 2.0));
      }
    }
    G3(dPhidxi, 0, inp) = slopelimit * G3(dPhidxi, 0, inp);
    G3(dPhidxi, 1, inp) = slopelimit * G3(dPhidxi, 1, inp);
    G3(dPhidxi, 2, inp) = slopelimit * G3(dPhidxi, 2, inp);
  }
}

/* Limiters of src-parallel/gradients.f90 on one rank: glomin / glomax arrive already reduced over the ranks
 * (global_min / global_max).  Barth-Jespersen and Venkatakrishnan are the serial loops (CSR neighbours only, the
 * unused phi_max / phi_min included); the modified Venkatakrishnan limiter of the parallel build takes the cell's
 * phimax / phimin from set_phi_min_max -- inner faces and processor faces, true min and max -- instead of the serial
 * routine's phi_min = min(phi_max, ...). */
void fco_slope_limiter_par(const fco_mesh *g, const fco_csr *m, int which, const double *phi, double *dPhidxi, double small,
                           double glomin, double glomax) {
  const int n = g->numCells;
  const double epsprim = 0.05;
  double *phimax = (double *)malloc(sizeof(double) * (size_t)(n + 1)), *phimin = (double *)malloc(sizeof(double) * (size_t)(n + 1));
  for (int i = 1; i <= n; ++i) { phimax[i] = A1(phi, i); phimin[i] = A1(phi, i); }
  for (int i = 1; i <= g->numInnerFaces; ++i) {
    int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
    phimax[ijp] = FCO_MAX2(phimax[ijp], A1(phi, ijn));
    phimax[ijn] = FCO_MAX2(phimax[ijn], A1(phi, ijp));
    phimin[ijp] = FCO_MIN2(phimin[ijp], A1(phi, ijn));
    phimin[ijn] = FCO_MIN2(phimin[ijn], A1(phi, ijp));
  }
  for (int i = 1; i <= g->npro; ++i) {
    int ijp = A1(g->owner, g->iProcFacesStart + i), ijn = n + i;
    phimax[ijp] = FCO_MAX2(phimax[ijp], A1(phi, ijn));
    phimin[ijp] = FCO_MIN2(phimin[ijp], A1(phi, ijn));
  }
  for (int inp = 1; inp <= n; ++inp) {
    double phi_p = A1(phi, inp);
    double deltamax = glomax - A1(phi, inp), deltamin = glomin - A1(phi, inp);
    double slopelimit = 1.0;
    for (int k = A1(m->ioffset, inp); k <= A1(m->ioffset, inp + 1) - 1; ++k) {
      if (k == A1(m->diag, inp)) continue;
      int ijn = A1(m->ja, k);
      double gradfiXdr = G3(dPhidxi, 0, inp) * (A1(g->xc, ijn) - A1(g->xc, inp)) +
                         G3(dPhidxi, 1, inp) * (A1(g->yc, ijn) - A1(g->yc, inp)) +
                         G3(dPhidxi, 2, inp) * (A1(g->zc, ijn) - A1(g->zc, inp));
      if (which == 3) {
        double cell_neighbour_value = phi_p + gradfiXdr;
        double deltam = cell_neighbour_value - phi_p, deltap;
        if (deltam > 0.0) deltap = phimax[inp] - phi_p; else deltap = phimin[inp] - phi_p;
        double epsi = epsprim * (glomax - glomin);
        double val = 1.0 / (deltam + small) * ((deltap * deltap + epsi * epsi) * deltam + 2 * (deltam * deltam) * deltap) /
                     (deltap * deltap + 2 * (deltam * deltam) + deltap * deltam + epsi * epsi + small);
        slopelimit = FCO_MAX2(FCO_MIN2(slopelimit, val), 0.0);
      } else {
        double r;
        if (fabs(gradfiXdr) < (double)1.e-6f) r = 1.0;
        else if (gradfiXdr > 0.0) r = deltamax / gradfiXdr;
        else r = deltamin / gradfiXdr;
        if (which == 1) slopelimit = FCO_MIN2(slopelimit, r);
        else slopelimit = FCO_MIN2(slopelimit, (r * r + 2.0 * r) / (r * r + r + 2.0));
      }
    }
    G3(dPhidxi, 0, inp) = slopelimit * G3(dPhidxi, 0, inp);
    G3(dPhidxi, 1, inp) = slopelimit * G3(dPhidxi, 1, inp);
    G3(dPhidxi, 2, inp) = slopelimit * G3(dPhidxi, 2, inp);
  }
  free(phimax); free(phimin);
}

/* ---- the dispatcher `grad(phi,dPhidxi)` (gradients.f90:95-151) with a process-wide configuration ---- */
static fco_gradient_cfg g_grad_cfg = {0, 0, 0, 0, 0.0};
void fco_set_gradient(const fco_gradient_cfg *c) {
  if (c) g_grad_cfg = *c;
  else { fco_gradient_cfg z = {0, 0, 0, 0, 0.0}; g_grad_cfg = z; }
}

void fco_limit_configured(const fco_mesh *g, const fco_csr *m, const double *phi, double *dPhidxi) {
  if (g_grad_cfg.limiter) fco_slope_limiter(g, m, g_grad_cfg.limiter, phi, dPhidxi, g_grad_cfg.small);
}

void fco_grad(const fco_mesh *g, const fco_csr *m, const double *phi, int nigrad, double *dPhidxi) {
  const fco_gradient_cfg *c = &g_grad_cfg;
  if (c->method == 0) {
    fco_grad_gauss(g, phi, nigrad, dPhidxi);
  } else {
    memset(dPhidxi, 0, sizeof(double) * 3 * (size_t)(g->numCells + g->npro));
    if (c->method == 1) fco_grad_lsq(g, 0, c->dmat, phi, dPhidxi);
    else if (c->method == 2) fco_grad_lsq_qr(g, c->dmatqr, phi, dPhidxi);
    else fco_grad_lsq(g, 1, c->dmat, phi, dPhidxi);
  }
  if (c->limiter) fco_slope_limiter(g, m, c->limiter, phi, dPhidxi, c->small);
}
