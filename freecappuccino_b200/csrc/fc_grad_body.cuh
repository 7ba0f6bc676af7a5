// Least-squares gradients and slope limiters behind the reference's `grad` dispatcher (src/gradients.f90:95-151,
// src/grad_lsq.f90, src/grad_lsq_dm.f90, src/grad_lsq_qr.f90, limiters src/gradients.f90:263-522; SURVEY.md 8(f)
// rank 3): per-cell bodies of the kernels in fc_gradients.cu.  One thread per cell walks the cell-to-face map, so
// the reference's face-loop scatters become ordered gathers (same sums, no atomics).
//
// Reference quirks kept (DESIGN.md lists them): dFidxi(2) of grad_lsq / grad_lsq_dm as
// written (grad_lsq.f90:303), grad_lsq_dm's boundary weights read xf(i) of the running index (grad_lsq_dm.f90:285),
// matrix stage in face order vs solve stage in kind order, grad_lsq_qr defined for exactly six neighbours,
// phi_min = min(phi_max, ...) in all three limiters (gradients.f90:301).
// grad_lsq_qr's LAPACK DGEQRF is restated as DGEQR2 (Householder reflectors); the result R^-1 Q^T is the
// pseudo-inverse, independent of the QR algorithm up to rounding.
#pragma once
#include "fc_momentum_body.cuh"

#define FCG_DM(k, c) dmat[9 * (size_t)(c) + (k) - 1]

// ---- grad_lsq (weighted = 0) / grad_lsq_dm (weighted = 1), matrix stage: dmat(9,numCells) ----
FCM_HD void fcg_lsq_acc(double &d1, double &d2, double &d3, double &d4, double &d5, double &d6, int weighted, double Dx,
                        double Dy, double Dz) {
  if (weighted) {
    const double w = 1.0 / (Dx * Dx + Dy * Dy + Dz * Dz);
    d1 = d1 + w * Dx * Dx; d4 = d4 + w * Dy * Dy; d6 = d6 + w * Dz * Dz;
    d2 = d2 + w * Dx * Dy; d3 = d3 + w * Dx * Dz; d5 = d5 + w * Dy * Dz;
  } else {
    d1 = d1 + Dx * Dx; d4 = d4 + Dy * Dy; d6 = d6 + Dz * Dz;
    d2 = d2 + Dx * Dy; d3 = d3 + Dx * Dz; d5 = d5 + Dy * Dz;
  }
}

FCM_HD void fcg_lsq_matrix_row(const fcm_geom &g, const fcm_c2f &m, int weighted, double *dmat, int c) {
  double d1 = 0.0, d2 = 0.0, d3 = 0.0, d4 = 0.0, d5 = 0.0, d6 = 0.0;
  const int qs = m.off[c], qe = m.off[c + 1];
  for (int q = qs; q < qe; ++q) {
    const int fe = m.face[q];
    if ((fe & 0x7fffffff) >= g.F) continue;
    const int ijp = fe < 0 ? m.other[q] : c, ijn = fe < 0 ? c : m.other[q];
    fcg_lsq_acc(d1, d2, d3, d4, d5, d6, weighted, g.xc[ijn] - g.xc[ijp], g.yc[ijn] - g.yc[ijp], g.zc[ijn] - g.zc[ijp]);
  }
  // boundary faces in FACE order (grad_lsq.f90:116-131 loops over numBoundaryFaces), not the map's kind order
  int last = -1;
  for (;;) {
    int best = 0x7fffffff;
    for (int q = qs; q < qe; ++q) {
      const int fc = m.face[q] & 0x7fffffff;
      if (fc >= g.F && fc > last && fc < best) best = fc;
    }
    if (best == 0x7fffffff) break;
    last = best;
    fcg_lsq_acc(d1, d2, d3, d4, d5, d6, weighted, g.xf[best] - g.xc[c], g.yf[best] - g.yc[c], g.zf[best] - g.zc[c]);
  }
  const double d11 = d1, d12 = d2, d13 = d3, d22 = d4, d23 = d5, d33 = d6;
  const double d21 = d12, d31 = d13, d32 = d23;
  const double tmp = 1.0 / (d11 * d22 * d33 - d11 * d23 * d32 - d12 * d21 * d33 + d12 * d23 * d31 + d13 * d21 * d32 -
                            d13 * d22 * d31);
  FCG_DM(1, c) = (d22 * d33 - d23 * d32) * tmp;
  FCG_DM(2, c) = (d21 * d33 - d23 * d31) * tmp;
  FCG_DM(3, c) = (d21 * d32 - d22 * d31) * tmp;
  FCG_DM(4, c) = (d11 * d33 - d13 * d31) * tmp;
  FCG_DM(5, c) = (d12 * d33 - d13 * d32) * tmp;
  FCG_DM(6, c) = (d11 * d32 - d12 * d31) * tmp;
  FCG_DM(7, c) = (d12 * d23 - d13 * d22) * tmp;
  FCG_DM(8, c) = (d11 * d23 - d13 * d21) * tmp;
  FCG_DM(9, c) = (d11 * d22 - d12 * d21) * tmp;
}

// solve stage (grad_lsq.f90:168-306, grad_lsq_dm.f90:240-445)
FCM_HD void fcg_grad_lsq_row(const fcm_geom &g, const fcm_c2f &m, const fcm_slots &sl, int weighted, const double *dmat,
                             const double *fi, double *out, int c) {
  double b1 = 0.0, b2 = 0.0, b3 = 0.0;
  for (int q = m.off[c]; q < m.off[c + 1]; ++q) {
    const int fe = m.face[q];
    const int fc = fe & 0x7fffffff;
    double Dx, Dy, Dz;
    if (fc < g.F) {
      const int ijp = fe < 0 ? m.other[q] : c, ijn = fe < 0 ? c : m.other[q];
      if (weighted) {
        const double dx = g.xc[ijn] - g.xc[ijp], dy = g.yc[ijn] - g.yc[ijp], dz = g.zc[ijn] - g.zc[ijp];
        const double w = 1.0 / (dx * dx + dy * dy + dz * dz);
        Dx = w * (g.xc[ijn] - g.xc[ijp]) * (fi[ijn] - fi[ijp]);
        Dy = w * (g.yc[ijn] - g.yc[ijp]) * (fi[ijn] - fi[ijp]);
        Dz = w * (g.zc[ijn] - g.zc[ijp]) * (fi[ijn] - fi[ijp]);
      } else {
        Dx = (g.xc[ijn] - g.xc[ijp]) * (fi[ijn] - fi[ijp]);
        Dy = (g.yc[ijn] - g.yc[ijp]) * (fi[ijn] - fi[ijp]);
        Dz = (g.zc[ijn] - g.zc[ijp]) * (fi[ijn] - fi[ijp]);
      }
    } else {
      const int ijn = m.other[q];
      if (weighted) {
        int kind = 0;
        for (int b = 0; b < 5; ++b)
          if (ijn >= sl.slot[b] && ijn < sl.slot[b] + sl.count[b]) kind = b;
        const int i = ijn - sl.slot[kind];   // 0-based running index of the kind: xf(i) (grad_lsq_dm.f90:285)
        const double ex = g.xf[i] - g.xc[c], ey = g.yf[i] - g.yc[c], ez = g.zf[i] - g.zc[c];
        const double w = 1.0 / (ex * ex + ey * ey + ez * ez);
        Dx = w * (fi[ijn] - fi[c]) * (g.xf[fc] - g.xc[c]);
        Dy = w * (fi[ijn] - fi[c]) * (g.yf[fc] - g.yc[c]);
        Dz = w * (fi[ijn] - fi[c]) * (g.zf[fc] - g.zc[c]);
      } else {
        Dx = (fi[ijn] - fi[c]) * (g.xf[fc] - g.xc[c]);
        Dy = (fi[ijn] - fi[c]) * (g.yf[fc] - g.yc[c]);
        Dz = (fi[ijn] - fi[c]) * (g.zf[fc] - g.zc[c]);
      }
    }
    b1 = b1 + Dx; b2 = b2 + Dy; b3 = b3 + Dz;
  }
  FCM_G3(out, 0, c) = b1 * FCG_DM(1, c) - b2 * FCG_DM(2, c) + b3 * FCG_DM(3, c);
  FCM_G3(out, 1, c) = b1 * FCG_DM(4, c) - b2 * FCG_DM(5, c) - b3 * FCG_DM(6, c);
  FCM_G3(out, 2, c) = b1 * FCG_DM(7, c) - b2 * FCG_DM(8, c) + b3 * FCG_DM(9, c);
}

// ---- grad_lsq_qr ----
FCM_HD double fcg_lapy2(double x, double y) {
  const double xa = fabs(x), ya = fabs(y), w = xa > ya ? xa : ya, z = xa > ya ? ya : xa;
  if (z == 0.0) return w;
  return w * sqrt(1.0 + (z / w) * (z / w));
}

// matrix stage (grad_lsq_qr.f90:62-247): D(3,6,numCells) = R1^-1 Q1^T.  Returns 1 when the cell does not have
// exactly six neighbours (the routine is not defined for it; its D is zeroed).
// `npro`: processor faces of a partitioned mesh count as cell neighbours (the halo cell n + i;
// src-parallel/grad_lsq_qr.f90:55-64), in the position the cell-to-face map gives them: after the inner faces.
FCM_HD int fcg_lsq_qr_matrix_row(const fcm_geom &g, const fcm_c2f &m, double *D, int c, int npro = 0) {
  double *Dc = D + 18 * (size_t)c;
  const int qs = m.off[c], qe = m.off[c + 1];
  if (qe - qs != 6) {
    for (int k = 0; k < 18; ++k) Dc[k] = 0.0;
    return 1;
  }
  double A[18], tau[3];   // column-major 6 x 3: A[j*6 + r]
  for (int r = 0; r < 6; ++r) {
    const int q = qs + r, fe = m.face[q], fc = fe & 0x7fffffff;
    if (fc < g.F || m.other[q] < g.n + npro) {
      const int o = m.other[q];
      A[r] = g.xc[o] - g.xc[c]; A[6 + r] = g.yc[o] - g.yc[c]; A[12 + r] = g.zc[o] - g.zc[c];
    } else {
      A[r] = g.xf[fc] - g.xc[c]; A[6 + r] = g.yf[fc] - g.yc[c]; A[12 + r] = g.zf[fc] - g.zc[c];
    }
  }
  for (int i = 0; i < 3; ++i) {   // DGEQR2
    const double alpha = A[i * 6 + i];
    double ss = 0.0;
    for (int r = i + 1; r < 6; ++r) ss = ss + A[i * 6 + r] * A[i * 6 + r];
    const double xnorm = sqrt(ss);
    if (xnorm == 0.0) { tau[i] = 0.0; continue; }
    const double beta = -copysign(fcg_lapy2(alpha, xnorm), alpha);
    tau[i] = (beta - alpha) / beta;
    const double sc = 1.0 / (alpha - beta);
    for (int r = i + 1; r < 6; ++r) A[i * 6 + r] = A[i * 6 + r] * sc;
    A[i * 6 + i] = beta;
    for (int j = i + 1; j < 3; ++j) {
      double w = A[j * 6 + i];
      for (int r = i + 1; r < 6; ++r) w = w + A[i * 6 + r] * A[j * 6 + r];
      A[j * 6 + i] = A[j * 6 + i] - tau[i] * w;
      for (int r = i + 1; r < 6; ++r) A[j * 6 + r] = A[j * 6 + r] - tau[i] * w * A[i * 6 + r];
    }
  }
  const double r11 = A[0], r12 = A[6], r13 = A[12], r22 = A[7], r23 = A[13], r33 = A[14];
  double v[3][6], Q12[6][6], Q[6][3];
  for (int i = 0; i < 3; ++i)
    for (int r = 0; r < 6; ++r) v[i][r] = r < i ? 0.0 : (r == i ? 1.0 : A[i * 6 + r]);
#define FCG_H(i, r, cc) (((r) == (cc) ? 1.0 : 0.0) + (-tau[i]) * v[i][r] * v[i][cc])
  for (int r = 0; r < 6; ++r)
    for (int cc = 0; cc < 6; ++cc) {
      double s = 0.0;
      for (int k = 0; k < 6; ++k) s = s + FCG_H(0, r, k) * FCG_H(1, k, cc);
      Q12[r][cc] = s;
    }
  for (int r = 0; r < 6; ++r)
    for (int cc = 0; cc < 3; ++cc) {
      double s = 0.0;
      for (int k = 0; k < 6; ++k) s = s + Q12[r][k] * FCG_H(2, k, cc);
      Q[r][cc] = s;
    }
#undef FCG_H
  for (int k = 0; k < 6; ++k) {
    const double q1 = Q[k][0], q2 = Q[k][1], q3 = Q[k][2];
    Dc[3 * k + 0] = q1 / r11 - (r12 * q2) / (r11 * r22) + (q3 * (r12 * r23 - r13 * r22)) / (r11 * r22 * r33);
    Dc[3 * k + 1] = q2 / r22 - (r23 * q3) / (r22 * r33);
    Dc[3 * k + 2] = q3 / r33;
  }
  return 0;
}

// solve stage (grad_lsq_qr.f90:250-330)
FCM_HD void fcg_grad_lsq_qr_row(const fcm_geom &g, const fcm_c2f &m, const double *D, const double *fi, double *out,
                                int c) {
  const double *Dc = D + 18 * (size_t)c;
  const int qs = m.off[c];
  int l = m.off[c + 1] - qs;
  if (l > 6) l = 6;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < l; ++k) {
    const double b = fi[m.other[qs + k]] - fi[c];
    s0 = s0 + Dc[3 * k + 0] * b;
    s1 = s1 + Dc[3 * k + 1] * b;
    s2 = s2 + Dc[3 * k + 2] * b;
  }
  FCM_G3(out, 0, c) = s0; FCM_G3(out, 1, c) = s1; FCM_G3(out, 2, c) = s2;
}

// ---- slope limiters (gradients.f90:263-522): 1 Barth-Jespersen, 2 Venkatakrishnan, 3 mVenkatakrishnan ----
FCM_HD void fcg_limiter_row(const fcm_geom &g, const int *ioffset, const int *ja, const int *diag, int which,
                            const double *phi, double *grad, double glomin, double glomax, double small, int c) {
  const double phi_p = phi[c];
  const int rs = ioffset[c], re = ioffset[c + 1];
  double phi_max = phi[ja[rs]], phi_min = phi_max;
  for (int k = rs + 1; k < re; ++k) {
    const double pv = phi[ja[k]];
    phi_max = FCM_MAX2(phi_max, pv);
    phi_min = FCM_MIN2(phi_max, pv);   // sic (gradients.f90:301)
  }
  const double deltamax = glomax - phi[c], deltamin = glomin - phi[c];
  const double gx = FCM_G3(grad, 0, c), gy = FCM_G3(grad, 1, c), gz = FCM_G3(grad, 2, c);
  double slopelimit = 1.0;
  for (int k = rs; k < re; ++k) {
    if (k == diag[c]) continue;
    const int ijn = ja[k];
    const double gradfiXdr = gx * (g.xc[ijn] - g.xc[c]) + gy * (g.yc[ijn] - g.yc[c]) + gz * (g.zc[ijn] - g.zc[c]);
    if (which == 3) {
      const double cell_neighbour_value = phi_p + gradfiXdr;
      const double deltam = cell_neighbour_value - phi_p;
      double deltap;
      if (deltam > 0.0) deltap = phi_max - phi_p; else deltap = phi_min - phi_p;
      const double epsi = 0.05 * (glomax - glomin);
      const double val = 1.0 / (deltam + small) * ((deltap * deltap + epsi * epsi) * deltam + 2 * (deltam * deltam) * deltap) /
                         (deltap * deltap + 2 * (deltam * deltam) + deltap * deltam + epsi * epsi + small);
      slopelimit = FCM_MAX2(FCM_MIN2(slopelimit, val), 0.0);
    } else {
      double r;
      if (fabs(gradfiXdr) < (double)1.e-6f) r = 1.0;
      else if (gradfiXdr > 0.0) r = deltamax / gradfiXdr;
      else r = deltamin / gradfiXdr;
      if (which == 1) slopelimit = FCM_MIN2(slopelimit, r);
      else slopelimit = FCM_MIN2(slopelimit, (r * r + 2.0 * r) / (r * r + r + 2.0));
    }
  }
  FCM_G3(grad, 0, c) = slopelimit * gx;
  FCM_G3(grad, 1, c) = slopelimit * gy;
  FCM_G3(grad, 2, c) = slopelimit * gz;
}

// The limiters of src-parallel/gradients.f90 (glomin / glomax already reduced over the ranks).  Barth-Jespersen and
// Venkatakrishnan are the serial loops; the modified Venkatakrishnan limiter of the parallel build takes phimax / phimin
// from set_phi_min_max -- the cell, its inner-face neighbours and its processor-face neighbours, true min and max (order
// independent, so the walk over the cell-to-face map gives the reference's values exactly).
FCM_HD void fcg_limiter_row_par(const fcm_geom &g, const fcm_c2f &m, int npro, const int *ioffset, const int *ja,
                                const int *diag, int which, const double *phi, double *grad, double glomin,
                                double glomax, double small, int c) {
  const double phi_p = phi[c];
  double phimax = phi_p, phimin = phi_p;
  if (which == 3)
    for (int q = m.off[c]; q < m.off[c + 1]; ++q) {
      const int fc = m.face[q] & 0x7fffffff, o = m.other[q];
      if (fc < g.F || o < g.n + npro) {
        phimax = FCM_MAX2(phimax, phi[o]);
        phimin = FCM_MIN2(phimin, phi[o]);
      }
    }
  const double deltamax = glomax - phi[c], deltamin = glomin - phi[c];
  const double gx = FCM_G3(grad, 0, c), gy = FCM_G3(grad, 1, c), gz = FCM_G3(grad, 2, c);
  double slopelimit = 1.0;
  for (int k = ioffset[c]; k < ioffset[c + 1]; ++k) {
    if (k == diag[c]) continue;
    const int ijn = ja[k];
    const double gradfiXdr = gx * (g.xc[ijn] - g.xc[c]) + gy * (g.yc[ijn] - g.yc[c]) + gz * (g.zc[ijn] - g.zc[c]);
    if (which == 3) {
      const double cell_neighbour_value = phi_p + gradfiXdr;
      const double deltam = cell_neighbour_value - phi_p;
      double deltap;
      if (deltam > 0.0) deltap = phimax - phi_p; else deltap = phimin - phi_p;
      const double epsi = 0.05 * (glomax - glomin);
      const double val = 1.0 / (deltam + small) * ((deltap * deltap + epsi * epsi) * deltam + 2 * (deltam * deltam) * deltap) /
                         (deltap * deltap + 2 * (deltam * deltam) + deltap * deltam + epsi * epsi + small);
      slopelimit = FCM_MAX2(FCM_MIN2(slopelimit, val), 0.0);
    } else {
      double r;
      if (fabs(gradfiXdr) < (double)1.e-6f) r = 1.0;
      else if (gradfiXdr > 0.0) r = deltamax / gradfiXdr;
      else r = deltamin / gradfiXdr;
      if (which == 1) slopelimit = FCM_MIN2(slopelimit, r);
      else slopelimit = FCM_MIN2(slopelimit, (r * r + 2.0 * r) / (r * r + r + 2.0));
    }
  }
  FCM_G3(grad, 0, c) = slopelimit * gx;
  FCM_G3(grad, 1, c) = slopelimit * gy;
  FCM_G3(grad, 2, c) = slopelimit * gz;
}

// ---- several Gauss gradients in one walk (opt-in, FC_TUNE_FUSED_GRAD) ----
// grad_gauss (grad_gauss.f90:43-113; gradco :128-190; gradbc :194-211) applied to NF fields at once: calcuvw
// (:59-61) and calcp (:38-40) always ask for the three velocity gradients together (calcuvw also for the first
// pressure stage right after them), and two thirds of what one pass reads -- the cell-to-face map, the face vectors
// and the interpolation factors -- does not depend on the field.  Per field the expressions and their order are those
// of k_grad_pass (fc_assemble.cu), so each gradient is bit-identical to the one-field pass.  npro / fpro: processor
// faces of the multi-rank build (src-parallel/grad_gauss.f90:68-75), halo cell = n + i.
constexpr int FCG_MAXF = 4;
struct fcg_gaussn {
  int npro;
  const double *fpro;
  const double *phi[FCG_MAXF];   // e.g. u, v, w, p  [numTotal]
  const double *old[FCG_MAXF];   // gradients of the previous pass (HAS_OLD) or unused
  double *out[FCG_MAXF];         // (3,numCells)
};

template <int NF, bool HAS_OLD>
FCM_HD void fcg_gaussn_row(const fcm_geom &g, const fcm_c2f &m, const fcg_gaussn &k, int c) {
  double gx[NF], gy[NF], gz[NF];
  for (int t = 0; t < NF; ++t) gx[t] = gy[t] = gz[t] = 0.0;
  const int s = m.off[c], e = m.off[c + 1];
  for (int q = s; q < e; ++q) {
    const int fe = m.face[q];
    const int f = fe & 0x7fffffff;
    const int o = m.other[q];
    const double sx = g.arx[f], sy = g.ary[f], sz = g.arz[f];
    if (f < g.F || o < g.n + k.npro) {
      const bool nb = fe < 0;
      const int ijp = nb ? o : c, ijn = nb ? c : o;
      const double fxn = (f < g.F) ? g.facint[f] : k.fpro[o - g.n], fxp = 1.0 - fxn;
      double dx = 0.0, dy = 0.0, dz = 0.0;
      if (HAS_OLD) {
        const double xi = g.xc[ijp] * fxp + g.xc[ijn] * fxn;
        const double yi = g.yc[ijp] * fxp + g.yc[ijn] * fxn;
        const double zi = g.zc[ijp] * fxp + g.zc[ijn] * fxn;
        dx = g.xf[f] - xi; dy = g.yf[f] - yi; dz = g.zf[f] - zi;
      }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int t = 0; t < NF; ++t) {
        const double *phi = k.phi[t];
        double fie = phi[ijp] * fxp + phi[ijn] * fxn;
        if (HAS_OLD) {
          const double *dfo = k.old[t];
          const double dfxi = FCM_G3(dfo, 0, ijp) * fxp + FCM_G3(dfo, 0, ijn) * fxn;
          const double dfyi = FCM_G3(dfo, 1, ijp) * fxp + FCM_G3(dfo, 1, ijn) * fxn;
          const double dfzi = FCM_G3(dfo, 2, ijp) * fxp + FCM_G3(dfo, 2, ijn) * fxn;
          fie = fie + dfxi * dx + dfyi * dy + dfzi * dz;
        }
        const double dfxe = fie * sx, dfye = fie * sy, dfze = fie * sz;
        if (nb) { gx[t] = gx[t] - dfxe; gy[t] = gy[t] - dfye; gz[t] = gz[t] - dfze; }
        else    { gx[t] = gx[t] + dfxe; gy[t] = gy[t] + dfye; gz[t] = gz[t] + dfze; }
      }
    } else {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int t = 0; t < NF; ++t) {
        const double fi = k.phi[t][o];
        gx[t] = gx[t] + fi * sx; gy[t] = gy[t] + fi * sy; gz[t] = gz[t] + fi * sz;
      }
    }
  }
  const double volr = 1.0 / g.vol[c];
  for (int t = 0; t < NF; ++t) {
    FCM_G3(k.out[t], 0, c) = gx[t] * volr;
    FCM_G3(k.out[t], 1, c) = gy[t] * volr;
    FCM_G3(k.out[t], 2, c) = gz[t] * volr;
  }
}
