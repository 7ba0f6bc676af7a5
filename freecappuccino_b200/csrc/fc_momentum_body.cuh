// Momentum predictor `calcuvw` (src/calcuvw.f90:3-557), the caller immediately before the
// pressure-correction path: per-index bodies of the kernels in fc_momentum.cu.
//
// Reference: calcuvw.f90; facefluxuvw / facefluxuvw_boundary src/faceflux_velocity.f90:37-196, 385-549;
// sngrad 'skewness' src/gradients.f90:547-668; face_value and its schemes src/interpolation.f90:19-403;
// calcPressDiv src/fieldManipulation.f90:57-165 with presFaceDivInner :395-445 (whose df(ijp,k)
// addressing of a (3,numCells) array is kept: flat elements ijp, ijp+3, ijp+6).
//
// Same decomposition as the pressure assembly (fc_assemble.cu): every face quantity is computed ONCE by
// a face-parallel kernel (fcm_face), and a cell-parallel kernel (fcm_row) walks the cell's faces through
// the cell-to-face map in the reference's loop order and accumulates left to right -- no scatter, no
// float atomics, and with -fmad=false bit-identical sums.  A third body (fcm_component) is the
// per-component diagonal / under-relaxation / ap* step that precedes each bicgstab call.
//
// The bodies are plain functions of one index so that the thin __global__ wrappers are the only
// device-specific code; tests/kernel_bodies_host compiles THIS header with g++ to check the index
// logic against the oracle on machines without a GPU (test infrastructure only -- the library itself
// has no host path and fc_create fails without a device).
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef __CUDACC__
#define FCM_HD __host__ __device__ __forceinline__
#else
#define FCM_HD static inline
#endif

struct fcm_geom {
  const int *owner, *neigh;
  const double *xc, *yc, *zc, *vol;
  const double *arx, *ary, *arz, *xf, *yf, *zf, *facint;
  int n, F;
};
struct fcm_c2f {
  const int *off, *face, *other, *pos;
};
struct fcm_slots {  // 0-based first slot / first face / count: inlet, outlet, symmetry, wall, prOutlet
  int slot[5], face[5], count[5];
};
struct fcm_flow {
  const double *u, *v, *w, *p, *den, *vis;           // [numTotal]
  const double *flmass, *fmi, *fmo;                  // [F], [ninl], [nout]
  const double *dU, *dV, *dW, *dP;                   // (3,numCells)
  const double *uo, *vo, *wo, *uoo, *voo, *woo, *t;  // [numTotal]
};
struct fcm_opts {
  int scheme, limiter;
  double gds;
  int bdf;
  double btime, timestep;
  int cn, const_mflux;
  double gradPcmf;
  int lbuoy, boussinesq;
  double beta, tref, densit, gravx, gravy, gravz, viscos;
};
struct fcm_faces {  // per inner face, written by fcm_face and read by fcm_row
  double *can, *cap, *sup, *svp, *swp, *fie;
};
struct fcm_rows {
  double *a, *su, *sv, *sw, *spu, *spv, *sp;
};
// processor-boundary faces (src-parallel): face = pface0 + i, halo cell = n + i, i = 0..npro-1
struct fcm_proc {
  int npro, pface0;
  const double *fpro, *fmpro;          // interpolation factor, mass flux of the processor faces
  double *apr;                         // written: can (the coupling coefficient of the SpMV strip)
  double *sup, *svp, *swp, *fie;       // per processor face, written by fcm_proc_face and read by fcm_row
};

#define FCM_G3(p, c, i) ((p)[3 * (size_t)(i) + (c)])
#define FCM_MAX2(a, b) (((a) > (b)) ? (a) : (b))
#define FCM_MIN2(a, b) (((a) < (b)) ? (a) : (b))

// sngrad_scalar_field, approach 'skewness', nrelax = 0 (gradients.f90:595-668)
FCM_HD void fcm_sngrad(const fcm_geom &g, int ijp, int ijn, double arx, double ary, double arz, double lambda,
                       const double *fi, const double *dF, double &dfixi, double &dfiyi, double &dfizi,
                       double &dfixii, double &dfiyii, double &dfizii) {
  const double fxn = lambda, fxp = 1.0 - lambda;
  const double xpn = g.xc[ijn] - g.xc[ijp];
  const double ypn = g.yc[ijn] - g.yc[ijp];
  const double zpn = g.zc[ijn] - g.zc[ijp];
  const double costn = 1.0;
  const double vole = xpn * arx + ypn * ary + zpn * arz;
  dfixi = FCM_G3(dF, 0, ijp) * fxp + FCM_G3(dF, 0, ijn) * fxn;
  dfiyi = FCM_G3(dF, 1, ijp) * fxp + FCM_G3(dF, 1, ijn) * fxn;
  dfizi = FCM_G3(dF, 2, ijp) * fxp + FCM_G3(dF, 2, ijn) * fxn;
  const double d2x = xpn * costn, d2y = ypn * costn, d2z = zpn * costn;
  const double rem = fi[ijn] - fi[ijp] - dfixi * d2x - dfiyi * d2y - dfizi * d2z;
  dfixii = dfixi * costn + arx / vole * rem;
  dfiyii = dfiyi * costn + ary / vole * rem;
  dfizii = dfizi * costn + arz / vole * rem;
}

// face_value (interpolation.f90:19-57) and the schemes it dispatches to.  `lambda` is what the caller
// passes: fxp for the p->n direction, fxn for n->p (faceflux_velocity.f90:160-170).
FCM_HD double fcm_face_value(const fcm_geom &g, int scheme, int limiter, int ijp, int ijn, double xf, double yf,
                             double zf, double lambda, const double *u, const double *dU) {
  if (scheme == 0) {  // face_value_cds :62-90
    const double fxn = lambda, fxp = 1.0 - lambda;
    return u[ijp] * fxp + u[ijn] * fxn;
  }
  if (scheme == 1) {  // face_value_cds_corrected :96-139
    const double fxn = lambda, fxp = 1.0 - lambda;
    const double xi = g.xc[ijp] * fxp + g.xc[ijn] * fxn;
    const double yi = g.yc[ijp] * fxp + g.yc[ijn] * fxn;
    const double zi = g.zc[ijp] * fxp + g.zc[ijn] * fxn;
    const double dfixi = FCM_G3(dU, 0, ijp) * fxp + FCM_G3(dU, 0, ijn) * fxn;
    const double dfiyi = FCM_G3(dU, 1, ijp) * fxp + FCM_G3(dU, 1, ijn) * fxn;
    const double dfizi = FCM_G3(dU, 2, ijp) * fxp + FCM_G3(dU, 2, ijn) * fxn;
    return u[ijp] * fxp + u[ijn] * fxn + (dfixi * (xf - xi) + dfiyi * (yf - yi) + dfizi * (zf - zi));
  }
  if (scheme == 2) {  // face_value_central :145-199
    const double gradfidr = FCM_G3(dU, 0, ijp) * (xf - g.xc[ijp]) + FCM_G3(dU, 1, ijp) * (yf - g.yc[ijp]) +
                            FCM_G3(dU, 2, ijp) * (zf - g.zc[ijp]) + FCM_G3(dU, 0, ijn) * (xf - g.xc[ijn]) +
                            FCM_G3(dU, 1, ijn) * (yf - g.yc[ijn]) + FCM_G3(dU, 2, ijn) * (zf - g.zc[ijn]);
    return 0.5 * (u[ijp] + u[ijn] + gradfidr);
  }
  if (scheme == 3) {  // face_value_2nd_upwind :205-252
    const double gradfidr = FCM_G3(dU, 0, ijp) * (xf - g.xc[ijp]) + FCM_G3(dU, 1, ijp) * (yf - g.yc[ijp]) +
                            FCM_G3(dU, 2, ijp) * (zf - g.zc[ijp]);
    return u[ijp] + gradfidr;
  }
  if (scheme == 5) {  // face_value_2nd_upwind_flux_limiter :324-401
    const double fxp = 1.0 - lambda;
    const double xpn = g.xc[ijn] - g.xc[ijp];
    const double ypn = g.yc[ijn] - g.yc[ijp];
    const double zpn = g.zc[ijn] - g.zc[ijp];
    const double r = (2 * FCM_G3(dU, 0, ijp) * xpn + 2 * FCM_G3(dU, 1, ijp) * ypn + 2 * FCM_G3(dU, 2, ijp) * zpn) /
                         (u[ijn] - u[ijp]) - 1.0;
    double psi;
    switch (limiter) {
      case 0: psi = FCM_MAX2(0.0, FCM_MIN2(FCM_MIN2(2.0 * r, 0.75 * r + 0.25), 4.0)); break;
      case 1: psi = FCM_MAX2(0.0, FCM_MIN2(FCM_MIN2(1.5 * r, 0.75 * r + 0.25), 2.5)); break;
      case 2: psi = FCM_MAX2(0.0, FCM_MIN2(FCM_MIN2(2.0 * r, 0.5 * r + 0.5), 2.0)); break;
      case 3: psi = FCM_MAX2(0.0, FCM_MIN2(FCM_MIN2(FCM_MIN2(2.0 * r, 0.75 * r + 0.25), 0.25 * r + 0.75), 2.0)); break;
      case 4: psi = FCM_MAX2(0.0, FCM_MIN2(FCM_MIN2(2.0 * r, 2.0 / 3.0 * r + 1.0 / 3.0), 2.0)); break;
      case 5: psi = (r + fabs(r)) * (3 * r + 1.0) / (2 * ((r + 1.0) * (r + 1.0))); break;
      case 6: psi = 1.5 * r * (r + 1.0) / (r * r + r + 1.0); break;
      default: psi = 1.0; break;
    }
    return u[ijp] + fxp * psi * (u[ijn] - u[ijp]);
  }
  {  // face_value_muscl :258-318 (also face_value's fall-through)
    const double theta = 0.125;
    const double up = FCM_G3(dU, 0, ijp) * (xf - g.xc[ijp]) + FCM_G3(dU, 1, ijp) * (yf - g.yc[ijp]) +
                      FCM_G3(dU, 2, ijp) * (zf - g.zc[ijp]);
    const double ce = FCM_G3(dU, 0, ijp) * (xf - g.xc[ijp]) + FCM_G3(dU, 1, ijp) * (yf - g.yc[ijp]) +
                      FCM_G3(dU, 2, ijp) * (zf - g.zc[ijp]) + FCM_G3(dU, 0, ijn) * (xf - g.xc[ijn]) +
                      FCM_G3(dU, 1, ijn) * (yf - g.yc[ijn]) + FCM_G3(dU, 2, ijn) * (zf - g.zc[ijn]);
    const double fv_up = (u[ijp] + up);
    const double fv_ce = 0.5 * (u[ijp] + u[ijn] + ce);
    return theta * fv_ce + (1.0 - theta) * fv_up;
  }
}

// facefluxuvw (faceflux_velocity.f90:37-196) + presFaceDivInner (fieldManipulation.f90:395-445) of one face between
// cells ijp and ijn (ijn may be a halo cell); fidx = the face's index in the geometry arrays
struct fcm_face_val {
  double can, cap, sup, svp, swp, fie;
};
FCM_HD fcm_face_val fcm_face_core(const fcm_geom &g, const fcm_flow &f, const fcm_opts &o, int ijp, int ijn, int fidx,
                                  double lambda, double flomass) {
  fcm_face_val r;
  const double xf = g.xf[fidx], yf = g.yf[fidx], zf = g.zf[fidx];
  const double arx = g.arx[fidx], ary = g.ary[fidx], arz = g.arz[fidx];
  const double gam = o.gds;
  const double fxn = lambda, fxp = 1.0 - lambda;
  const double xpn = g.xc[ijn] - g.xc[ijp];
  const double ypn = g.yc[ijn] - g.yc[ijp];
  const double zpn = g.zc[ijn] - g.zc[ijp];
  const double dpn = sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
  const double are = sqrt(arx * arx + ary * ary + arz * arz);
  const double game = f.vis[ijp] * fxp + f.vis[ijn] * fxn;
  const double de = game * are / dpn;
  r.can = -de + FCM_MIN2(flomass, 0.0);
  r.cap = -de - FCM_MAX2(flomass, 0.0);
  double duxi, duyi, duzi, dvxi, dvyi, dvzi, dwxi, dwyi, dwzi;
  double duxii, duyii, duzii, dvxii, dvyii, dvzii, dwxii, dwyii, dwzii;
  fcm_sngrad(g, ijp, ijn, arx, ary, arz, lambda, f.u, f.dU, duxi, duyi, duzi, duxii, duyii, duzii);
  fcm_sngrad(g, ijp, ijn, arx, ary, arz, lambda, f.v, f.dV, dvxi, dvyi, dvzi, dvxii, dvyii, dvzii);
  fcm_sngrad(g, ijp, ijn, arx, ary, arz, lambda, f.w, f.dW, dwxi, dwyi, dwzi, dwxii, dwyii, dwzii);
  const double fdue = game * ((duxii + duxii) * arx + (duyii + dvxii) * ary + (duzii + dwxii) * arz);
  const double fdve = game * ((duyii + dvxii) * arx + (dvyii + dvyii) * ary + (dvzii + dwyii) * arz);
  const double fdwe = game * ((duzii + dwxii) * arx + (dwyii + dvzii) * ary + (dwzii + dwzii) * arz);
  const double fdui = game * are / dpn * (duxi * xpn + duyi * ypn + duzi * zpn);
  const double fdvi = game * are / dpn * (dvxi * xpn + dvyi * ypn + dvzi * zpn);
  const double fdwi = game * are / dpn * (dwxi * xpn + dwyi * ypn + dwzi * zpn);
  const double fuuds = FCM_MAX2(flomass, 0.0) * f.u[ijp] + FCM_MIN2(flomass, 0.0) * f.u[ijn];
  const double fvuds = FCM_MAX2(flomass, 0.0) * f.v[ijp] + FCM_MIN2(flomass, 0.0) * f.v[ijn];
  const double fwuds = FCM_MAX2(flomass, 0.0) * f.w[ijp] + FCM_MIN2(flomass, 0.0) * f.w[ijn];
  double ue, ve, we;
  if (flomass >= 0.0) {
    ue = fcm_face_value(g, o.scheme, o.limiter, ijp, ijn, xf, yf, zf, fxp, f.u, f.dU);
    ve = fcm_face_value(g, o.scheme, o.limiter, ijp, ijn, xf, yf, zf, fxp, f.v, f.dV);
    we = fcm_face_value(g, o.scheme, o.limiter, ijp, ijn, xf, yf, zf, fxp, f.w, f.dW);
  } else {
    ue = fcm_face_value(g, o.scheme, o.limiter, ijn, ijp, xf, yf, zf, fxn, f.u, f.dU);
    ve = fcm_face_value(g, o.scheme, o.limiter, ijn, ijp, xf, yf, zf, fxn, f.v, f.dV);
    we = fcm_face_value(g, o.scheme, o.limiter, ijn, ijp, xf, yf, zf, fxn, f.w, f.dW);
  }
  const double fuhigh = flomass * ue, fvhigh = flomass * ve, fwhigh = flomass * we;
  r.sup = -gam * (fuhigh - fuuds) + fdue - fdui;
  r.svp = -gam * (fvhigh - fvuds) + fdve - fdvi;
  r.swp = -gam * (fwhigh - fwuds) + fdwe - fdwi;
  // pressure at the face centre with the reference's df(ijp,k) addressing: flat ijp, ijp+3, ijp+6
  {
    const double xi = g.xc[ijp] * fxp + g.xc[ijn] * fxn;
    const double yi = g.yc[ijp] * fxp + g.yc[ijn] * fxn;
    const double zi = g.zc[ijp] * fxp + g.zc[ijn] * fxn;
    const double dfxi = f.dP[(size_t)ijp] * fxp + f.dP[(size_t)ijn] * fxn;
    const double dfyi = f.dP[(size_t)ijp + 3] * fxp + f.dP[(size_t)ijn + 3] * fxn;
    const double dfzi = f.dP[(size_t)ijp + 6] * fxp + f.dP[(size_t)ijn + 6] * fxn;
    r.fie = f.p[ijp] * fxp + f.p[ijn] * fxn + dfxi * (xf - xi) + dfyi * (yf - yi) + dfzi * (zf - zi);
  }
  return r;
}

// one inner face
FCM_HD void fcm_face(const fcm_geom &g, const fcm_flow &f, const fcm_opts &o, const fcm_faces &out, int i) {
  const fcm_face_val r = fcm_face_core(g, f, o, g.owner[i], g.neigh[i], i, g.facint[i], f.flmass[i]);
  out.can[i] = r.can; out.cap[i] = r.cap;
  out.sup[i] = r.sup; out.svp[i] = r.svp; out.swp[i] = r.swp;
  out.fie[i] = r.fie;
}

// one processor-boundary face (src-parallel/calcuvw.f90:225-254, src-parallel/fieldManipulation.f90:130-146):
// the halo cell n + i is the neighbour, fpro(i) the interpolation factor, fmpro(i) the mass flux
FCM_HD void fcm_proc_face(const fcm_geom &g, const fcm_flow &f, const fcm_opts &o, const fcm_proc &P, int i) {
  const int fidx = P.pface0 + i;
  const fcm_face_val r = fcm_face_core(g, f, o, g.owner[fidx], g.n + i, fidx, P.fpro[i], P.fmpro[i]);
  P.apr[i] = r.can;
  P.sup[i] = r.sup; P.svp[i] = r.svp; P.swp[i] = r.swp;
  P.fie[i] = r.fie;
}

// one cell: everything calcuvw.f90:48-383 accumulates into su/sv/sw, spu/spv/sp and the row's
// off-diagonals, in the reference's order: calcPressDiv (inner faces, then every boundary kind),
// the volume sources, the inner-face fluxes, then inlet, outlet, symmetry and wall faces.
FCM_HD void fcm_row(const fcm_geom &g, const fcm_c2f &m, const fcm_slots &sl, const fcm_flow &f, const fcm_opts &o,
                    const fcm_faces &fa, const fcm_proc &P, const fcm_rows &r, int c) {
  double su = 0.0, sv = 0.0, sw = 0.0, spu = 0.0, spv = 0.0, sp = 0.0;
  const int qs = m.off[c], qe = m.off[c + 1];
  // ---- calcPressDiv (fieldManipulation.f90:91-165) ----
  for (int q = qs; q < qe; ++q) {
    const int fe = m.face[q];
    const int fc = fe & 0x7fffffff;
    const double sx = g.arx[fc], sy = g.ary[fc], sz = g.arz[fc];
    if (fc < g.F) {
      const double fie = fa.fie[fc];
      const double dfxe = fie * sx, dfye = fie * sy, dfze = fie * sz;
      if (fe < 0) { su = su + dfxe; sv = sv + dfye; sw = sw + dfze; }
      else        { su = su - dfxe; sv = sv - dfye; sw = sw - dfze; }
    } else if (m.other[q] < g.n + P.npro) {  // processor face: the owner side only (src-parallel :130-146)
      const double fie = P.fie[m.other[q] - g.n];
      su = su - fie * sx; sv = sv - fie * sy; sw = sw - fie * sz;
    } else {  // presFaceDivBoundary :532-554, all five kinds
      const double pb = f.p[m.other[q]];
      su = su - pb * sx; sv = sv - pb * sy; sw = sw - pb * sz;
    }
  }
  // ---- volume sources (calcuvw.f90:75-141) ----
  const double vol = g.vol[c];
  if (o.const_mflux) su = su + o.gradPcmf * vol;
  if (o.lbuoy) {
    double heat;
    if (o.boussinesq) heat = o.beta * o.densit * (f.t[c] - o.tref) * vol;
    else heat = (o.densit - f.den[c]) * vol;
    su = su - o.gravx * heat;
    sv = sv - o.gravy * heat;
    sw = sw - o.gravz * heat;
  }
  if (o.bdf) {
    const double apotime = f.den[c] * vol / o.timestep;
    double sut = apotime * ((1 + o.btime) * f.uo[c]);
    double svt = apotime * ((1 + o.btime) * f.vo[c]);
    double swt = apotime * ((1 + o.btime) * f.wo[c]);
    if (o.btime > (double)0.99f) {
      sut = sut - apotime * (0.5 * o.btime * f.uoo[c]);
      svt = svt - apotime * (0.5 * o.btime * f.voo[c]);
      swt = swt - apotime * (0.5 * o.btime * f.woo[c]);
    }
    su = su + sut; sv = sv + svt; sw = sw + swt;
    spu = spu + apotime * (1 + 0.5 * o.btime);
    spv = spv + apotime * (1 + 0.5 * o.btime);
    sp = sp + apotime * (1 + 0.5 * o.btime);
  }
  // ---- face fluxes (calcuvw.f90:156-383) ----
  for (int q = qs; q < qe; ++q) {
    const int fe = m.face[q];
    const int fc = fe & 0x7fffffff;
    if (fc < g.F) {
      if (fe < 0) {  // this cell is the face's neighbour: a(jcell,icell) = cap, sources with the other sign
        r.a[m.pos[q]] = fa.cap[fc];
        su = su - fa.sup[fc]; sv = sv - fa.svp[fc]; sw = sw - fa.swp[fc];
      } else {
        r.a[m.pos[q]] = fa.can[fc];
        su = su + fa.sup[fc]; sv = sv + fa.svp[fc]; sw = sw + fa.swp[fc];
      }
      continue;
    }
    const int ijb = m.other[q];
    if (ijb < g.n + P.npro) {  // processor face (src-parallel/calcuvw.f90:239-251): coupling stays outside the CSR (apr)
      const int i = ijb - g.n;
      const double can = P.apr[i];
      spu = spu - can; spv = spv - can; sp = sp - can;
      su = su + P.sup[i]; sv = sv + P.svp[i]; sw = sw + P.swp[i];
      continue;
    }
    int kind = -1;
    for (int b = 0; b < 5; ++b)
      if (ijb >= sl.slot[b] && ijb < sl.slot[b] + sl.count[b]) kind = b;
    if (kind < 0 || kind == 4) continue;  // prOutlet has no loop in calcuvw
    const double ax = g.arx[fc], ay = g.ary[fc], az = g.arz[fc];
    const double are = sqrt(ax * ax + ay * ay + az * az);
    if (kind <= 1) {  // inlet / outlet: facefluxuvw_boundary, only cb = can is used (:225-265)
      const double flomass = kind == 0 ? f.fmi[ijb - sl.slot[0]] : f.fmo[ijb - sl.slot[1]];
      const double xpn = g.xf[fc] - g.xc[c], ypn = g.yf[fc] - g.yc[c], zpn = g.zf[fc] - g.zc[c];
      const double dpn = sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
      const double game = f.vis[ijb];
      const double de = game * are / dpn;
      const double cb = -de + FCM_MIN2(flomass, 0.0);
      spu = spu - cb; spv = spv - cb; sp = sp - cb;
      su = su - cb * f.u[ijb];
      sv = sv - cb * f.v[ijb];
      sw = sw - cb * f.w[ijb];
    } else {  // symmetry (:269-312) / wall (:315-383); srds, srdw = are / ((x_f - x_P).n) (init.f90:987-1029)
      const double nxf = ax / are, nyf = ay / are, nzf = az / are;
      const double dn = (g.xf[fc] - g.xc[c]) * nxf + (g.yf[fc] - g.yc[c]) * nyf + (g.zf[fc] - g.zc[c]) * nzf;
      const double srd = are / dn;
      const double visc = kind == 3 ? o.viscos : f.vis[ijb];
      const double cf = visc * srd;
      const double dx = g.xc[c] - g.xf[fc], dy = g.yc[c] - g.yf[fc], dz = g.zc[c] - g.zf[fc];
      const double dpb = sqrt(dx * dx + dy * dy + dz * dz);
      const double vsol = visc * are / dpb;
      const double upb = f.u[c] - f.u[ijb], vpb = f.v[c] - f.v[ijb], wpb = f.w[c] - f.w[ijb];
      spu = spu + vsol; spv = spv + vsol; sp = sp + vsol;
      if (kind == 2) {
        const double fdne = 2 * cf * (upb * nxf + vpb * nyf + wpb * nzf);
        su = su + vsol * f.u[c] - fdne * nxf;
        sv = sv + vsol * f.v[c] - fdne * nyf;
        sw = sw + vsol * f.w[c] - fdne * nzf;
      } else {
        const double vnp = upb * nxf + vpb * nyf + wpb * nzf;
        const double utp = upb - vnp * nxf, vtp = vpb - vnp * nyf, wtp = wpb - vnp * nzf;
        su = su + vsol * f.u[c] - cf * utp;
        sv = sv + vsol * f.v[c] - cf * vtp;
        sw = sw + vsol * f.w[c] - cf * wtp;
      }
    }
  }
  r.su[c] = su; r.sv[c] = sv; r.sw[c] = sw;
  r.spu[c] = spu; r.spv[c] = spv; r.sp[c] = sp;
}

// one cell of one velocity component: Crank-Nicolson sources, main diagonal, under-relaxation, ap*
// (calcuvw.f90:391-436 U, :447-495 V, :504-553 W).  `s`/`spc` = sources of the component (su|sv|sw,
// spu|spv|sp); `su` = the right-hand side the solver reads; zero_diag: V and W zero a(diag) and su
// first (:473-476), U sums the row with the stale diagonal of the previous solve in place (:423).
struct fcm_comp {
  const int *ioffset, *diag;
  double *a, *s, *spc, *su, *ap;
  const double *phi, *phio, *den;
  double urfrs, urfms, small, timestep;
  int cn, zero_diag;
  int parallel;          // src-parallel/calcuvw.f90: running-subtraction diagonal (:485-493), cn terms of the
  int npro;              // processor faces (:450-463) with apr
  const double *apr;
};
FCM_HD void fcm_component(const fcm_geom &g, const fcm_c2f &m, const fcm_comp &k, int c) {
  double s = k.s[c], spc = k.spc[c];
  const int rs = k.ioffset[c], re = k.ioffset[c + 1], dg = k.diag[c];
  if (k.cn) {
    for (int q = m.off[c]; q < m.off[c + 1]; ++q) {
      const int fc = m.face[q] & 0x7fffffff;
      if (fc < g.F) s = s - k.a[m.pos[q]] * k.phio[m.other[q]];
    }
    for (int q = m.off[c]; q < m.off[c + 1]; ++q) {   // processor faces come after the inner ones in the map
      const int fc = m.face[q] & 0x7fffffff;
      if (fc >= g.F && m.other[q] < g.n + k.npro) {
        const double ap_ = k.apr[m.other[q] - g.n];
        s = s - ap_ * k.phio[m.other[q]];
        s = s + ap_ * k.phio[c];
      }
    }
    const double apotime = k.den[c] * g.vol[c] / k.timestep;
    double sum = 0.0;
    for (int p = rs; p < re; ++p) sum = sum + k.a[p];
    const double off = sum - k.a[dg];
    s = s + (apotime + off) * k.phio[c];
    spc = spc + apotime;
    k.s[c] = s;
    k.spc[c] = spc;
  }
  if (k.parallel) {
    double d = spc;
    for (int p = rs; p < re; ++p)
      if (p != dg) d = d - k.a[p];
    d = d * k.urfrs;
    k.a[dg] = d;
    k.su[c] = s + k.urfms * d * k.phi[c];
    k.ap[c] = 1.0 / (d + k.small);
    return;
  }
  const double stale = k.zero_diag ? 0.0 : k.a[dg];
  double sum = 0.0;
  for (int p = rs; p < re; ++p) sum = sum + (p == dg ? stale : k.a[p]);
  const double off = sum - stale;
  double d = spc - off;
  d = d * k.urfrs;
  k.a[dg] = d;
  k.su[c] = s + k.urfms * d * k.phi[c];
  k.ap[c] = 1.0 / (d + k.small);
}
