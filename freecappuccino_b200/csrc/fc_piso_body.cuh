// PISO / PIMPLE pressure equation (src/PISO_multiple_correction.f90, src/PIMPLE_multiple_correction.f90,
// src/get_rAU_x_UEqnH.f90): per-index bodies of the kernels the driver fc_piso_dev
// (fc_assemble.cu) launches on top of the pressure-correction path's own kernels (facefluxmass_piso face
// kernel, row gather, adjustMassFlow, iccg, bpres, Gauss gradient, correctBoundaryConditionsVelocity).
//
// Like fc_momentum_body.cuh these are plain functions of one index so that tests/kernel_bodies_host can run
// them with g++ against the oracle on a GPU-less machine (test infrastructure; the library has no host path).
#pragma once
#include "fc_momentum_body.cuh"

// get_rAU_x_UEqnH.f90:24-208: su, sv, sw = volume sources [+ Crank-Nicolson terms] - sum_nb h(P,nb) phi_nb,
// accumulated per cell in the reference's order (sources, cn face terms ascending, cn cell term, H face terms
// ascending).  h = copy of the momentum matrix taken before the corrector loop (`h = a`, PISO :84).
struct fcp_hbya {
  const int *ioffset, *diag;
  const double *h;
  const double *u, *v, *w, *uo, *vo, *wo, *uoo, *voo, *woo, *t, *den;
  double *su, *sv, *sw;
  int bdf;
  double btime, timestep;
  int cn, lbuoy, boussinesq;
  double beta, tref, densit, gravx, gravy, gravz;
  // several ranks (src-parallel/get_rAU_x_UEqnH.f90): processor-face terms with the CURRENT apr (momentum coefficients
  // in the first corrector, the pressure equation's afterwards).  The reference writes them to su for ALL THREE
  // components (`su(ijp) = su(ijp) - apr(i)*v(ijn)` in the v and w blocks too); restated as written.
  const double *apr;
  int npro;
};

// One component.  `s` = its own source so far; `*su` = the u source, which also collects the processor-face terms of
// this component (for the u component the caller passes su = &s's storage, see fcp_hbya_row).
FCM_HD double fcp_hbya_component(const fcm_geom &g, const fcm_c2f &m, const fcp_hbya &k, int c, double s,
                                 const double *phi, const double *phio, double *su, bool is_u) {
  const int qs = m.off[c], qe = m.off[c + 1];
  if (k.cn) {
    for (int q = qs; q < qe; ++q)
      if ((m.face[q] & 0x7fffffff) < g.F) s = s - k.h[m.pos[q]] * phio[m.other[q]];
    if (k.npro > 0) {
      double t = is_u ? s : *su;
      for (int q = qs; q < qe; ++q) {
        const int fc = m.face[q] & 0x7fffffff, ijn = m.other[q];
        if (fc >= g.F && ijn < g.n + k.npro) {
          t = t - k.apr[ijn - g.n] * phio[ijn];
          t = t + k.apr[ijn - g.n] * phio[c];
        }
      }
      if (is_u) s = t; else *su = t;
    }
    const double apotime = k.den[c] * g.vol[c] / k.timestep;
    double sum = 0.0;
    for (int p = k.ioffset[c]; p < k.ioffset[c + 1]; ++p) sum = sum + k.h[p];
    const double off = sum - k.h[k.diag[c]];
    s = s + (apotime + off) * phio[c];
  }
  for (int q = qs; q < qe; ++q)
    if ((m.face[q] & 0x7fffffff) < g.F) s = s - k.h[m.pos[q]] * phi[m.other[q]];
  if (k.npro > 0) {
    double t = is_u ? s : *su;
    for (int q = qs; q < qe; ++q) {
      const int fc = m.face[q] & 0x7fffffff, ijn = m.other[q];
      if (fc >= g.F && ijn < g.n + k.npro) t = t - k.apr[ijn - g.n] * phi[ijn];
    }
    if (is_u) s = t; else *su = t;
  }
  return s;
}

FCM_HD void fcp_hbya_row(const fcm_geom &g, const fcm_c2f &m, const fcp_hbya &k, int c) {
  double su = 0.0, sv = 0.0, sw = 0.0;
  const double vol = g.vol[c];
  if (k.lbuoy) {
    double heat = 0.0;
    if (k.boussinesq) heat = k.beta * k.densit * (k.t[c] - k.tref) * vol;
    else heat = (k.densit - k.den[c]) * vol;
    su = su - k.gravx * heat;
    sv = sv - k.gravy * heat;
    sw = sw - k.gravz * heat;
  }
  if (k.bdf) {
    const double apotime = k.den[c] * vol / k.timestep;
    double sut = apotime * ((1 + k.btime) * k.uo[c]);
    double svt = apotime * ((1 + k.btime) * k.vo[c]);
    double swt = apotime * ((1 + k.btime) * k.wo[c]);
    if (k.btime > (double)0.99f) {
      sut = sut - apotime * (0.5 * k.btime * k.uoo[c]);
      svt = svt - apotime * (0.5 * k.btime * k.voo[c]);
      swt = swt - apotime * (0.5 * k.btime * k.woo[c]);
    }
    su = su + sut; sv = sv + svt; sw = sw + swt;
  }
  su = fcp_hbya_component(g, m, k, c, su, k.u, k.uo, &su, true);
  k.sv[c] = fcp_hbya_component(g, m, k, c, sv, k.v, k.vo, &su, false);
  k.sw[c] = fcp_hbya_component(g, m, k, c, sw, k.w, k.wo, &su, false);
  k.su[c] = su;
}

// u(1:numCells) = apu*su ...   (get_rAU_x_UEqnH.f90:203-205) -- a separate pass: the H sums read the old field
FCM_HD void fcp_hbya_scale(int c, const double *apu, const double *apv, const double *apw, const double *su,
                           const double *sv, const double *sw, double *u, double *v, double *w) {
  u[c] = apu[c] * su[c];
  v[c] = apv[c] * sv[c];
  w[c] = apw[c] * sw[c];
}

// reference pressure: clear the ROW of pRefCell, unit diagonal, su = p(pRefCell) | pp(pRefCell)   (PISO :188-192)
FCM_HD void fcp_pin_row(const int *ioffset, const int *diag, double *a, double *su, const double *src, int pref) {
  for (int k = ioffset[pref]; k < ioffset[pref + 1]; ++k) a[k] = 0.0;
  a[diag[pref]] = 1.0;
  su[pref] = src[pref];
}

// flmass(iface) += a(icell_jcell(iface)) * (pp(ijn) - pp(ijp)), read from the MATRIX (PISO :253-264)
FCM_HD void fcp_flux_correct(const fcm_geom &g, const int *icj, const double *a, const double *pp, double *flmass,
                             int i) {
  flmass[i] = flmass[i] + a[icj[i]] * (pp[g.neigh[i]] - pp[g.owner[i]]);
}

// u(inp) = u(inp) - apu(inp)*dPdxi(1,inp)*vol(inp)   (PISO :298-302; note the order of the product)
FCM_HD void fcp_velocity_correct(const fcm_geom &g, const double *apu, const double *apv, const double *apw,
                                 const double *dP, double *u, double *v, double *w, int c) {
  const double vol = g.vol[c];
  u[c] = u[c] - apu[c] * FCM_G3(dP, 0, c) * vol;
  v[c] = v[c] - apv[c] * FCM_G3(dP, 1, c) * vol;
  w[c] = w[c] - apw[c] * FCM_G3(dP, 2, c) * vol;
}

// p(inp) = p(inp) + urf(ip)*( pp(inp) - p(inp) )   (PIMPLE :282-284)
FCM_HD void fcp_relax_p(double urf, const double *pp, double *p, int c) { p[c] = p[c] + urf * (pp[c] - p[c]); }
