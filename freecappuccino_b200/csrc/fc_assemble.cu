// Assembly half of the pressure-correction path: Gauss gradients, Rhie-Chow face fluxes,
// the CSR row gather, Laplacian, boundary pressure extrapolation and the post-solve
// corrections of calcp.
//
// Reference (all under src/): calcp-multiple_correction_SIMPLE.f90, facefluxmass.f90,
// interpolation.f90:145-199, grad_gauss.f90, grad_gauss_corrected.f90, fvm_laplacian.f90,
// bpres.f90, adjustMassFlow.f90, correctBoundaryConditionsVelocity.f90, continuityErrors.h.
//
// The reference scatters per face into cells (a(diag(P)) -= can, su(P) -= flux, dudx(P) += ...),
// which on a GPU would need floating-point atomics and would make the sums order-dependent.
// Here every face quantity is computed ONCE by a face-parallel kernel, and a cell-parallel
// kernel then walks the cell's faces through the precomputed cell-to-face map in the
// reference's own loop order (inner faces ascending, then inlet, outlet, symmetry, wall,
// prOutlet) and accumulates left to right.  The result is deterministic and -- because
// the library is built with -fmad=false and IEEE division / sqrt -- bit-identical to the
// Fortran loops.
#include "fc_body_views.cuh"
#include "fc_piso_body.cuh"
#include "fc_reduce.cuh"

namespace {

struct geom_t {
  const int *owner, *neigh;
  const double *xc, *yc, *zc, *vol;
  const double *arx, *ary, *arz, *xf, *yf, *zf, *facint;
  int n, F;
  int npro, pface0;       // processor faces: face = pface0 + i, halo cell = n + i (src-parallel)
  const double *fpro;
};

struct c2f_t {
  const int *off, *face, *other, *pos;
};

struct slots_t {  // 0-based first slot / first face / count per boundary kind: inlet, outlet, symmetry, wall, prOutlet
  int slot[5], face[5], count[5];
};

#define G3(p, c, i) ((p)[3 * (size_t)(i) + (c)])

inline geom_t geom_of(const fc_context *ctx) {
  return geom_t{ctx->owner, ctx->neigh, ctx->xc, ctx->yc, ctx->zc, ctx->vol, ctx->arx, ctx->ary, ctx->arz,
                ctx->xf, ctx->yf, ctx->zf, ctx->facint, ctx->n, ctx->F, ctx->npro, ctx->m.iProcFacesStart, ctx->fpro};
}
inline c2f_t c2f_of(const fc_context *ctx) { return c2f_t{ctx->c2f_off, ctx->c2f_face, ctx->c2f_other, ctx->c2f_pos}; }
inline slots_t slots_of(const fc_context *ctx) {
  const fc_mesh_desc &m = ctx->m;
  slots_t s;
  const int cnt[5] = {m.ninl, m.nout, m.nsym, m.nwal, m.npru};
  const int fst[5] = {m.iInletFacesStart, m.iOutletFacesStart, m.iSymmetryFacesStart, m.iWallFacesStart,
                      m.iPressOutletFacesStart};
  int slot = ctx->n + ctx->npro;
  for (int b = 0; b < 5; ++b) { s.count[b] = cnt[b]; s.face[b] = fst[b]; s.slot[b] = slot; slot += cnt[b]; }
  return s;
}

// ------------------------------------------------------------------------------------------
// Gauss gradient, one pass (grad_gauss.f90:43-113; gradco :128-190; gradbc :194-211)
// ------------------------------------------------------------------------------------------
template <bool HAS_OLD>
__global__ void __launch_bounds__(256)
k_grad_pass(geom_t g, c2f_t m, const double *__restrict__ phi, const double *__restrict__ dfo,
            double *__restrict__ df) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.n) return;
  double gx = 0.0, gy = 0.0, gz = 0.0;
  const int s = m.off[c], e = m.off[c + 1];
  for (int q = s; q < e; ++q) {
    const int fe = m.face[q];
    const int f = fe & 0x7fffffff;
    const int o = m.other[q];
    const double sx = g.arx[f], sy = g.ary[f], sz = g.arz[f];
    if (f < g.F || o < g.n + g.npro) {  // inner face, or processor face (src-parallel/grad_gauss.f90:68-75)
      const bool nb = fe < 0;
      const int ijp = nb ? o : c, ijn = nb ? c : o;
      const double fxn = (f < g.F) ? g.facint[f] : g.fpro[o - g.n], fxp = 1.0 - fxn;
      double fie = phi[ijp] * fxp + phi[ijn] * fxn;
      if (HAS_OLD) {
        const double xi = g.xc[ijp] * fxp + g.xc[ijn] * fxn;
        const double yi = g.yc[ijp] * fxp + g.yc[ijn] * fxn;
        const double zi = g.zc[ijp] * fxp + g.zc[ijn] * fxn;
        const double dfxi = G3(dfo, 0, ijp) * fxp + G3(dfo, 0, ijn) * fxn;
        const double dfyi = G3(dfo, 1, ijp) * fxp + G3(dfo, 1, ijn) * fxn;
        const double dfzi = G3(dfo, 2, ijp) * fxp + G3(dfo, 2, ijn) * fxn;
        fie = fie + dfxi * (g.xf[f] - xi) + dfyi * (g.yf[f] - yi) + dfzi * (g.zf[f] - zi);
      }
      const double dfxe = fie * sx, dfye = fie * sy, dfze = fie * sz;
      if (nb) { gx = gx - dfxe; gy = gy - dfye; gz = gz - dfze; }
      else    { gx = gx + dfxe; gy = gy + dfye; gz = gz + dfze; }
    } else {
      const double fi = phi[o];
      gx = gx + fi * sx; gy = gy + fi * sy; gz = gz + fi * sz;
    }
  }
  const double volr = 1.0 / g.vol[c];
  G3(df, 0, c) = gx * volr;
  G3(df, 1, c) = gy * volr;
  G3(df, 2, c) = gz * volr;
}

// ------------------------------------------------------------------------------------------
// Rhie-Chow mass flux and p' coefficient of one face (facefluxmass.f90:34-199 variant 0,
// facefluxmass2 :204-288 variant 1, facefluxmass_piso :294-358 variant 2;
// face_value_central interpolation.f90:145-199)
// ------------------------------------------------------------------------------------------
struct flow_t {
  const double *u, *v, *w, *p, *den;
  const double *dU, *dV, *dW, *dP;
  const double *apu, *apv, *apw;
};

__device__ __forceinline__ double face_value_central(const geom_t &g, int inp, int inn, double xf, double yf,
                                                     double zf, const double *__restrict__ fi,
                                                     const double *__restrict__ gr) {
  const double gradfidr = G3(gr, 0, inp) * (xf - g.xc[inp]) + G3(gr, 1, inp) * (yf - g.yc[inp]) +
                          G3(gr, 2, inp) * (zf - g.zc[inp]) + G3(gr, 0, inn) * (xf - g.xc[inn]) +
                          G3(gr, 1, inn) * (yf - g.yc[inn]) + G3(gr, 2, inn) * (zf - g.zc[inn]);
  return 0.5 * (fi[inp] + fi[inn] + gradfidr);
}

template <int VARIANT>
__device__ __forceinline__ void facefluxmass(const geom_t &g, const flow_t &f, int ijp, int ijn, double xf, double yf,
                                             double zf, double arx, double ary, double arz, double lambda,
                                             double &cap, double &fluxmass) {
  const double fxn = lambda, fxp = 1.0 - lambda;
  const double xpn = g.xc[ijn] - g.xc[ijp];
  const double ypn = g.yc[ijn] - g.yc[ijp];
  const double zpn = g.zc[ijn] - g.zc[ijp];
  const double dpn = sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
  const double are = sqrt(arx * arx + ary * ary + arz * arz);
  const double dene = f.den[ijp] * fxp + f.den[ijn] * fxn;
  const double ui = face_value_central(g, ijp, ijn, xf, yf, zf, f.u, f.dU);
  const double vi = face_value_central(g, ijp, ijn, xf, yf, zf, f.v, f.dV);
  const double wi = face_value_central(g, ijp, ijn, xf, yf, zf, f.w, f.dW);
  const double *dP = f.dP;
  if (VARIANT == 1) {
    const double Kj = g.vol[ijp] * f.apu[ijp] * fxp + g.vol[ijn] * f.apu[ijn] * fxn;
    cap = -dene * Kj * are / dpn;
    const double dpxi = (G3(dP, 0, ijn) * fxp + G3(dP, 0, ijp) * fxn) * xpn;
    const double dpyi = (G3(dP, 1, ijn) * fxp + G3(dP, 1, ijp) * fxn) * ypn;
    const double dpzi = (G3(dP, 2, ijn) * fxp + G3(dP, 2, ijp) * fxn) * zpn;
    fluxmass = dene * (ui * arx + vi * ary + wi * arz) + cap * (f.p[ijn] - f.p[ijp] - dpxi - dpyi - dpzi);
    return;
  }
  if (VARIANT == 2) {
    cap = -dene * (fxp * g.vol[ijp] * f.apu[ijp] + fxn * g.vol[ijn] * f.apu[ijn]) * are / dpn;
    fluxmass = dene * (ui * arx + vi * ary + wi * arz);
    return;
  }
  const double nxx = arx / are, nyy = ary / are, nzz = arz / are;
  const double Dpu = (fxn * g.vol[ijn] * f.apu[ijn] + fxp * g.vol[ijp] * f.apu[ijp]);
  const double Dpv = (fxn * g.vol[ijn] * f.apv[ijn] + fxp * g.vol[ijp] * f.apv[ijp]);
  const double Dpw = (fxn * g.vol[ijn] * f.apw[ijn] + fxp * g.vol[ijp] * f.apw[ijp]);
  const double sfdpnr = 1.0 / (arx * xpn * nxx + ary * ypn * nyy + arz * zpn * nzz);
  const double smdpn = (arx * arx + ary * ary + arz * arz) * sfdpnr;
  cap = -dene * Dpu * smdpn;
  const double dpxi = Dpu * (fxn * G3(dP, 0, ijn) + fxp * G3(dP, 0, ijp)) * xpn * nxx;
  const double dpyi = Dpv * (fxn * G3(dP, 1, ijn) + fxp * G3(dP, 1, ijp)) * ypn * nyy;
  const double dpzi = Dpw * (fxn * G3(dP, 2, ijn) + fxp * G3(dP, 2, ijp)) * zpn * nzz;
  double xpp = xf - (xf - g.xc[ijp]) * nxx;
  double ypp = yf - (yf - g.yc[ijp]) * nyy;
  double zpp = zf - (zf - g.zc[ijp]) * nzz;
  double xep = xf - (xf - g.xc[ijn]) * nxx;
  double yep = yf - (yf - g.yc[ijn]) * nyy;
  double zep = zf - (zf - g.zc[ijn]) * nzz;
  xpp = xpp - g.xc[ijp]; ypp = ypp - g.yc[ijp]; zpp = zpp - g.zc[ijp];
  xep = xep - g.xc[ijn]; yep = yep - g.yc[ijn]; zep = zep - g.zc[ijn];
  double dpe = (f.p[ijn] - f.p[ijp]);
  // facefluxmass.f90:170-171 -- the reference's operator precedence is kept: only the first
  // ijp term is subtracted
  const double dpecorr = (G3(dP, 0, ijn) * xep + G3(dP, 1, ijn) * yep + G3(dP, 2, ijn) * zep - G3(dP, 0, ijp) * xpp +
                          G3(dP, 1, ijp) * ypp + G3(dP, 2, ijp) * zpp);
  dpe = dpe + dpecorr;
  const double dpex = Dpu * dpe * sfdpnr * arx;
  const double dpey = Dpv * dpe * sfdpnr * ary;
  const double dpez = Dpw * dpe * sfdpnr * arz;
  const double ue = ui - dpex + dpxi;
  const double ve = vi - dpey + dpyi;
  const double we = wi - dpez + dpzi;
  fluxmass = dene * (ue * arx + ve * ary + we * arz);
}

// OCC = CTAs per SM the register allocation must allow (FC_TUNE_FACE_OCC): the kernel is bound by the latency of its
// ~55 gathers per face (ncu: long-scoreboard stalls, 23 % of the warps resident at 100 registers)
template <int VARIANT, int OCC>
__global__ void __launch_bounds__(256, OCC)
k_calcp_faces(geom_t g, flow_t f, double *__restrict__ coef, double *__restrict__ flmass) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.F) return;
  double cap, fl;
  facefluxmass<VARIANT>(g, f, g.owner[i], g.neigh[i], g.xf[i], g.yf[i], g.zf[i], g.arx[i], g.ary[i], g.arz[i],
                        g.facint[i], cap, fl);
  coef[i] = cap;
  flmass[i] = fl;
}

// fluxmc (facefluxmass.f90:520-607): non-orthogonal corrector flux; flmass += fmcor (calcp :195-205)
// facefluxlaplacian (fvm_laplacian.f90:171-221)
__global__ void __launch_bounds__(256)
k_laplacian_faces(geom_t g, const double *__restrict__ mu, double *__restrict__ coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.F) return;
  const int ijp = g.owner[i], ijn = g.neigh[i];
  const double arx = g.arx[i], ary = g.ary[i], arz = g.arz[i];
  const double fxn = g.facint[i], fxp = 1.0 - fxn;
  const double xpn = g.xc[ijn] - g.xc[ijp];
  const double ypn = g.yc[ijn] - g.yc[ijp];
  const double zpn = g.zc[ijn] - g.zc[ijp];
  const double smdpn = (arx * arx + ary * ary + arz * arz) / (arx * xpn + ary * ypn + arz * zpn);
  coef[i] = (fxp * mu[ijp] + fxn * mu[ijn]) * smdpn;
}

// processor-boundary faces with the halo cell as neighbour: facefluxmass2 in calcp (src-parallel/calcp :107-128,
// VARIANT 1), facefluxmass_piso in PISO / PIMPLE (src-parallel/PISO_multiple_correction.f90:181-202, VARIANT 2)
template <int VARIANT>
__global__ void __launch_bounds__(256)
k_calcp_proc_faces(geom_t g, flow_t f, double *__restrict__ apr, double *__restrict__ fmpro) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.npro) return;
  const int fc = g.pface0 + i;
  double cap, fl;
  facefluxmass<VARIANT>(g, f, g.owner[fc], g.n + i, g.xf[fc], g.yf[fc], g.zf[fc], g.arx[fc], g.ary[fc], g.arz[fc], g.fpro[i],
                        cap, fl);
  apr[i] = cap;
  fmpro[i] = fl;
}

__device__ __forceinline__ double fluxmc(const geom_t &g, const flow_t &f, int ijp, int ijn, int fc, double lambda) {
  const double xf = g.xf[fc], yf = g.yf[fc], zf = g.zf[fc], arx = g.arx[fc], ary = g.ary[fc], arz = g.arz[fc];
  const double fxn = lambda, fxp = 1.0 - fxn;
  const double xpn = g.xc[ijn] - g.xc[ijp];
  const double ypn = g.yc[ijn] - g.yc[ijp];
  const double zpn = g.zc[ijn] - g.zc[ijp];
  const double are = sqrt(arx * arx + ary * ary + arz * arz);
  const double nxx = arx / are, nyy = ary / are, nzz = arz / are;
  const double dppnnr = 1.0 / ((xpn * nxx) + (ypn * nyy) + (zpn * nzz));
  double xpp = xf - (xf - g.xc[ijp]) * nxx;
  double ypp = yf - (yf - g.yc[ijp]) * nyy;
  double zpp = zf - (zf - g.zc[ijp]) * nzz;
  double xep = xf - (xf - g.xc[ijn]) * nxx;
  double yep = yf - (yf - g.yc[ijn]) * nyy;
  double zep = zf - (zf - g.zc[ijn]) * nzz;
  xpp = xpp - g.xc[ijp]; ypp = ypp - g.yc[ijp]; zpp = zpp - g.zc[ijp];
  xep = xep - g.xc[ijn]; yep = yep - g.yc[ijn]; zep = zep - g.zc[ijn];
  const double rapr = (f.apu[ijp] * f.den[ijp] * g.vol[ijp] * fxp + f.apu[ijn] * f.den[ijn] * g.vol[ijn] * fxn);
  const double *dP = f.dP;
  return rapr * are *
         ((G3(dP, 0, ijn) * xep - G3(dP, 0, ijp) * xpp) + (G3(dP, 1, ijn) * yep - G3(dP, 1, ijp) * ypp) +
          (G3(dP, 2, ijn) * zep - G3(dP, 2, ijp) * zpp)) *
         dppnnr;
}

__global__ void __launch_bounds__(256)
k_fluxmc_faces(geom_t g, flow_t f, double *__restrict__ fmcor_out, double *__restrict__ flmass) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.F) return;
  const double fmcor = fluxmc(g, f, g.owner[i], g.neigh[i], i, g.facint[i]);
  fmcor_out[i] = fmcor;
  flmass[i] = flmass[i] + fmcor;
}

__global__ void k_fluxmc_proc_faces(geom_t g, flow_t f, double *__restrict__ fmcor_out, double *__restrict__ fmpro) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.npro) return;
  const int fc = g.pface0 + i;
  const double fmcor = fluxmc(g, f, g.owner[fc], g.n + i, fc, g.fpro[i]);
  fmcor_out[i] = fmcor;
  fmpro[i] = fmpro[i] + fmcor;
}

__global__ void k_laplacian_proc_faces(geom_t g, const double *__restrict__ mu, double *__restrict__ apr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.npro) return;
  const int fc = g.pface0 + i, ijp = g.owner[fc], ijn = g.n + i;
  const double arx = g.arx[fc], ary = g.ary[fc], arz = g.arz[fc];
  const double fxn = g.fpro[i], fxp = 1.0 - fxn;
  const double xpn = g.xc[ijn] - g.xc[ijp];
  const double ypn = g.yc[ijn] - g.yc[ijp];
  const double zpn = g.zc[ijn] - g.zc[ijp];
  const double smdpn = (arx * arx + ary * ary + arz * arz) / (arx * xpn + ary * ypn + arz * zpn);
  apr[i] = (fxp * mu[ijp] + fxn * mu[ijn]) * smdpn;
}

// fmpro(i) += apr(i) * (pp(halo) - pp(owner))   (src-parallel/calcp :200-208)
__global__ void k_fmpro_correct(geom_t g, const double *__restrict__ apr, const double *__restrict__ pp, double *fmpro) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.npro) return;
  fmpro[i] = fmpro[i] + apr[i] * (pp[g.n + i] - pp[g.owner[g.pface0 + i]]);
}

// ------------------------------------------------------------------------------------------
// Row gather: what the reference's face loop scatters into a(diag), a(off-diagonals) and su
// (calcp :56-75 + adjustMassFlow; laplacian :38-65, :140-154), one thread per cell.
// ------------------------------------------------------------------------------------------
enum { ROWS_CALCP = 0, ROWS_LAPLACIAN = 1, ROWS_SU_ONLY = 2 };

template <int KIND>
__global__ void __launch_bounds__(256)
k_rows_gather(geom_t g, c2f_t m, slots_t sl, const int *__restrict__ diag, const double *__restrict__ coef,
              const double *__restrict__ flux, const double *__restrict__ apr, const double *__restrict__ fpr,
              const double *__restrict__ fmi, const double *__restrict__ fmo,
              int mass_bc, const double *__restrict__ mu, const double *__restrict__ phi, double *__restrict__ a,
              double *__restrict__ su) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.n) return;
  double d = 0.0;
  double s = (KIND == ROWS_LAPLACIAN) ? su[c] : 0.0;
  const int qs = m.off[c], qe = m.off[c + 1];
  for (int q = qs; q < qe; ++q) {
    const int fe = m.face[q];
    const int f = fe & 0x7fffffff;
    if (f < g.F) {
      if (KIND != ROWS_SU_ONLY) {
        const double cf = coef[f];
        a[m.pos[q]] = cf;     // a(icell_jcell) = can / a(jcell_icell) = cap (cap == can)
        d = d - cf;
      }
      if (KIND != ROWS_LAPLACIAN) {
        const double fl = flux[f];
        s = (fe < 0) ? s + fl : s - fl;   // su(owner) -= flux ; su(neighbour) += flux
      }
    } else if (m.other[q] < g.n + g.npro) {   // processor face: coupling stays outside the CSR (apr)
      const int i = m.other[q] - g.n;
      if (KIND != ROWS_SU_ONLY) d = d - apr[i];
      if (KIND != ROWS_LAPLACIAN) s = s - fpr[i];
    } else if (KIND == ROWS_CALCP) {
      const int o = m.other[q];
      if (mass_bc) {
        if (o >= sl.slot[0] && o < sl.slot[0] + sl.count[0]) s = s - fmi[o - sl.slot[0]];       // adjustMassFlow.f90:33-38
        else if (o >= sl.slot[1] && o < sl.slot[1] + sl.count[1]) s = s - fmo[o - sl.slot[1]];  // :59-66
      }
    } else if (KIND == ROWS_LAPLACIAN) {
      const int o = m.other[q];
      if (o >= sl.slot[3] && o < sl.slot[3] + sl.count[3]) {  // wall: fvm_laplacian.f90:140-154
        const double ax = g.arx[f], ay = g.ary[f], az = g.arz[f];
        const double are = sqrt(ax * ax + ay * ay + az * az);
        const double dx = g.xc[c] - g.xf[f], dy = g.yc[c] - g.yf[f], dz = g.zc[c] - g.zf[f];
        const double dpw = sqrt(dx * dx + dy * dy + dz * dz);
        d = d - mu[c] * are / dpw;
        s = s + d * phi[o];
      }
    }
  }
  if (KIND != ROWS_SU_ONLY) a[diag[c]] = d;
  su[c] = s;
}

// ------------------------------------------------------------------------------------------
// adjustMassFlow / correctBoundaryConditionsVelocity: outlet extrapolation and scaling
// (adjustMassFlow.f90:41-66), symmetry projection (correctBoundaryConditionsVelocity.f90:55-75)
// ------------------------------------------------------------------------------------------
__global__ void k_outlet_extrapolate(geom_t g, slots_t sl, double *u, double *v, double *w,
                                     const double *__restrict__ den, double *fmo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sl.count[1]) return;
  const int f = sl.face[1] + i, ijp = g.owner[f], ijb = sl.slot[1] + i;
  u[ijb] = u[ijp]; v[ijb] = v[ijp]; w[ijb] = w[ijp];
  fmo[i] = den[ijp] * (u[ijb] * g.arx[f] + v[ijb] * g.ary[f] + w[ijb] * g.arz[f]);
}

// flowo = sum(fmo) in the reference's sequential order (one thread; outlet patches are O(boundary))
__global__ void k_outlet_sum(int nout, const double *__restrict__ fmo, fc_scalars *sc) {
  double flowo = 0.0;
  for (int i = 0; i < nout; ++i) flowo = flowo + fmo[i];
  sc->red[0] = flowo;
}
__global__ void k_outlet_factor(const fc_scalars *sc, double flomas, double small, double *fac) {
  *fac = flomas / (sc->red[0] + small);
}

__global__ void k_outlet_scale(slots_t sl, double *u, double *v, double *w, double *fmo,
                               const double *__restrict__ fac_p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sl.count[1]) return;
  const double fac = *fac_p;
  const int ijb = sl.slot[1] + i;
  fmo[i] = fmo[i] * fac;
  u[ijb] = u[ijb] * fac; v[ijb] = v[ijb] * fac; w[ijb] = w[ijb] * fac;
}

__global__ void k_symmetry_project(geom_t g, slots_t sl, double *u, double *v, double *w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sl.count[2]) return;
  const int f = sl.face[2] + i, ijp = g.owner[f], ijb = sl.slot[2] + i;
  const double ax = g.arx[f], ay = g.ary[f], az = g.arz[f];
  const double Unmag = u[ijp] * ax + v[ijp] * ay + w[ijp] * az;
  u[ijb] = u[ijp] - Unmag * ax;
  v[ijb] = v[ijp] - Unmag * ay;
  w[ijb] = w[ijp] - Unmag * az;
}

// ------------------------------------------------------------------------------------------
// bpres (bpres.f90:37-150)
// ------------------------------------------------------------------------------------------
__global__ void k_bpres(geom_t g, slots_t sl, double *p, const double *__restrict__ dPdxi, int istage, int total) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int b = 0, i = t;
  while (i >= sl.count[b]) { i -= sl.count[b]; ++b; }
  const int f = sl.face[b] + i, ijp = g.owner[f], ijb = sl.slot[b] + i;
  if (istage == 1) {
    p[ijb] = p[ijp];
  } else if (b == 3 || b == 4) {  // wall and prOutlet only (:103-131)
    const double xpb = g.xf[f] - g.xc[ijp], ypb = g.yf[f] - g.yc[ijp], zpb = g.zf[f] - g.zc[ijp];
    p[ijb] = p[ijp] + G3(dPdxi, 0, ijp) * xpb + G3(dPdxi, 1, ijp) * ypb + G3(dPdxi, 2, ijp) * zpb;
  }
}

// ------------------------------------------------------------------------------------------
// post-solve corrections (calcp :146-181)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_flux_correct(geom_t g, const double *__restrict__ coef, const double *__restrict__ pp, double *flmass) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.F) return;
  flmass[i] = flmass[i] + coef[i] * (pp[g.neigh[i]] - pp[g.owner[i]]);
}

__global__ void __launch_bounds__(256)
k_cell_correct(geom_t g, double *u, double *v, double *w, double *p, const double *__restrict__ pp,
               const double *__restrict__ dP, const double *__restrict__ apu, const double *__restrict__ apv,
               const double *__restrict__ apw, double urf, const double *__restrict__ ppref_p) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.n) return;
  const double ppref = *ppref_p;
  const double vol = g.vol[c];
  u[c] = u[c] - G3(dP, 0, c) * vol * apu[c];
  v[c] = v[c] - G3(dP, 1, c) * vol * apv[c];
  w[c] = w[c] - G3(dP, 2, c) * vol * apw[c];
  p[c] = p[c] + urf * (pp[c] - ppref);
}

__global__ void k_pick(const double *src, int idx, double *dst) { *dst = src[idx]; }
__global__ void k_mean(fc_scalars *sc, double gloCells, double *dst) { *dst = sc->red[0] / gloCells; }

// continuityErrors.h: res = net flux per cell -- including its flmass(ijp) (owner CELL id used
// as a face index, :18-19) -- then sum|res| and sum res.  Report only.
__global__ void __launch_bounds__(FC_RED_BLOCK)
k_continuity(geom_t g, c2f_t m, slots_t sl, const double *__restrict__ flmass, const double *__restrict__ fmpro,
             const double *__restrict__ fmi, const double *__restrict__ fmo, double *res, double *partials,
             fc_scalars *sc) {
  __shared__ double s_red[64];
  double a0 = 0.0, a1 = 0.0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < g.n; c += gridDim.x * blockDim.x) {
    double r = 0.0;
    for (int q = m.off[c]; q < m.off[c + 1]; ++q) {
      const int fe = m.face[q];
      const int f = fe & 0x7fffffff;
      const int o = m.other[q];
      if (f < g.F) {
        const int own = (fe < 0) ? o : c;
        const double fl = flmass[own < g.F ? own : g.F - 1];
        r = (fe < 0) ? r + fl : r - fl;
      } else if (o < g.n + g.npro) r = r - fmpro[o - g.n];
      else if (o >= sl.slot[0] && o < sl.slot[0] + sl.count[0]) r = r - fmi[o - sl.slot[0]];
      else if (o >= sl.slot[1] && o < sl.slot[1] + sl.count[1]) r = r - fmo[o - sl.slot[1]];
    }
    res[c] = r;
    a0 += fabs(r);
    a1 += r;
  }
  double v[2] = {a0, a1};
  if (fc_grid_sum<2>(v, partials, &sc->ticket[2], s_red)) {
    sc->red[0] = v[0];
    sc->red[1] = v[1];
  }
}

__global__ void __launch_bounds__(FC_RED_BLOCK)
k_sum(int n, const double *__restrict__ x, double *partials, fc_scalars *sc) {
  __shared__ double s_red[32];
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += x[i];
  double v[1] = {acc};
  if (fc_grid_sum<1>(v, partials, &sc->ticket[2], s_red)) sc->red[0] = v[0];
}

int grad_pass(fc_context *ctx, const double *phi, const double *dfo, double *df) {
  const int B = 256, G = fc_blocks(ctx->n, B);
  if (dfo) k_grad_pass<true><<<G, B, 0, ctx->stream>>>(geom_of(ctx), c2f_of(ctx), phi, dfo, df);
  else k_grad_pass<false><<<G, B, 0, ctx->stream>>>(geom_of(ctx), c2f_of(ctx), phi, nullptr, df);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

flow_t flow_of(fc_context *ctx) {
  return flow_t{ctx->field[FC_U], ctx->field[FC_V], ctx->field[FC_W], ctx->field[FC_P], ctx->field[FC_DEN],
                ctx->field[FC_DUDXI], ctx->field[FC_DVDXI], ctx->field[FC_DWDXI], ctx->field[FC_DPDXI],
                ctx->field[FC_APU], ctx->field[FC_APV], ctx->field[FC_APW]};
}

// `global`: adjustMassFlow sums flowo over the ranks (src-parallel/adjustMassFlow.f90:49),
// correctBoundaryConditionsVelocity does not
int outlet_extrapolate_and_scale(fc_context *ctx, double flomas, double small, bool global) {
  const slots_t sl = slots_of(ctx);
  const int nout = sl.count[1];
  const bool comm = global && ctx->nranks > 1;
  if (nout == 0 && !comm) return FC_OK;
  const int B = 256, G = fc_blocks(nout > 0 ? nout : 1, B);
  double *fac = &ctx->sc->aux[2];
  if (nout > 0) {
    k_outlet_extrapolate<<<G, B, 0, ctx->stream>>>(geom_of(ctx), sl, ctx->field[FC_U], ctx->field[FC_V],
                                                   ctx->field[FC_W], ctx->field[FC_DEN], ctx->field[FC_FMO]);
    FC_LAUNCH_CHECK();
  }
  k_outlet_sum<<<1, 1, 0, ctx->stream>>>(nout, ctx->field[FC_FMO], ctx->sc);
  FC_LAUNCH_CHECK();
  if (comm) FC_CHECK(fc_allreduce_scalars(ctx, ctx->sc->red, 1));
  if (nout == 0) return FC_OK;
  k_outlet_factor<<<1, 1, 0, ctx->stream>>>(ctx->sc, flomas, small, fac);
  FC_LAUNCH_CHECK();
  k_outlet_scale<<<G, B, 0, ctx->stream>>>(sl, ctx->field[FC_U], ctx->field[FC_V], ctx->field[FC_W],
                                           ctx->field[FC_FMO], fac);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

}  // namespace

static int need_mesh(fc_context *ctx, const char *who) {
  if (!ctx->has_mesh || !ctx->has_csr || !ctx->c2f_off)
    FC_FAIL(FC_ERR_ARG, std::string(who) + ": call fc_set_mesh and fc_create_csr first");
  return FC_OK;
}

// grad(phi,dPhidxi): src-parallel/gradients.f90:95-160 brackets the Gauss passes with exchange(phi)
// and the exchange of the three gradient components (one packed NCCL exchange here).  For
// nigrad > 1 the previous-pass gradient is exchanged too (the reference indexes a numCells-sized
// copy with halo cells there, i.e. out of bounds).
int fc_grad_gauss_dev(fc_context *ctx, double *phi, double *grad, int nigrad) {
  FC_CHECK(need_mesh(ctx, "fc_grad_gauss"));
  if (nigrad < 1) FC_FAIL(FC_ERR_ARG, "fc_grad_gauss: nigrad < 1");
  if (ctx->npro > 0) FC_CHECK(fc_halo_exchange(ctx, phi));
  for (int lc = 1; lc <= nigrad; ++lc) {
    if (lc == 1) FC_CHECK(grad_pass(ctx, phi, nullptr, grad));
    else {
      if (ctx->npro > 0) FC_CHECK(fc_halo_exchange3(ctx, grad));
      FC_CUDA(cudaMemcpyAsync(ctx->gtmp, grad, sizeof(double) * 3 * (size_t)ctx->NP, cudaMemcpyDeviceToDevice,
                              ctx->stream));
      FC_CHECK(grad_pass(ctx, phi, ctx->gtmp, grad));
    }
  }
  if (ctx->npro > 0) FC_CHECK(fc_halo_exchange3(ctx, grad));
  return FC_OK;
}

int fc_grad_gauss_corrected_dev(fc_context *ctx, double *phi, double *grad, int zero_seed) {
  FC_CHECK(need_mesh(ctx, "fc_grad_gauss_corrected"));
  if (ctx->npro > 0) FC_CHECK(fc_halo_exchange(ctx, phi));
  if (zero_seed) FC_CHECK(grad_pass(ctx, phi, nullptr, grad));  // seed 0: the correction terms vanish identically
  else {
    FC_CUDA(cudaMemcpyAsync(ctx->gtmp, grad, sizeof(double) * 3 * (size_t)ctx->NP, cudaMemcpyDeviceToDevice,
                            ctx->stream));
    FC_CHECK(grad_pass(ctx, phi, ctx->gtmp, grad));
  }
  if (ctx->npro > 0) FC_CHECK(fc_halo_exchange3(ctx, grad));
  return FC_OK;
}

int fc_bpres_dev(fc_context *ctx, double *p, const double *dPdxi, int istage) {
  FC_CHECK(need_mesh(ctx, "fc_bpres"));
  const slots_t sl = slots_of(ctx);
  int total = 0;
  for (int b = 0; b < 5; ++b) total += sl.count[b];
  if (total == 0) return FC_OK;
  k_bpres<<<fc_blocks(total, 256), 256, 0, ctx->stream>>>(geom_of(ctx), sl, p, dPdxi, istage, total);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

int fc_laplacian_dev(fc_context *ctx, double *mu, const double *phi) {
  FC_CHECK(need_mesh(ctx, "fc_laplacian"));
  const int B = 256;
  if (ctx->npro > 0) {  // src-parallel/fvm_laplacian.f90:36, :93-114
    FC_CHECK(fc_halo_exchange(ctx, mu));
    k_laplacian_proc_faces<<<fc_blocks(ctx->npro, B), B, 0, ctx->stream>>>(geom_of(ctx), mu, ctx->field[FC_APR]);
    FC_LAUNCH_CHECK();
  }
  if (ctx->csr_dup) FC_CUDA(cudaMemsetAsync(ctx->field[FC_A], 0, sizeof(double) * (size_t)ctx->nnz, ctx->stream));
  if (ctx->F > 0) {
    k_laplacian_faces<<<fc_blocks(ctx->F, B), B, 0, ctx->stream>>>(geom_of(ctx), mu, ctx->coef);
    FC_LAUNCH_CHECK();
  }
  k_rows_gather<ROWS_LAPLACIAN><<<fc_blocks(ctx->n, B), B, 0, ctx->stream>>>(
      geom_of(ctx), c2f_of(ctx), slots_of(ctx), ctx->diag, ctx->coef, nullptr, ctx->field[FC_APR], nullptr, nullptr,
      nullptr, 0, mu, phi, ctx->field[FC_A], ctx->field[FC_SU]);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

int fc_calcp_assemble_dev(fc_context *ctx, const fc_calcp_opts *o) {
  FC_CHECK(need_mesh(ctx, "fc_calcp_assemble"));
  const int B = 256;
  FC_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (ctx->csr_dup) FC_CUDA(cudaMemsetAsync(ctx->field[FC_A], 0, sizeof(double) * (size_t)ctx->nnz, ctx->stream));
  // grad(U), grad(V), grad(W)   (calcp :38-40)
  FC_CHECK(fc_grad_uvw_dev(ctx, o->nigrad, false));
  if (ctx->F > 0) {
    const int G = fc_blocks(ctx->F, B);
#define FC_FACES(V)                                                                                                   \
  do {                                                                                                                \
    if (ctx->tune_face_occ >= 4)                                                                                      \
      k_calcp_faces<V, 4><<<G, B, 0, ctx->stream>>>(geom_of(ctx), flow_of(ctx), ctx->coef, ctx->field[FC_FLMASS]);    \
    else if (ctx->tune_face_occ == 3)                                                                                 \
      k_calcp_faces<V, 3><<<G, B, 0, ctx->stream>>>(geom_of(ctx), flow_of(ctx), ctx->coef, ctx->field[FC_FLMASS]);    \
    else                                                                                                              \
      k_calcp_faces<V, 2><<<G, B, 0, ctx->stream>>>(geom_of(ctx), flow_of(ctx), ctx->coef, ctx->field[FC_FLMASS]);    \
  } while (0)
    if (o->flux_variant == 0)
      FC_FACES(0);
    else if (o->flux_variant == 1)
      FC_FACES(1);
    else
      FC_FACES(2);
#undef FC_FACES
    FC_LAUNCH_CHECK();
  }
  if (ctx->npro > 0) {
    if (o->flux_variant == 2)
      k_calcp_proc_faces<2><<<fc_blocks(ctx->npro, B), B, 0, ctx->stream>>>(geom_of(ctx), flow_of(ctx), ctx->field[FC_APR],
                                                                            ctx->field[FC_FMPRO]);
    else
      k_calcp_proc_faces<1><<<fc_blocks(ctx->npro, B), B, 0, ctx->stream>>>(geom_of(ctx), flow_of(ctx), ctx->field[FC_APR],
                                                                            ctx->field[FC_FMPRO]);
    FC_LAUNCH_CHECK();
  }
  if (!o->const_mflux) FC_CHECK(outlet_extrapolate_and_scale(ctx, o->flomas, o->sol.small, true));  // adjustMassFlow
  k_rows_gather<ROWS_CALCP><<<fc_blocks(ctx->n, B), B, 0, ctx->stream>>>(
      geom_of(ctx), c2f_of(ctx), slots_of(ctx), ctx->diag, ctx->coef, ctx->field[FC_FLMASS], ctx->field[FC_APR],
      ctx->field[FC_FMPRO], ctx->field[FC_FMI], ctx->field[FC_FMO], o->const_mflux ? 0 : 1, nullptr, nullptr,
      ctx->field[FC_A], ctx->field[FC_SU]);
  FC_LAUNCH_CHECK();
  FC_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
  return FC_OK;
}

// Post-solve part of one pressure corrector (calcp :132-223): boundary pressure + gradient of pp, reference value,
// flux / velocity / pressure correction, boundary velocities and, unless it is the last corrector, the
// non-orthogonal corrector source in su.  PP holds the solved correction.
int fc_calcp_correct_dev(fc_context *ctx, const fc_calcp_opts *o, int ipcorr) {
  FC_CHECK(need_mesh(ctx, "fc_calcp_correct"));
  if (o->npcor < 1 || o->npcor > 8) FC_FAIL(FC_ERR_ARG, "fc_calcp: npcor must be in 1..8");
  if (o->pRefCell < 1 || o->pRefCell > ctx->n) FC_FAIL(FC_ERR_ARG, "fc_calcp: pRefCell out of range");
  if (ipcorr < 1 || ipcorr > o->npcor) FC_FAIL(FC_ERR_ARG, "fc_calcp_correct: ipcorr must be in 1..npcor");
  const int B = 256, n = ctx->n;
  cudaStream_t st = ctx->stream;
  double *pp = ctx->field[FC_PP], *dP = ctx->field[FC_DPDXI];
  for (int istage = 1; istage <= o->nipgrad; ++istage) {                                    // :132-140
    FC_CHECK(fc_bpres_dev(ctx, pp, dP, istage));
    FC_CHECK(fc_grad_dev(ctx, pp, dP, o->nigrad));
  }
  if (o->lsq_flag) {                                                                        // :143
    FC_CHECK(fc_grad_gauss_corrected_dev(ctx, pp, dP, 1));
    FC_CHECK(fc_limit_gradient_dev(ctx, pp, dP));   // grad_scalar_field_w_option ends with the limiter (gradients.f90:240-255)
  }
  double *ppref = &ctx->sc->aux[3];
  if (ctx->nranks == 1) {
    k_pick<<<1, 1, 0, st>>>(pp, o->pRefCell - 1, ppref);                                    // :146
    FC_LAUNCH_CHECK();
  } else {  // src-parallel/calcp :175-177: ppref = global mean of pp
    k_sum<<<FC_RED_GRID, FC_RED_BLOCK, 0, st>>>(n, pp, ctx->partials, ctx->sc);
    FC_LAUNCH_CHECK();
    FC_CHECK(fc_allreduce_scalars(ctx, ctx->sc->red, 1));
    k_mean<<<1, 1, 0, st>>>(ctx->sc, (double)ctx->m.gloCells, ppref);
    FC_LAUNCH_CHECK();
  }
  if (ctx->F > 0) {
    k_flux_correct<<<fc_blocks(ctx->F, B), B, 0, st>>>(geom_of(ctx), ctx->coef, pp, ctx->field[FC_FLMASS]);
    FC_LAUNCH_CHECK();
  }
  if (ctx->npro > 0) {
    k_fmpro_correct<<<fc_blocks(ctx->npro, B), B, 0, st>>>(geom_of(ctx), ctx->field[FC_APR], pp,
                                                           ctx->field[FC_FMPRO]);
    FC_LAUNCH_CHECK();
  }
  k_cell_correct<<<fc_blocks(n, B), B, 0, st>>>(geom_of(ctx), ctx->field[FC_U], ctx->field[FC_V], ctx->field[FC_W],
                                                ctx->field[FC_P], pp, dP, ctx->field[FC_APU], ctx->field[FC_APV],
                                                ctx->field[FC_APW], o->urf_p, ppref);
  FC_LAUNCH_CHECK();
  // correctBoundaryConditionsVelocity (:184)
  FC_CHECK(outlet_extrapolate_and_scale(ctx, o->flomas, o->sol.small, false));
  const slots_t sl = slots_of(ctx);
  if (sl.count[2] > 0) {
    k_symmetry_project<<<fc_blocks(sl.count[2], B), B, 0, st>>>(geom_of(ctx), sl, ctx->field[FC_U],
                                                                ctx->field[FC_V], ctx->field[FC_W]);
    FC_LAUNCH_CHECK();
  }
  if (ipcorr != o->npcor) {  // non-orthogonal corrector source (:187-223)
    if (ctx->F > 0) {
      k_fluxmc_faces<<<fc_blocks(ctx->F, B), B, 0, st>>>(geom_of(ctx), flow_of(ctx), ctx->facev,
                                                         ctx->field[FC_FLMASS]);
      FC_LAUNCH_CHECK();
    }
    if (ctx->npro > 0) {
      k_fluxmc_proc_faces<<<fc_blocks(ctx->npro, B), B, 0, st>>>(geom_of(ctx), flow_of(ctx), ctx->facev + ctx->F,
                                                                 ctx->field[FC_FMPRO]);
      FC_LAUNCH_CHECK();
    }
    k_rows_gather<ROWS_SU_ONLY><<<fc_blocks(n, B), B, 0, st>>>(
        geom_of(ctx), c2f_of(ctx), sl, ctx->diag, ctx->coef, ctx->facev, nullptr, ctx->facev + ctx->F, nullptr,
        nullptr, 0, nullptr, nullptr, ctx->field[FC_A], ctx->field[FC_SU]);
    FC_LAUNCH_CHECK();
  }
  return FC_OK;
}

// End of calcp: halo of the corrected fields on several ranks (src-parallel/calcp :289-292) and continuityErrors.h
int fc_calcp_finish_dev(fc_context *ctx, fc_calcp_report *rep) {
  FC_CHECK(need_mesh(ctx, "fc_calcp_finish"));
  cudaStream_t st = ctx->stream;
  if (ctx->npro > 0)  // src-parallel/calcp :289-292
    for (int fld : {FC_U, FC_V, FC_W, FC_P}) FC_CHECK(fc_halo_exchange(ctx, ctx->field[fld]));
  k_continuity<<<FC_RED_GRID, FC_RED_BLOCK, 0, st>>>(geom_of(ctx), c2f_of(ctx), slots_of(ctx), ctx->field[FC_FLMASS],
                                                     ctx->field[FC_FMPRO], ctx->field[FC_FMI], ctx->field[FC_FMO],
                                                     ctx->field[FC_RES], ctx->partials, ctx->sc);
  FC_LAUNCH_CHECK();
  FC_CHECK(fc_allreduce_scalars(ctx, ctx->sc->red, 2));
  FC_CUDA(cudaMemcpyAsync(ctx->sc_host, ctx->sc, sizeof(fc_scalars), cudaMemcpyDeviceToHost, st));
  FC_CUDA(cudaStreamSynchronize(st));
  rep->sumLocalContErr = ctx->sc_host->red[0];
  rep->globalContErr = ctx->sc_host->red[1];
  return FC_OK;
}

int fc_calcp_dev(fc_context *ctx, const fc_calcp_opts *o, fc_calcp_report *rep) {
  FC_CHECK(need_mesh(ctx, "fc_calcp"));
  if (o->npcor < 1 || o->npcor > 8) FC_FAIL(FC_ERR_ARG, "fc_calcp: npcor must be in 1..8");
  if (o->pRefCell < 1 || o->pRefCell > ctx->n) FC_FAIL(FC_ERR_ARG, "fc_calcp: pRefCell out of range");
  cudaStream_t st = ctx->stream;
  FC_CHECK(fc_calcp_assemble_dev(ctx, o));
  double *pp = ctx->field[FC_PP];
  double solve_ms = 0.0;
  cudaEvent_t c0, c1;
  FC_CUDA(cudaEventCreate(&c0));
  FC_CUDA(cudaEventCreate(&c1));
  float corr_ms = 0.f;
  for (int ipcorr = 1; ipcorr <= o->npcor; ++ipcorr) {
    FC_CUDA(cudaMemsetAsync(pp, 0, sizeof(double) * (size_t)ctx->NT, st));                    // pp = 0 (:115)
    FC_CHECK(fc_solve_device(ctx, o->solver, pp, &o->sol, &rep->rep[ipcorr - 1], nullptr));   // :118-120
    solve_ms += ctx->tm.solve_ms;
    FC_CUDA(cudaEventRecord(c0, st));
    FC_CHECK(fc_calcp_correct_dev(ctx, o, ipcorr));
    FC_CUDA(cudaEventRecord(c1, st));
    FC_CUDA(cudaEventSynchronize(c1));
    float ms = 0.f;
    FC_CUDA(cudaEventElapsedTime(&ms, c0, c1));
    corr_ms += ms;
  }
  FC_CHECK(fc_calcp_finish_dev(ctx, rep));
  float ams = 0.f;
  FC_CUDA(cudaEventElapsedTime(&ams, ctx->ev[2], ctx->ev[3]));
  ctx->tm.assemble_ms = ams;
  ctx->tm.correct_ms = corr_ms;
  ctx->tm.solve_ms = solve_ms;
  cudaEventDestroy(c0);
  cudaEventDestroy(c1);
  return FC_OK;
}

// ------------------------------------------------------------------------------------------
// PISO / PIMPLE pressure equation (PISO_multiple_correction.f90, PIMPLE_multiple_correction.f90,
// get_rAU_x_UEqnH.f90).  Bodies: fc_piso_body.cuh.
// ------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) k_hbya_rows(fcm_geom g, fcm_c2f m, fcp_hbya k) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < g.n) fcp_hbya_row(g, m, k, c);
}
__global__ void __launch_bounds__(256)
k_hbya_scale(int n, const double *apu, const double *apv, const double *apw, const double *su, const double *sv,
             const double *sw, double *u, double *v, double *w) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) fcp_hbya_scale(c, apu, apv, apw, su, sv, sw, u, v, w);
}
__global__ void k_pin_row(const int *ioffset, const int *diag, double *a, double *su, const double *src, int pref) {
  fcp_pin_row(ioffset, diag, a, su, src, pref);
}
__global__ void __launch_bounds__(256)
k_flux_correct_matrix(fcm_geom g, const int *icj, const double *a, const double *pp, double *flmass) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.F) fcp_flux_correct(g, icj, a, pp, flmass, i);
}
__global__ void __launch_bounds__(256)
k_piso_velocity_correct(fcm_geom g, const double *apu, const double *apv, const double *apw, const double *dP,
                        double *u, double *v, double *w) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < g.n) fcp_velocity_correct(g, apu, apv, apw, dP, u, v, w, c);
}
__global__ void __launch_bounds__(256) k_relax_p(int n, double urf, const double *pp, double *p) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) fcp_relax_p(urf, pp, p, c);
}

int piso_continuity(fc_context *ctx, fc_piso_report *rep) {   // continuityErrors.h
  k_continuity<<<FC_RED_GRID, FC_RED_BLOCK, 0, ctx->stream>>>(geom_of(ctx), c2f_of(ctx), slots_of(ctx),
                                                             ctx->field[FC_FLMASS], ctx->field[FC_FMPRO],
                                                             ctx->field[FC_FMI], ctx->field[FC_FMO], ctx->field[FC_RES],
                                                             ctx->partials, ctx->sc);
  FC_LAUNCH_CHECK();
  FC_CHECK(fc_allreduce_scalars(ctx, ctx->sc->red, 2));   // global_sum of src-parallel's continuityErrors.h (no-op on one rank)
  FC_CUDA(cudaMemcpyAsync(ctx->sc_host, ctx->sc, sizeof(fc_scalars), cudaMemcpyDeviceToHost, ctx->stream));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  rep->sumLocalContErr = ctx->sc_host->red[0];
  rep->globalContErr = ctx->sc_host->red[1];
  return FC_OK;
}

// p(inp) = urf(ip)*pp(inp) + (1-urf(ip))*p(inp)   (src-parallel/PIMPLE_multiple_correction.f90:92)
__global__ void __launch_bounds__(256) k_relax_p_par(int n, double urf, const double *pp, double *p) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) p[c] = urf * pp[c] + (1.0 - urf) * p[c];
}

}  // namespace

int fc_piso_dev(fc_context *ctx, const fc_piso_opts *o, fc_piso_report *rep) {
  FC_CHECK(need_mesh(ctx, "fc_piso"));
  // Several ranks = src-parallel/PISO_multiple_correction.f90 / PIMPLE_multiple_correction.f90: processor faces with
  // facefluxmass_piso, the reference pressure pinned on rank 0 only (iPrefProcess = 0, read_input.f90:316; su = p(pRefCell)
  // in both variants), flux correction and continuity report once after the npcor loop, PIMPLE's relaxation written as
  // urf*pp + (1-urf)*p, u, v, w, p exchanged once at the end, get_rAU_x_UEqnH with the processor terms of the parallel
  // routine (fc_piso_body.cuh).
  const bool par = ctx->nranks > 1;
  if (ctx->npro > 0 && !ctx->comm) FC_FAIL(FC_ERR_ARG, "fc_piso: processor faces without fc_comm_init");
  if (o->ncorr < 1 || o->npcor < 1 || o->nigrad < 1 || o->nipgrad < 0) FC_FAIL(FC_ERR_ARG, "fc_piso: bad corrector counts");
  if ((!par || ctx->rank == 0) && (o->pRefCell < 1 || o->pRefCell > ctx->n))
    FC_FAIL(FC_ERR_ARG, "fc_piso: pRefCell out of range (with several ranks it is a cell of rank 0)");
  if ((o->bdf || o->cn) && !(o->timestep > 0.0)) FC_FAIL(FC_ERR_ARG, "fc_piso: timestep must be > 0");
  FC_CHECK(fc_momentum_fields(ctx));
  if (!ctx->hcoef) FC_CHECK(fc_dev_alloc(ctx, &ctx->hcoef, (size_t)ctx->nnz + 2));
  const int B = 256, n = ctx->n;
  cudaStream_t st = ctx->stream;
  double **fl = ctx->field;
  const fcm_geom g = fcm_geom_of(ctx);
  const fcm_c2f m = fcm_c2f_of(ctx);
  // h = a   (PISO :84): the momentum matrix calcuvw left behind
  FC_CUDA(cudaMemcpyAsync(ctx->hcoef, fl[FC_A], sizeof(double) * (size_t)ctx->nnz, cudaMemcpyDeviceToDevice, st));
  fc_calcp_opts co{};
  co.npcor = 1; co.nigrad = o->nigrad; co.nipgrad = o->nipgrad; co.pRefCell = o->pRefCell;
  co.flux_variant = 2;   // facefluxmass_piso
  co.const_mflux = o->const_mflux; co.flomas = o->flomas; co.sol = o->sol;
  fc_solver_opts so = o->sol;
  if (par) { co.sol.parallel = 1; so.parallel = 1; }   // the parallel iccg (`+small`, rank-local DIC, exchange(pp) at exit)
  rep->nsolves = 0;
  rep->sumLocalContErr = rep->globalContErr = 0.0;
  double solve_ms = 0.0;
  for (int icorr = 1; icorr <= o->ncorr; ++icorr) {
    const fcp_hbya hk{ctx->ioffset, ctx->diag, ctx->hcoef, fl[FC_U], fl[FC_V], fl[FC_W], fl[FC_UO], fl[FC_VO],
                      fl[FC_WO], fl[FC_UOO], fl[FC_VOO], fl[FC_WOO], fl[FC_T], fl[FC_DEN], fl[FC_SU], fl[FC_SV],
                      fl[FC_SW], o->bdf, o->btime, o->timestep, o->cn, o->lbuoy, o->boussinesq, o->beta, o->tref,
                      o->densit, o->gravx, o->gravy, o->gravz, fl[FC_APR], ctx->npro};
    k_hbya_rows<<<fc_blocks(n, B), B, 0, st>>>(g, m, hk);                                     // get_rAU_x_UEqnH
    FC_LAUNCH_CHECK();
    k_hbya_scale<<<fc_blocks(n, B), B, 0, st>>>(n, fl[FC_APU], fl[FC_APV], fl[FC_APW], fl[FC_SU], fl[FC_SV],
                                                fl[FC_SW], fl[FC_U], fl[FC_V], fl[FC_W]);
    FC_LAUNCH_CHECK();
    // grad(U,V,W); a = 0; su = 0; facefluxmass_piso face loop; adjustMassFlow   (PISO :104-181)
    FC_CHECK(fc_calcp_assemble_dev(ctx, &co));
    if (!par || ctx->rank == 0) {
      k_pin_row<<<1, 1, 0, st>>>(ctx->ioffset, ctx->diag, fl[FC_A], fl[FC_SU], (o->pimple && !par) ? fl[FC_PP] : fl[FC_P],
                                 o->pRefCell - 1);                                            // :188-192
      FC_LAUNCH_CHECK();
    }
    for (int ipcorr = 1; ipcorr <= o->npcor; ++ipcorr) {
      fc_solver_report *r = &rep->rep[rep->nsolves < 16 ? rep->nsolves : 15];
      FC_CHECK(fc_solve_device(ctx, FC_ICCG, fl[FC_PP], &so, r, nullptr));                    // call iccg(pp,ip)
      solve_ms += ctx->tm.solve_ms;
      rep->nsolves++;
      if (!o->pimple && !par) {
        if (ipcorr == o->npcor && ctx->F > 0) {
          k_flux_correct_matrix<<<fc_blocks(ctx->F, B), B, 0, st>>>(g, ctx->icj, fl[FC_A], fl[FC_PP], fl[FC_FLMASS]);
          FC_LAUNCH_CHECK();
        }
        FC_CHECK(piso_continuity(ctx, rep));
      }
    }
    if (o->pimple || par) {   // one flux correction and one continuity report after the npcor loop
      if (ctx->F > 0) {
        k_flux_correct_matrix<<<fc_blocks(ctx->F, B), B, 0, st>>>(g, ctx->icj, fl[FC_A], fl[FC_PP], fl[FC_FLMASS]);
        FC_LAUNCH_CHECK();
      }
      if (ctx->npro > 0) {    // fmpro(i) += apr(i)*(pp(halo) - pp(owner)); iccg left pp's halo current
        k_fmpro_correct<<<fc_blocks(ctx->npro, B), B, 0, st>>>(geom_of(ctx), fl[FC_APR], fl[FC_PP], fl[FC_FMPRO]);
        FC_LAUNCH_CHECK();
      }
      FC_CHECK(piso_continuity(ctx, rep));
    }
    if (o->pimple) {
      if (par) k_relax_p_par<<<fc_blocks(n, B), B, 0, st>>>(n, o->urf_p, fl[FC_PP], fl[FC_P]);
      else k_relax_p<<<fc_blocks(n, B), B, 0, st>>>(n, o->urf_p, fl[FC_PP], fl[FC_P]);
      FC_LAUNCH_CHECK();
    } else {
      FC_CUDA(cudaMemcpyAsync(fl[FC_P], fl[FC_PP], sizeof(double) * (size_t)ctx->NT, cudaMemcpyDeviceToDevice, st));  // p = pp
    }
    for (int istage = 1; istage <= o->nipgrad; ++istage) {
      FC_CHECK(fc_bpres_dev(ctx, fl[FC_P], fl[FC_DPDXI], istage));
      FC_CHECK(fc_grad_dev(ctx, fl[FC_P], fl[FC_DPDXI], o->nigrad));
    }
    k_piso_velocity_correct<<<fc_blocks(n, B), B, 0, st>>>(g, fl[FC_APU], fl[FC_APV], fl[FC_APW], fl[FC_DPDXI],
                                                           fl[FC_U], fl[FC_V], fl[FC_W]);
    FC_LAUNCH_CHECK();
    // correctBoundaryConditionsVelocity
    FC_CHECK(outlet_extrapolate_and_scale(ctx, o->flomas, o->sol.small, false));
    const slots_t sl = slots_of(ctx);
    if (sl.count[2] > 0) {
      k_symmetry_project<<<fc_blocks(sl.count[2], B), B, 0, st>>>(geom_of(ctx), sl, fl[FC_U], fl[FC_V], fl[FC_W]);
      FC_LAUNCH_CHECK();
    }
  }
  if (ctx->npro > 0)   // src-parallel/PISO_multiple_correction.f90:383-386
    for (int fld : {FC_U, FC_V, FC_W, FC_P}) FC_CHECK(fc_halo_exchange(ctx, fl[fld]));
  FC_CUDA(cudaStreamSynchronize(st));
  ctx->tm.solve_ms = solve_ms;
  return FC_OK;
}
