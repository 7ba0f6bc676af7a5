// See fcapp_host.hpp.  Bodies = the C-ABI calls a Fortran maintainer would put behind the same
// subroutine names (fortran/fcapp_shim.f90).
#include "fcapp_host.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace fcapp {

namespace geometry {
int numCells, numInnerFaces, numFaces, numBoundaryFaces, numTotal, nnz;
int ninl, nout, nsym, nwal, npru, noc;
int iInletFacesStart, iOutletFacesStart, iSymmetryFacesStart, iWallFacesStart, iPressOutletFacesStart, iOCFacesStart;
std::vector<int> owner, neighbour;
std::vector<dp> xc, yc, zc, vol, arx, ary, arz, xf, yf, zf, facint;
}  // namespace geometry
namespace sparse_matrix {
std::vector<int> ioffset, ja, diag, icell_jcell_csr_value_index, jcell_icell_csr_value_index;
std::vector<dp> a, su, sv, sw, spu, spv, sp, res, apu, apv, apw;
}  // namespace sparse_matrix
namespace parameters {
dp small = (dp)1e-20f;
dp sor[nphi + 1], urf[nphi + 1], resor[nphi + 1];
int nsw[nphi + 1];
int npcor = 1, nigrad = 1, nipgrad = 2, pRefCell = 1, ncorr = 1;
bool const_mflux = false, ltest = false, lstsq_qr = false, lstsq_dm = false;
dp flomas = 0.0;
dp densit = 1.0, viscos = 0.01, gds[nphi + 1];
bool bdf = false, cn = false, lturb = false, lbuoy = false, boussinesq = true, lcal[nphi + 1];
dp btime = 0.0, timestep = 1e20, gradPcmf = 0.0, beta = 0.0, tref = 0.0, gravx = 0.0, gravy = 0.0, gravz = 0.0;
std::string convective_scheme = "muscl-f";
}  // namespace parameters
namespace variables {
std::vector<dp> u, v, w, p, pp, den, flmass, fmi, fmo;
std::vector<dp> vis, uo, vo, wo, uoo, voo, woo, t;
std::vector<dp> dUdxi, dVdxi, dWdxi, dPdxi;
dp sumLocalContErr = 0, globalContErr = 0, cumulativeContErr = 0;
}  // namespace variables
namespace title_mod {
const char *chvarSolver[parameters::nphi + 1] = {"", "U", "V", "W", "p", "k", "epsilon", "Energy", "Visc", "VisT", "Conc"};
}

static fc_context *ctx = nullptr;

// the reference has no status codes: a failing step stops the program
static void check(int status, const char *who) {
  if (status != FC_OK) {
    std::fprintf(stderr, "  libfcapp_cuda: %s failed with status %d: %s\n", who, status, fc_last_error(ctx));
    std::exit(1);
  }
}

static fc_solver_opts solver_opts(int ifi) {
  fc_solver_opts o;
  o.sor = parameters::sor[ifi];
  o.nsw = parameters::nsw[ifi];
  o.small = parameters::small;
  o.tol = (dp)1e-13f;  // dpcg.f90:37
  o.parallel = 0;
  return o;
}

void mesh_geometry_box(int nx, int ny, int nz, dp lx, dp ly, dp lz, const char *kinds[6]) {
  using namespace geometry;
  const dp dx = lx / nx, dy = ly / ny, dz = lz / nz;
  numCells = nx * ny * nz;
  xc.assign(numCells, 0); yc.assign(numCells, 0); zc.assign(numCells, 0); vol.assign(numCells, dx * dy * dz);
  auto id = [&](int i, int j, int k) { return i + nx * (j + ny * k); };
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int c = id(i, j, k);
        xc[c] = (i + 0.5) * dx; yc[c] = (j + 0.5) * dy; zc[c] = (k + 0.5) * dz;
      }
  owner.clear(); neighbour.clear();
  arx.clear(); ary.clear(); arz.clear(); xf.clear(); yf.clear(); zf.clear(); facint.clear();
  auto face = [&](int own, dp sx, dp sy, dp sz, dp fx, dp fy, dp fz) {
    owner.push_back(own + 1);
    arx.push_back(sx); ary.push_back(sy); arz.push_back(sz); xf.push_back(fx); yf.push_back(fy); zf.push_back(fz);
  };
  // inner faces: by owner, then +x, +y, +z neighbour (OpenFOAM upper-triangular order)
  for (int c = 0; c < numCells; ++c) {
    const int i = c % nx, j = (c / nx) % ny, k = c / (nx * ny);
    if (i < nx - 1) { face(c, dy * dz, 0, 0, xc[c] + 0.5 * dx, yc[c], zc[c]); neighbour.push_back(c + 1 + 1); }
    if (j < ny - 1) { face(c, 0, dx * dz, 0, xc[c], yc[c] + 0.5 * dy, zc[c]); neighbour.push_back(c + nx + 1); }
    if (k < nz - 1) { face(c, 0, 0, dx * dy, xc[c], yc[c], zc[c] + 0.5 * dz); neighbour.push_back(c + nx * ny + 1); }
  }
  numInnerFaces = (int)neighbour.size();
  facint.assign(numInnerFaces, 0.0);
  for (int f = 0; f < numInnerFaces; ++f) {  // mesh_geometry_and_topology.f90:1040-1062
    const int P = owner[f] - 1, N = neighbour[f] - 1;
    const dp dpn = std::sqrt((xc[N] - xc[P]) * (xc[N] - xc[P]) + (yc[N] - yc[P]) * (yc[N] - yc[P]) +
                             (zc[N] - zc[P]) * (zc[N] - zc[P]));
    const dp djn = std::sqrt((xf[f] - xc[P]) * (xf[f] - xc[P]) + (yf[f] - yc[P]) * (yf[f] - yc[P]) +
                             (zf[f] - zc[P]) * (zf[f] - zc[P]));
    facint[f] = djn / dpn;
  }
  // boundary patches x-,x+,y-,y+,z-,z+ regrouped so that equal kinds are contiguous
  const char *order[5] = {"inlet", "outlet", "symmetry", "wall", "prOutlet"};
  int *cnt[5] = {&ninl, &nout, &nsym, &nwal, &npru};
  int *fst[5] = {&iInletFacesStart, &iOutletFacesStart, &iSymmetryFacesStart, &iWallFacesStart, &iPressOutletFacesStart};
  for (int b = 0; b < 5; ++b) { *cnt[b] = 0; *fst[b] = 0; }
  std::vector<std::string> seen;
  for (int p = 0; p < 6; ++p) {
    bool have = false;
    for (auto &s : seen) have = have || s == kinds[p];
    if (!have) seen.push_back(kinds[p]);
  }
  for (auto &kind : seen) {
    int b = -1;
    for (int q = 0; q < 5; ++q)
      if (kind == order[q]) b = q;
    if (b < 0) { std::fprintf(stderr, "unknown boundary kind %s\n", kind.c_str()); std::exit(1); }
    *fst[b] = (int)owner.size();
    for (int p = 0; p < 6; ++p) {
      if (kind != kinds[p]) continue;
      const int ax = p / 2;
      const dp sg = (p % 2) ? 1.0 : -1.0;
      for (int c = 0; c < numCells; ++c) {
        const int i = c % nx, j = (c / nx) % ny, k = c / (nx * ny);
        const int pos[3] = {i, j, k}, lim[3] = {nx, ny, nz};
        if (pos[ax] != ((p % 2) ? lim[ax] - 1 : 0)) continue;
        const dp area[3] = {dy * dz, dx * dz, dx * dy}, half[3] = {0.5 * dx, 0.5 * dy, 0.5 * dz};
        dp s[3] = {0, 0, 0}, fcen[3] = {xc[c], yc[c], zc[c]};
        s[ax] = sg * area[ax];
        fcen[ax] += sg * half[ax];
        face(c, s[0], s[1], s[2], fcen[0], fcen[1], fcen[2]);
        ++*cnt[b];
      }
    }
  }
  noc = 0; iOCFacesStart = 0;
  numFaces = (int)owner.size();
  numBoundaryFaces = numFaces - numInnerFaces;
  numTotal = numCells + numBoundaryFaces;
  nnz = numCells + 2 * numInnerFaces;  // mesh_geometry_and_topology.f90:580
}

void fcapp_init(int device) {
  using namespace geometry;
  check(fc_create(device, &ctx), "fc_create");
  fc_mesh_desc m{};
  m.numCells = numCells; m.numInnerFaces = numInnerFaces; m.numFaces = numFaces; m.numTotal = numTotal;
  m.ninl = ninl; m.nout = nout; m.nsym = nsym; m.nwal = nwal; m.npru = npru; m.noc = noc;
  m.iInletFacesStart = iInletFacesStart; m.iOutletFacesStart = iOutletFacesStart;
  m.iSymmetryFacesStart = iSymmetryFacesStart; m.iWallFacesStart = iWallFacesStart;
  m.iPressOutletFacesStart = iPressOutletFacesStart; m.iOCFacesStart = iOCFacesStart;
  m.owner = owner.data(); m.neighbour = neighbour.data();
  m.xc = xc.data(); m.yc = yc.data(); m.zc = zc.data(); m.vol = vol.data();
  m.arx = arx.data(); m.ary = ary.data(); m.arz = arz.data(); m.xf = xf.data(); m.yf = yf.data(); m.zf = zf.data();
  m.facint = facint.data();
  m.gloCells = numCells;
  check(fc_set_mesh(ctx, &m), "fc_set_mesh");
}

void fcapp_finalize() {
  fc_destroy(ctx);
  ctx = nullptr;
}

void allocate_arrays() {
  using namespace geometry;
  using namespace variables;
  for (auto *f : {&u, &v, &w, &p, &pp, &uo, &vo, &wo, &uoo, &voo, &woo, &t}) f->assign(numTotal, 0.0);
  vis.assign(numTotal, parameters::viscos);
  den.assign(numTotal, 1.0);
  flmass.assign(numFaces, 0.0);
  fmi.assign(ninl > 0 ? ninl : 1, 0.0);
  fmo.assign(nout > 0 ? nout : 1, 0.0);
  for (auto *g : {&dUdxi, &dVdxi, &dWdxi, &dPdxi}) g->assign(3 * (size_t)numCells, 0.0);
}

void create_CSR_matrix_from_mesh_data() {
  using namespace geometry;
  using namespace sparse_matrix;
  ioffset.assign(numCells + 1, 0); ja.assign(nnz, 0); diag.assign(numCells, 0); a.assign(nnz, 0.0);
  for (auto *x : {&su, &sv, &sw, &spu, &spv, &sp, &res, &apu, &apv, &apw}) x->assign(numCells, 0.0);
  icell_jcell_csr_value_index.assign(numInnerFaces, 0);
  jcell_icell_csr_value_index.assign(numInnerFaces, 0);
  check(fc_create_csr(ctx, ioffset.data(), ja.data(), diag.data(), icell_jcell_csr_value_index.data(),
                      jcell_icell_csr_value_index.data()), "fc_create_csr");
}

void laplacian(const dp *mu, const dp *phi) {
  using namespace geometry;
  using namespace sparse_matrix;
  check(fc_upload(ctx, FC_APU, mu, numCells), "upload mu");
  check(fc_upload(ctx, FC_SCRATCH_T, phi, numTotal), "upload phi");
  check(fc_upload(ctx, FC_SU, su.data(), numCells), "upload su");
  check(fc_laplacian(ctx, FC_APU, FC_SCRATCH_T), "fc_laplacian");
  check(fc_download(ctx, FC_A, a.data(), nnz), "download a");
  check(fc_download(ctx, FC_SU, su.data(), numCells), "download su");
}

void grad(const dp *phi, dp *dPhidxi) {
  using namespace geometry;
  check(fc_upload(ctx, FC_SCRATCH_T, phi, numTotal), "upload phi");
  check(fc_grad_gauss(ctx, FC_SCRATCH_T, FC_DPDXI, parameters::nigrad), "fc_grad_gauss");
  check(fc_download(ctx, FC_DPDXI, dPhidxi, 3 * (size_t)numCells), "download gradient");
}

static void solve(int solver, const char *label, dp *fi, int ifi) {
  using namespace sparse_matrix;
  fc_solver_opts o = solver_opts(ifi);
  fc_solver_report rep;
  check(fc_solve_host(ctx, solver, a.data(), su.data(), fi, res.data(), &o, &rep), label);
  if (rep.iters > 0) parameters::resor[ifi] = rep.res0;  // dpcg.f90:139
  // '(3a,1PE10.3,a,1PE10.3,a,I0)' -- the format examples/*/plotResiduals parse
  std::printf("%s  Solving for %s, Initial residual = %10.3E, Final residual = %10.3E, No Iterations %d\n", label,
              title_mod::chvarSolver[ifi], rep.res0, rep.resl, rep.iters);
}

void dpcg(dp *fi, int ifi) { solve(FC_DPCG, "PCG(Jacobi):", fi, ifi); }
void iccg(dp *fi, int ifi) { solve(FC_ICCG, "  PCG(IC0):", fi, ifi); }
void bicgstab(dp *fi, int ifi) { solve(FC_BICGSTAB, "  BiCGStab(ILU(0)):", fi, ifi); }

void calcp() {
  using namespace geometry;
  using namespace parameters;
  using namespace variables;
  using namespace sparse_matrix;
  fc_calcp_opts o{};
  o.npcor = npcor; o.nigrad = nigrad; o.nipgrad = nipgrad; o.pRefCell = pRefCell;
  o.urf_p = urf[ip]; o.solver = FC_ICCG;  // calcp :119
  o.const_mflux = const_mflux; o.flomas = flomas;
  o.lsq_flag = lstsq_qr || lstsq_dm; o.flux_variant = 0;
  o.sol = solver_opts(ip);
  fc_calcp_report rep;
  check(fc_upload(ctx, FC_DEN, den.data(), numTotal), "upload den");
  if (ninl > 0) check(fc_upload(ctx, FC_FMI, fmi.data(), ninl), "upload fmi");
  check(fc_upload(ctx, FC_DPDXI, dPdxi.data(), 3 * (size_t)numCells), "upload dPdxi");
  check(fc_calcp_host(ctx, &o, u.data(), v.data(), w.data(), p.data(), pp.data(), apu.data(), apv.data(), apw.data(),
                      flmass.data(), &rep), "fc_calcp_host");
  check(fc_download(ctx, FC_DPDXI, dPdxi.data(), 3 * (size_t)numCells), "download dPdxi");
  check(fc_download(ctx, FC_SU, su.data(), numCells), "download su");
  for (int k = 0; k < npcor; ++k) {
    if (rep.rep[k].iters > 0) resor[ip] = rep.rep[k].res0;
    std::printf("  PCG(IC0):  Solving for %s, Initial residual = %10.3E, Final residual = %10.3E, No Iterations %d\n",
                title_mod::chvarSolver[ip], rep.rep[k].res0, rep.rep[k].resl, rep.rep[k].iters);
  }
  sumLocalContErr = rep.sumLocalContErr;  // continuityErrors.h
  globalContErr = rep.globalContErr;
  cumulativeContErr += globalContErr;
  std::printf("  time step continuity errors : sum local = %10.3E, global = %10.3E, cumulative = %10.3E\n",
              sumLocalContErr, globalContErr, cumulativeContErr);
}

// `call calcuvw` (src/calcuvw.f90:3-557), laminar: the momentum predictor on the device; every module array the
// remaining host routines read is brought back (same transfers as fortran/fcapp_shim.f90)
void calcuvw() {
  using namespace geometry;
  using namespace parameters;
  using namespace variables;
  using namespace sparse_matrix;
  if (lturb) { std::fprintf(stderr, "calcuvw: turbulent stresses (calcstress) are not on the GPU path\n"); std::exit(1); }
  fc_calcuvw_opts o{};
  o.nigrad = nigrad; o.nipgrad = nipgrad;
  // read_input.f90:97-133 -> face_value (interpolation.f90:36-57)
  static const struct { const char *name; int scheme, limiter; } table[] = {
      {"central", 0, 7}, {"cds-corrected", 1, 7}, {"central-f", 2, 7}, {"linear-f", 3, 7}, {"muscl-f", 4, 7},
      {"smart", 5, 0}, {"avl-smart", 5, 1}, {"muscl", 5, 2}, {"umist", 5, 3}, {"koren", 5, 4}, {"charm", 5, 5},
      {"ospre", 5, 6}, {"linear", 5, 7}};
  // an unknown name only renames the scheme ('Convective scheme not chosen, assigning default muscl scheme',
  // read_input.f90:123-126) without setting a flag, so face_value falls through to face_value_muscl
  o.scheme = 4; o.limiter = 7;
  for (auto &e : table)
    if (convective_scheme == e.name) { o.scheme = e.scheme; o.limiter = e.limiter; }
  o.gds = gds[iu];
  for (int k = 0; k < 3; ++k) { o.urf[k] = urf[iu + k]; o.sor[k] = sor[iu + k]; o.nsw[k] = nsw[iu + k]; }
  o.bdf = bdf; o.btime = btime; o.timestep = timestep; o.cn = cn;
  o.const_mflux = const_mflux; o.gradPcmf = gradPcmf;
  o.lbuoy = lcal[ien] && lbuoy; o.boussinesq = boussinesq;
  o.beta = beta; o.tref = tref; o.densit = densit; o.gravx = gravx; o.gravy = gravy; o.gravz = gravz;
  o.viscos = viscos;
  o.sol = solver_opts(iu);
  check(fc_upload(ctx, FC_DEN, den.data(), numTotal), "upload den");
  if (ninl > 0) check(fc_upload(ctx, FC_FMI, fmi.data(), ninl), "upload fmi");
  if (nout > 0) check(fc_upload(ctx, FC_FMO, fmo.data(), nout), "upload fmo");
  // `a` is NOT uploaded: the U row sums read the stale diagonal of the previous solve (:423), and every solve of
  // this build runs on the device, so the device copy of `a` is the current one by construction
  if (bdf || cn) {
    const struct { int f; std::vector<dp> *h; } old[] = {{FC_UO, &uo}, {FC_VO, &vo}, {FC_WO, &wo},
                                                         {FC_UOO, &uoo}, {FC_VOO, &voo}, {FC_WOO, &woo}};
    for (auto &e : old) check(fc_upload(ctx, e.f, e.h->data(), numTotal), "upload old time level");
  }
  if (o.lbuoy) check(fc_upload(ctx, FC_T, t.data(), numTotal), "upload t");
  fc_calcuvw_report rep;
  check(fc_calcuvw_host(ctx, &o, u.data(), v.data(), w.data(), p.data(), vis.data(), flmass.data(), apu.data(),
                        apv.data(), apw.data(), &rep), "fc_calcuvw_host");
  const struct { int f; std::vector<dp> *h; } grads[] = {{FC_DUDXI, &dUdxi}, {FC_DVDXI, &dVdxi}, {FC_DWDXI, &dWdxi},
                                                         {FC_DPDXI, &dPdxi}};
  for (auto &e : grads) check(fc_download(ctx, e.f, e.h->data(), 3 * (size_t)numCells), "download gradient");
  check(fc_download(ctx, FC_A, a.data(), nnz), "download a");
  for (int k = 0; k < 3; ++k) {
    if (rep.rep[k].iters > 0) resor[iu + k] = rep.rep[k].res0;
    std::printf("  BiCGStab(ILU(0)):  Solving for %s, Initial residual = %10.3E, Final residual = %10.3E, No Iterations %d\n",
                title_mod::chvarSolver[iu + k], rep.rep[k].res0, rep.rep[k].resl, rep.rep[k].iters);
  }
}

// PISO_multiple_correction.f90:2 / PIMPLE_multiple_correction.f90:2 -- directly after calcuvw: the device still holds
// the momentum matrix (backed up as h = a), ap*, u, v, w, the old time levels and flmass
static void piso(bool pimple) {
  using namespace geometry;
  using namespace parameters;
  using namespace variables;
  fc_piso_opts o{};
  o.ncorr = ncorr; o.npcor = npcor; o.nigrad = nigrad; o.nipgrad = nipgrad; o.pRefCell = pRefCell;
  o.pimple = pimple; o.urf_p = urf[ip];
  o.const_mflux = const_mflux; o.flomas = flomas;
  o.bdf = bdf; o.btime = btime; o.timestep = timestep; o.cn = cn;
  o.lbuoy = lcal[ien] && lbuoy; o.boussinesq = boussinesq;
  o.beta = beta; o.tref = tref; o.densit = densit; o.gravx = gravx; o.gravy = gravy; o.gravz = gravz;
  o.sol = solver_opts(ip);
  fc_piso_report rep;
  check(fc_upload(ctx, FC_PP, pp.data(), numTotal), "upload pp");  // pp is not reset between calls (PISO :199)
  check(fc_piso(ctx, &o, &rep), "fc_piso");
  const struct { int f; std::vector<dp> *h; size_t n; } dn[] = {
      {FC_U, &u, (size_t)numTotal}, {FC_V, &v, (size_t)numTotal}, {FC_W, &w, (size_t)numTotal},
      {FC_P, &p, (size_t)numTotal}, {FC_PP, &pp, (size_t)numTotal}, {FC_FLMASS, &flmass, (size_t)numInnerFaces},
      {FC_DPDXI, &dPdxi, 3 * (size_t)numCells}};
  for (auto &e : dn) check(fc_download(ctx, e.f, e.h->data(), e.n), "download");
  for (int k = 0; k < rep.nsolves && k < 16; ++k) {
    if (rep.rep[k].iters > 0) resor[ip] = rep.rep[k].res0;
    std::printf("  PCG(IC0):  Solving for %s, Initial residual = %10.3E, Final residual = %10.3E, No Iterations %d\n",
                title_mod::chvarSolver[ip], rep.rep[k].res0, rep.rep[k].resl, rep.rep[k].iters);
  }
  sumLocalContErr = rep.sumLocalContErr;
  globalContErr = rep.globalContErr;
  cumulativeContErr += globalContErr;
  std::printf("  time step continuity errors : sum local = %10.3E, global = %10.3E, cumulative = %10.3E\n",
              sumLocalContErr, globalContErr, cumulativeContErr);
}
void PISO_multiple_correction() { piso(false); }
void PIMPLE_multiple_correction() { piso(true); }

}  // namespace fcapp
