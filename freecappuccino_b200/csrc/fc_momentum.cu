// Momentum predictor `call calcuvw` (src/calcuvw.f90:3-557) on device-resident fields: the step
// immediately before the pressure-correction path.  It produces apu/apv/apw,
// u/v/w and leaves the boundary pressure and dPdxi that calcp reads, so a whole SIMPLE iteration
// (calcuvw + calcp) runs without a host round trip of any field.
//
// Several ranks (src-parallel/calcuvw.f90): processor faces get their own face kernel (apr = can, the SpMV strip
// of the BiCGStab solve), the diagonal is the running subtraction of the parallel build, and u, v, w, apu are
// exchanged at the end; `vis` must arrive with a current halo (fc_exchange), the reference exchanges it where it
// is updated.
//
// Launches per call: 3 x grad_gauss (U, V, W) + nipgrad x (bpres + grad_gauss) of the pressure,
// one face kernel (F threads), one row kernel (n threads), then per velocity component one
// diagonal / under-relaxation kernel (n threads) and the BiCGStab(DILU) solve of fc_krylov.cu.
// The per-index arithmetic lives in fc_momentum_body.cuh.
//
// Algorithmic bytes (n cells, F inner faces, B boundary faces), each datum once:
//   face kernel   2 idx 8 + 7 geometry 56 + flmass 8 + 6 results 48 per face = 120 F, plus per cell the
//                 gathered xc,yc,zc,vis,u,v,w,p (64) and four gradients (96) = 160 n
//   row kernel    6 face results 48 + 3 area components 24 (each face is read by its two cells: x2)
//                 + 3 map ints 12 per entry -> (72 + 12) * 2 F + off-diagonals 16 F = 184 F, per cell
//                 vol, den, 3 old velocities, 6 results = 88 n
//   component     row of a (8 nnz) + diag/ioffset 8 n + s, sp, phi, su, ap 40 n = 8 nnz + 48 n, x3
// All HBM-bound; no tensor cores (FP64 gather work).
#include "fc_momentum_body.cuh"
#include "fc_body_views.cuh"
#include "fc_reduce.cuh"

// fc_assemble.cu
int fc_grad_gauss_dev(fc_context *ctx, double *phi, double *grad, int nigrad);
int fc_bpres_dev(fc_context *ctx, double *p, const double *dPdxi, int istage);

namespace {

template <int OCC>   // CTAs per SM the register allocation must allow (FC_TUNE_FACE_OCC), see k_calcp_faces
__global__ void __launch_bounds__(256, OCC)
k_uvw_faces(fcm_geom g, fcm_flow f, fcm_opts o, fcm_faces out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.F) fcm_face(g, f, o, out, i);
}

__global__ void __launch_bounds__(256)
k_uvw_proc_faces(fcm_geom g, fcm_flow f, fcm_opts o, fcm_proc P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.npro) fcm_proc_face(g, f, o, P, i);
}

__global__ void __launch_bounds__(256)
k_uvw_rows(fcm_geom g, fcm_c2f m, fcm_slots sl, fcm_flow f, fcm_opts o, fcm_faces fa, fcm_proc P, fcm_rows r) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < g.n) fcm_row(g, m, sl, f, o, fa, P, r, c);
}

__global__ void __launch_bounds__(256)
k_uvw_component(fcm_geom g, fcm_c2f m, fcm_comp k) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < g.n) fcm_component(g, m, k, c);
}

// a = 0.5 a, the whole array (calcuvw.f90:386-389)
__global__ void k_halve(double *a, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = 0.5 * a[i];
}

fcm_flow flow_of(fc_context *ctx) {
  double **fl = ctx->field;
  return fcm_flow{fl[FC_U], fl[FC_V], fl[FC_W], fl[FC_P], fl[FC_DEN], fl[FC_VIS], fl[FC_FLMASS], fl[FC_FMI],
                  fl[FC_FMO], fl[FC_DUDXI], fl[FC_DVDXI], fl[FC_DWDXI], fl[FC_DPDXI], fl[FC_UO], fl[FC_VO],
                  fl[FC_WO], fl[FC_UOO], fl[FC_VOO], fl[FC_WOO], fl[FC_T]};
}
fcm_opts opts_of(const fc_calcuvw_opts *o) {
  return fcm_opts{o->scheme, o->limiter, o->gds, o->bdf, o->btime, o->timestep, o->cn, o->const_mflux,
                  o->gradPcmf, o->lbuoy, o->boussinesq, o->beta, o->tref, o->densit, o->gravx, o->gravy,
                  o->gravz, o->viscos};
}
fcm_faces faces_of(const fc_context *ctx) {
  const size_t F = (size_t)ctx->F;
  double *b = ctx->uvw_face;
  return fcm_faces{b, b + F, b + 2 * F, b + 3 * F, b + 4 * F, b + 5 * F};
}
fcm_proc proc_of(fc_context *ctx) {   // the four per-processor-face arrays follow the six per-inner-face ones
  const size_t np = (size_t)ctx->npro;
  double *b = ctx->uvw_face + 6 * (size_t)ctx->F;
  return fcm_proc{ctx->npro, ctx->m.iProcFacesStart, ctx->fpro, ctx->field[FC_FMPRO], ctx->field[FC_APR],
                  b, b + np, b + 2 * np, b + 3 * np};
}

int check_opts(fc_context *ctx, const fc_calcuvw_opts *o, const char *who) {
  if (!ctx->has_mesh || !ctx->has_csr || !ctx->c2f_off)
    FC_FAIL(FC_ERR_ARG, std::string(who) + ": call fc_set_mesh and fc_create_csr first");
  if (ctx->npro > 0 && ctx->nranks == 1)
    FC_FAIL(FC_ERR_ARG, std::string(who) + ": processor faces without fc_comm_init");
  if (o->nigrad < 1 || o->nipgrad < 0) FC_FAIL(FC_ERR_ARG, std::string(who) + ": nigrad >= 1, nipgrad >= 0");
  if (o->scheme < 0 || o->scheme > 5 || o->limiter < 0 || o->limiter > 7)
    FC_FAIL(FC_ERR_ARG, std::string(who) + ": unknown convection scheme / limiter");
  if ((o->bdf || o->cn) && !(o->timestep > 0.0)) FC_FAIL(FC_ERR_ARG, std::string(who) + ": timestep must be > 0");
  for (int k = 0; k < 3; ++k)
    if (!(o->urf[k] > 0.0)) FC_FAIL(FC_ERR_ARG, std::string(who) + ": urf must be > 0");
  if (ctx->F > 0 && ctx->n < 3)  // fieldManipulation.f90:433-435 reads flat element ijn+6 of dPdxi(3,numCells)
    FC_FAIL(FC_ERR_UNSUPPORTED, std::string(who) + ": needs numCells >= 3 (the reference's df(ijp,3) addressing)");
  return FC_OK;
}

}  // namespace

// calcuvw.f90:48-389: everything before the first per-component block
int fc_calcuvw_assemble_dev(fc_context *ctx, const fc_calcuvw_opts *o) {
  FC_CHECK(check_opts(ctx, o, "fc_calcuvw_assemble"));
  FC_CHECK(fc_momentum_fields(ctx));
  const int B = 256;
  cudaStream_t st = ctx->stream;
  FC_CUDA(cudaEventRecord(ctx->ev[2], st));
  // grad(U), grad(V), grad(W) (:59-61) and the first stage of calcPressDiv
  FC_CHECK(fc_grad_uvw_dev(ctx, o->nigrad, o->nipgrad >= 1));
  // calcPressDiv, remaining stages: boundary pressure + pressure gradient (fieldManipulation.f90:82-87)
  for (int istage = 2; istage <= o->nipgrad; ++istage) {
    FC_CHECK(fc_bpres_dev(ctx, ctx->field[FC_P], ctx->field[FC_DPDXI], istage));
    FC_CHECK(fc_grad_dev(ctx, ctx->field[FC_P], ctx->field[FC_DPDXI], o->nigrad));
  }
  const fcm_geom g = fcm_geom_of(ctx);
  const fcm_flow f = flow_of(ctx);
  const fcm_opts fo = opts_of(o);
  const fcm_faces fa = faces_of(ctx);
  if (ctx->F > 0) {
    // 128 registers, two CTAs per SM: tighter budgets spill and were measured slower (4.97 / 5.03 / 5.75 ms at 216^3
    // for 2 / 3 / 4 CTAs per SM, profiles/r02_face_occ.jsonl)
    k_uvw_faces<2><<<fc_blocks(ctx->F, B), B, 0, st>>>(g, f, fo, fa);
    FC_LAUNCH_CHECK();
  }
  double **fl = ctx->field;
  const fcm_rows r{fl[FC_A], fl[FC_SU], fl[FC_SV], fl[FC_SW], fl[FC_SPU], fl[FC_SPV], fl[FC_SP]};
  const fcm_proc P = proc_of(ctx);
  if (ctx->npro > 0) {   // src-parallel/calcuvw.f90:225-254; the halo of u, v, w, p and the gradients is current
    k_uvw_proc_faces<<<fc_blocks(ctx->npro, B), B, 0, st>>>(g, f, fo, P);   // (grad exchanges them), vis: the caller's
    FC_LAUNCH_CHECK();
  }
  k_uvw_rows<<<fc_blocks(ctx->n, B), B, 0, st>>>(g, fcm_c2f_of(ctx), fcm_slots_of(ctx), f, fo, fa, P, r);
  FC_LAUNCH_CHECK();
  if (o->cn) {
    k_halve<<<fc_blocks((size_t)ctx->nnz, B), B, 0, st>>>(fl[FC_A], (size_t)ctx->nnz);
    FC_LAUNCH_CHECK();
    if (ctx->npro > 0) {   // apr = 0.5 apr (src-parallel/calcuvw.f90:427)
      k_halve<<<fc_blocks((size_t)ctx->npro, B), B, 0, st>>>(fl[FC_APR], (size_t)ctx->npro);
      FC_LAUNCH_CHECK();
    }
  }
  FC_CUDA(cudaEventRecord(ctx->ev[3], st));
  return FC_OK;
}

// one velocity component (comp = 0, 1, 2): Crank-Nicolson sources, diagonal, under-relaxation, ap*,
// then `call bicgstab(u|v|w, iu|iv|iw)`   (calcuvw.f90:391-441, :447-498, :504-556)
int fc_calcuvw_component_dev(fc_context *ctx, const fc_calcuvw_opts *o, int comp, fc_solver_report *rep) {
  FC_CHECK(check_opts(ctx, o, "fc_calcuvw_component"));
  if (comp < 0 || comp > 2) FC_FAIL(FC_ERR_ARG, "fc_calcuvw_component: comp must be 0, 1 or 2");
  FC_CHECK(fc_momentum_fields(ctx));
  double **fl = ctx->field;
  const int s_f[3] = {FC_SU, FC_SV, FC_SW}, sp_f[3] = {FC_SPU, FC_SPV, FC_SP}, phi_f[3] = {FC_U, FC_V, FC_W};
  const int old_f[3] = {FC_UO, FC_VO, FC_WO}, ap_f[3] = {FC_APU, FC_APV, FC_APW};
  fcm_comp k{ctx->ioffset, ctx->diag, fl[FC_A], fl[s_f[comp]], fl[sp_f[comp]], fl[FC_SU], fl[ap_f[comp]],
             fl[phi_f[comp]], fl[old_f[comp]], fl[FC_DEN], 1.0 / o->urf[comp], 1.0 - o->urf[comp],   // init.f90:80-81
             o->sol.small, o->timestep, o->cn, comp > 0 ? 1 : 0, ctx->nranks > 1 ? 1 : 0, ctx->npro, fl[FC_APR]};
  k_uvw_component<<<fc_blocks(ctx->n, 256), 256, 0, ctx->stream>>>(fcm_geom_of(ctx), fcm_c2f_of(ctx), k);
  FC_LAUNCH_CHECK();
  fc_solver_opts so = o->sol;
  so.sor = o->sor[comp];
  so.nsw = o->nsw[comp];
  if (ctx->nranks > 1) so.parallel = 1;   // src-parallel/bicgstab.f90
  return fc_solve_device(ctx, FC_BICGSTAB, fl[phi_f[comp]], &so, rep, nullptr);
}

int fc_calcuvw_dev(fc_context *ctx, const fc_calcuvw_opts *o, fc_calcuvw_report *rep) {
  FC_CHECK(fc_calcuvw_assemble_dev(ctx, o));
  double solve_ms = 0.0;
  for (int comp = 0; comp < 3; ++comp) {
    FC_CHECK(fc_calcuvw_component_dev(ctx, o, comp, &rep->rep[comp]));
    solve_ms += ctx->tm.solve_ms;
  }
  if (ctx->npro > 0)   // exchange(u), (v), (w), (apu)   (src-parallel/calcuvw.f90:657-665)
    for (int fld : {FC_U, FC_V, FC_W, FC_APU}) FC_CHECK(fc_halo_exchange(ctx, ctx->field[fld]));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  float ams = 0.f;
  FC_CUDA(cudaEventElapsedTime(&ams, ctx->ev[2], ctx->ev[3]));
  ctx->tm.uvw_assemble_ms = ams;
  ctx->tm.uvw_solve_ms = solve_ms;
  return FC_OK;
}
