// extern "C" entry points of libfcapp_cuda.so (declared in include/fcapp.h): context life
// cycle, mesh / pattern set-up, field transfer and the thin wrappers that bind field ids to
// the device kernels.  No entry point has a host-side compute path: without a CUDA device
// fc_create fails with FC_ERR_NODEVICE and nothing else can be called.
#include "fc_internal.cuh"

int fc_grad_gauss_dev(fc_context *ctx, double *phi, double *grad, int nigrad);
int fc_grad_gauss_corrected_dev(fc_context *ctx, double *phi, double *grad, int zero_seed);
int fc_bpres_dev(fc_context *ctx, double *p, const double *dPdxi, int istage);
int fc_laplacian_dev(fc_context *ctx, double *mu, const double *phi);
int fc_calcp_assemble_dev(fc_context *ctx, const fc_calcp_opts *o);
int fc_calcp_dev(fc_context *ctx, const fc_calcp_opts *o, fc_calcp_report *rep);
int fc_calcp_correct_dev(fc_context *ctx, const fc_calcp_opts *o, int ipcorr);
int fc_calcp_finish_dev(fc_context *ctx, fc_calcp_report *rep);

namespace {

std::string g_create_error;

__global__ void k_add_int(int *p, int v, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] += v;
}
__global__ void k_copy_add_int(const int *src, int *dst, int v, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] + v;
}
__global__ void k_fill_double(double *p, double v, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// upload a 1-based Fortran index array and make it 0-based on the device
int upload_index(fc_context *ctx, int **dst, const int *host, size_t count) {
  FC_CHECK(fc_dev_alloc(ctx, dst, count + 2));
  if (count == 0) return FC_OK;
  FC_CUDA(cudaMemcpyAsync(*dst, host, sizeof(int) * count, cudaMemcpyHostToDevice, ctx->stream));
  k_add_int<<<fc_blocks(count, 256), 256, 0, ctx->stream>>>(*dst, -1, count);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

int upload_real(fc_context *ctx, double **dst, const double *host, size_t count) {
  FC_CHECK(fc_dev_alloc(ctx, dst, count));
  if (count == 0 || !host) return FC_OK;
  FC_CUDA(cudaMemcpyAsync(*dst, host, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
  return FC_OK;
}

// download a 0-based device index array as 1-based
int download_index(fc_context *ctx, const int *dev, int *host, size_t count, int *tmp) {
  if (!host || count == 0) return FC_OK;
  k_copy_add_int<<<fc_blocks(count, 256), 256, 0, ctx->stream>>>(dev, tmp, 1, count);
  FC_LAUNCH_CHECK();
  FC_CUDA(cudaMemcpyAsync(host, tmp, sizeof(int) * count, cudaMemcpyDeviceToHost, ctx->stream));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  return FC_OK;
}

int alloc_field(fc_context *ctx, int f, size_t count) {
  FC_CHECK(fc_dev_alloc(ctx, &ctx->field[f], count));
  ctx->field_n[f] = count;
  FC_CUDA(cudaMemsetAsync(ctx->field[f], 0, sizeof(double) * (count ? count : 1), ctx->stream));
  return FC_OK;
}

int alloc_fields(fc_context *ctx) {
  const size_t n = ctx->n, NT = ctx->NT, NP = (size_t)ctx->n + ctx->npro, F = ctx->F;
  for (int f : {FC_U, FC_V, FC_W, FC_P, FC_PP, FC_DEN, FC_SCRATCH_T, FC_USER0, FC_USER1, FC_USER2, FC_USER3}) FC_CHECK(alloc_field(ctx, f, NT > NP ? NT : NP));
  FC_CHECK(alloc_field(ctx, FC_FLMASS, F));
  for (int f : {FC_APU, FC_APV, FC_APW}) FC_CHECK(alloc_field(ctx, f, NP));
  for (int f : {FC_DUDXI, FC_DVDXI, FC_DWDXI, FC_DPDXI}) FC_CHECK(alloc_field(ctx, f, 3 * NP));  // (3,numPCells)
  FC_CHECK(alloc_field(ctx, FC_A, (size_t)ctx->nnz + 2));
  ctx->field_n[FC_A] = ctx->nnz;
  FC_CHECK(alloc_field(ctx, FC_SU, n));
  FC_CHECK(alloc_field(ctx, FC_RES, n));
  FC_CHECK(alloc_field(ctx, FC_FMI, ctx->m.ninl));
  FC_CHECK(alloc_field(ctx, FC_FMO, ctx->m.nout));
  FC_CHECK(alloc_field(ctx, FC_APR, ctx->npro));
  FC_CHECK(alloc_field(ctx, FC_FMPRO, ctx->npro));
  k_fill_double<<<fc_blocks(ctx->field_n[FC_DEN], 256), 256, 0, ctx->stream>>>(ctx->field[FC_DEN], 1.0,
                                                                             ctx->field_n[FC_DEN]);
  FC_LAUNCH_CHECK();
  FC_CHECK(fc_dev_alloc(ctx, &ctx->coef, F + ctx->npro));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->facev, (size_t)ctx->NF));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->gtmp, 3 * NP));
  return FC_OK;
}

int check_field(fc_context *ctx, int f, size_t min_n, const char *who) {
  if (f >= FC_VIS && f < FC_NUM_FIELDS && !ctx->field[f] && ctx->has_mesh) FC_CHECK(fc_momentum_fields(ctx));
  if (f < 0 || f >= FC_NUM_FIELDS || !ctx->field[f])
    FC_FAIL(FC_ERR_ARG, std::string(who) + ": unknown or unallocated field id " + std::to_string(f));
  if (ctx->field_n[f] < min_n)
    FC_FAIL(FC_ERR_ARG, std::string(who) + ": field " + std::to_string(f) + " too small");
  return FC_OK;
}

void free_all(fc_context *ctx) {
  for (void *p : {(void *)ctx->owner, (void *)ctx->neigh, (void *)ctx->xc, (void *)ctx->yc, (void *)ctx->zc,
                  (void *)ctx->vol, (void *)ctx->arx, (void *)ctx->ary, (void *)ctx->arz, (void *)ctx->xf,
                  (void *)ctx->yf, (void *)ctx->zf, (void *)ctx->facint, (void *)ctx->fpro, (void *)ctx->c2f_off,
                  (void *)ctx->c2f_face, (void *)ctx->c2f_other, (void *)ctx->c2f_pos, (void *)ctx->ioffset,
                  (void *)ctx->ja, (void *)ctx->diag, (void *)ctx->icj, (void *)ctx->jci, (void *)ctx->tpos,
                  (void *)ctx->pk, (void *)ctx->zk, (void *)ctx->dd, (void *)ctx->reso, (void *)ctx->uk,
                  (void *)ctx->vk, (void *)ctx->adiag, (void *)ctx->tt, (void *)ctx->coef, (void *)ctx->facev,
                  (void *)ctx->gtmp, (void *)ctx->partials, (void *)ctx->sc, (void *)ctx->bufind,
                  (void *)ctx->sendbuf, (void *)ctx->strip_off, (void *)ctx->strip_idx,
                  (void *)ctx->strip_any32, (void *)ctx->persist, (void *)ctx->uvw_face, (void *)ctx->hcoef, (void *)ctx->dmat, (void *)ctx->tja, (void *)ctx->gtmp3, (void *)ctx->sweep_chk,
                  (void *)ctx->dmatqr, (void *)ctx->hist, (void *)ctx->jcode, (void *)ctx->jdict})
    if (p) cudaFree(p);
  for (int f = 0; f < FC_NUM_FIELDS; ++f)
    if (ctx->field[f]) cudaFree(ctx->field[f]);
  fc_levels_free(ctx->lower);
  fc_levels_free(ctx->upper);
  fc_levels_free(ctx->tile_lower);
  fc_levels_free(ctx->tile_upper);
  if (ctx->sc_host) cudaFreeHost(ctx->sc_host);
  if (ctx->persist_host) cudaFreeHost(ctx->persist_host);
}

}  // namespace

// Fields of the momentum predictor (module variables: vis, uo.., t; sparse_matrix: sv, sw, spu, spv, sp) and its
// per-face scratch: allocated at the first fc_upload / fc_calcuvw that names them, not per call.
int fc_momentum_fields(fc_context *ctx) {
  if (!ctx->has_mesh) FC_FAIL(FC_ERR_ARG, "momentum fields: call fc_set_mesh first");
  if (ctx->uvw_face) return FC_OK;
  const size_t n = ctx->n, NT = ctx->NT, NP = (size_t)ctx->n + ctx->npro;
  for (int f : {FC_VIS, FC_UO, FC_VO, FC_WO, FC_UOO, FC_VOO, FC_WOO, FC_T}) FC_CHECK(alloc_field(ctx, f, NT > NP ? NT : NP));
  for (int f : {FC_SV, FC_SW, FC_SPU, FC_SPV, FC_SP}) FC_CHECK(alloc_field(ctx, f, n));
  FC_CHECK(fc_dev_alloc(ctx, &ctx->uvw_face, 6 * (size_t)ctx->F + 4 * (size_t)ctx->npro));
  return FC_OK;
}

extern "C" {

int fc_version(void) { return 101; }

const char *fc_last_error(const fc_context *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int fc_create(int device, fc_context **out) {
  if (!out) return FC_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("fc_create: no CUDA device (") + cudaGetErrorString(e) +
                     "); libfcapp_cuda has no CPU fallback";
    return FC_ERR_NODEVICE;
  }
  if (device < 0 || device >= count) {
    g_create_error = "fc_create: device index out of range";
    return FC_ERR_ARG;
  }
  fc_context *ctx = new fc_context();
  ctx->device = device;
  auto fail = [&](int code) {
    g_create_error = ctx->err;
    delete ctx;
    return code;
  };
  auto init = [&]() -> int {
    FC_CUDA(cudaSetDevice(device));
    FC_CUDA(cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device));
    FC_CUDA(cudaDeviceGetAttribute(&ctx->l2_bytes, cudaDevAttrL2CacheSize, device));
    if (ctx->sms < 1) ctx->sms = FC_SMS;
    FC_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (auto &ev : ctx->ev) FC_CUDA(cudaEventCreate(&ev));
    FC_CHECK(fc_dev_alloc(ctx, &ctx->sc, 1));
    FC_CUDA(cudaMemset(ctx->sc, 0, sizeof(fc_scalars)));
    FC_CUDA(cudaMallocHost((void **)&ctx->sc_host, sizeof(fc_scalars)));
    memset(ctx->sc_host, 0, sizeof(fc_scalars));
    FC_CHECK(fc_dev_alloc(ctx, &ctx->partials, (size_t)FC_MAX_RED * 2048));
    FC_CHECK(fc_dev_alloc(ctx, &ctx->persist, 1));
    FC_CUDA(cudaMemset(ctx->persist, 0, sizeof(fc_persist_state)));
    FC_CUDA(cudaMallocHost((void **)&ctx->persist_host, sizeof(fc_persist_state)));
    memset(ctx->persist_host, 0, sizeof(fc_persist_state));
    return FC_OK;
  };
  int s = init();
  if (s != FC_OK) return fail(s);
  *out = ctx;
  return FC_OK;
}

int fc_destroy(fc_context *ctx) {
  if (!ctx) return FC_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto &e : ctx->spmv_ev) cudaEventDestroy(e);
  fc_p2p_close(ctx);
  fc_comm_destroy(ctx);
  free_all(ctx);
  for (auto &ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return FC_OK;
}

void *fc_stream(fc_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int fc_synchronize(fc_context *ctx) {
  if (!ctx) return FC_ERR_ARG;
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  return FC_OK;
}

int fc_set_mesh(fc_context *ctx, const fc_mesh_desc *m) {
  if (!ctx || !m) return FC_ERR_ARG;
  FC_CUDA(cudaSetDevice(ctx->device));
  if (ctx->csr_external) FC_FAIL(FC_ERR_ARG, "fc_set_mesh: context already adopted an explicit CSR pattern");
  fc_p2p_close(ctx);
  if (m->noc > 0)
    FC_FAIL(FC_ERR_UNSUPPORTED, "O-C grid cuts (noc > 0) are not on the GPU path (none of the target cases has them)");
  if (m->numCells < 1 || m->numInnerFaces < 0 || m->numFaces < m->numInnerFaces ||
      m->numTotal < m->numCells + m->npro)
    FC_FAIL(FC_ERR_ARG, "fc_set_mesh: inconsistent sizes");
  if (!m->owner || (m->numInnerFaces && !m->neighbour) || !m->xc || !m->yc || !m->zc || !m->vol || !m->arx ||
      !m->ary || !m->arz || !m->xf || !m->yf || !m->zf || (m->numInnerFaces && !m->facint))
    FC_FAIL(FC_ERR_ARG, "fc_set_mesh: NULL geometry array");
  if (m->npro > 0 && (!m->fpro || !m->neighbProcNo || !m->neighbProcOffset || m->numConnections < 1))
    FC_FAIL(FC_ERR_ARG, "fc_set_mesh: processor boundary without fpro / neighbProcNo / neighbProcOffset");
  // every kernel indexes with these arrays unchecked: reject a mesh that would send them out of bounds
  {
    const int nb = m->ninl + m->nout + m->nsym + m->nwal + m->npru;
    if (m->ninl < 0 || m->nout < 0 || m->nsym < 0 || m->nwal < 0 || m->npru < 0 || m->npro < 0 ||
        m->numTotal < m->numCells + m->npro + nb)
      FC_FAIL(FC_ERR_ARG, "fc_set_mesh: numTotal is smaller than numCells + npro + the boundary face counts");
    const struct { const char *name; int start, count; } patch[] = {
        {"processor", m->iProcFacesStart, m->npro},   {"inlet", m->iInletFacesStart, m->ninl},
        {"outlet", m->iOutletFacesStart, m->nout},    {"symmetry", m->iSymmetryFacesStart, m->nsym},
        {"wall", m->iWallFacesStart, m->nwal},        {"prOutlet", m->iPressOutletFacesStart, m->npru}};
    for (const auto &pt : patch)
      if (pt.count > 0 && (pt.start < m->numInnerFaces || pt.start > m->numFaces - pt.count))
        FC_FAIL(FC_ERR_ARG, std::string("fc_set_mesh: ") + pt.name + " faces lie outside numInnerFaces+1..numFaces");
    for (int i = 0; i < m->numFaces; ++i)
      if (m->owner[i] < 1 || m->owner[i] > m->numCells)
        FC_FAIL(FC_ERR_ARG, "fc_set_mesh: owner(" + std::to_string(i + 1) + ") = " + std::to_string(m->owner[i]) +
                                " is outside 1..numCells");
    for (int i = 0; i < m->numInnerFaces; ++i)
      if (m->neighbour[i] < 1 || m->neighbour[i] > m->numCells || m->neighbour[i] == m->owner[i])
        FC_FAIL(FC_ERR_ARG, "fc_set_mesh: neighbour(" + std::to_string(i + 1) + ") = " + std::to_string(m->neighbour[i]) +
                                " is outside 1..numCells or equals its owner");
  }
  if (ctx->hcoef) { cudaFree(ctx->hcoef); ctx->hcoef = nullptr; }
  ctx->grad_method = ctx->grad_limiter = 0;   // the least-squares matrices belong to the old mesh
  if (ctx->uvw_face) {  // momentum fields are sized by the mesh: drop them, they come back on first use
    cudaFree(ctx->uvw_face);
    ctx->uvw_face = nullptr;
    for (int f = FC_VIS; f < FC_NUM_FIELDS; ++f) {
      if (ctx->field[f]) cudaFree(ctx->field[f]);
      ctx->field[f] = nullptr;
      ctx->field_n[f] = 0;
    }
  }
  ctx->m = *m;
  ctx->n = m->numCells; ctx->F = m->numInnerFaces; ctx->NF = m->numFaces; ctx->NT = m->numTotal;
  ctx->npro = m->npro; ctx->NP = m->numCells + m->npro;
  ctx->nnz = m->numCells + 2 * m->numInnerFaces;  // mesh_geometry_and_topology.f90:580
  const size_t NP = ctx->NP, NF = ctx->NF, F = ctx->F;
  FC_CHECK(upload_index(ctx, &ctx->owner, m->owner, NF));
  FC_CHECK(upload_index(ctx, &ctx->neigh, m->neighbour, F));
  FC_CHECK(upload_real(ctx, &ctx->xc, m->xc, NP));
  FC_CHECK(upload_real(ctx, &ctx->yc, m->yc, NP));
  FC_CHECK(upload_real(ctx, &ctx->zc, m->zc, NP));
  FC_CHECK(upload_real(ctx, &ctx->vol, m->vol, NP));
  FC_CHECK(upload_real(ctx, &ctx->arx, m->arx, NF));
  FC_CHECK(upload_real(ctx, &ctx->ary, m->ary, NF));
  FC_CHECK(upload_real(ctx, &ctx->arz, m->arz, NF));
  FC_CHECK(upload_real(ctx, &ctx->xf, m->xf, NF));
  FC_CHECK(upload_real(ctx, &ctx->yf, m->yf, NF));
  FC_CHECK(upload_real(ctx, &ctx->zf, m->zf, NF));
  FC_CHECK(upload_real(ctx, &ctx->facint, m->facint, F));
  FC_CHECK(upload_real(ctx, &ctx->fpro, m->fpro, (size_t)m->npro));
  ctx->nbr_rank.clear();
  ctx->nbr_off.clear();
  if (m->npro > 0) {
    for (int c = 0; c < m->numConnections; ++c) ctx->nbr_rank.push_back(m->neighbProcNo[c]);
    for (int c = 0; c <= m->numConnections; ++c) ctx->nbr_off.push_back(m->neighbProcOffset[c] - 1);
    // the pack kernels and the send / receive lengths index with this table unchecked
    if (ctx->nbr_off.front() != 0) FC_FAIL(FC_ERR_ARG, "fc_set_mesh: neighbProcOffset(1) must be 1");
    for (int c = 0; c < m->numConnections; ++c) {
      if (ctx->nbr_off[c + 1] < ctx->nbr_off[c])
        FC_FAIL(FC_ERR_ARG, "fc_set_mesh: neighbProcOffset decreases at connection " + std::to_string(c + 1));
      if (ctx->nbr_rank[c] < 0) FC_FAIL(FC_ERR_ARG, "fc_set_mesh: negative neighbProcNo");
      if (ctx->nranks > 1 && (ctx->nbr_rank[c] >= ctx->nranks || ctx->nbr_rank[c] == ctx->rank))
        FC_FAIL(FC_ERR_ARG, "fc_set_mesh: neighbProcNo(" + std::to_string(c + 1) + ") = " + std::to_string(ctx->nbr_rank[c]) +
                                " is not another rank of the communicator");
    }
    if (ctx->nbr_off.back() != m->npro) FC_FAIL(FC_ERR_ARG, "fc_set_mesh: neighbProcOffset does not cover npro");
    // bufind(i) = owner(iProcFacesStart + i)   (src-parallel/mesh_geometry_and_topology.f90:879-881)
    FC_CHECK(fc_dev_alloc(ctx, &ctx->bufind, (size_t)m->npro));
    FC_CUDA(cudaMemcpyAsync(ctx->bufind, ctx->owner + m->iProcFacesStart, sizeof(int) * (size_t)m->npro,
                            cudaMemcpyDeviceToDevice, ctx->stream));
    FC_CHECK(fc_dev_alloc(ctx, &ctx->sendbuf, 3 * (size_t)m->npro));
  }
  // the descriptor's pointers are the caller's; keep only the scalars
  ctx->m.owner = ctx->m.neighbour = nullptr;
  ctx->m.xc = ctx->m.yc = ctx->m.zc = ctx->m.vol = nullptr;
  ctx->m.arx = ctx->m.ary = ctx->m.arz = ctx->m.xf = ctx->m.yf = ctx->m.zf = nullptr;
  ctx->m.facint = ctx->m.fpro = nullptr;
  ctx->m.neighbProcNo = ctx->m.neighbProcOffset = nullptr;
  FC_CHECK(alloc_fields(ctx));
  ctx->scratch_n = 0;
  ctx->has_mesh = true;
  ctx->has_csr = false;
  ctx->has_levels = false;
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  return FC_OK;
}

int fc_create_csr(fc_context *ctx, int *ioffset, int *ja, int *diag, int *icell_jcell, int *jcell_icell) {
  if (!ctx) return FC_ERR_ARG;
  if (!ctx->has_mesh) FC_FAIL(FC_ERR_ARG, "fc_create_csr: call fc_set_mesh first");
  FC_CUDA(cudaSetDevice(ctx->device));
  FC_CHECK(fc_csr_build(ctx));
  FC_CHECK(fc_c2f_build(ctx));
  ctx->has_levels = false;
  FC_CHECK(fc_alloc_solver_scratch(ctx));
  int *tmp = nullptr;
  FC_CHECK(fc_dev_alloc(ctx, &tmp, (size_t)ctx->nnz + 1));
  int s = download_index(ctx, ctx->ioffset, ioffset, (size_t)ctx->n + 1, tmp);
  if (s == FC_OK) s = download_index(ctx, ctx->ja, ja, (size_t)ctx->nnz, tmp);
  if (s == FC_OK) s = download_index(ctx, ctx->diag, diag, (size_t)ctx->n, tmp);
  if (s == FC_OK) s = download_index(ctx, ctx->icj, icell_jcell, (size_t)ctx->F, tmp);
  if (s == FC_OK) s = download_index(ctx, ctx->jci, jcell_icell, (size_t)ctx->F, tmp);
  cudaFree(tmp);
  return s;
}

int fc_field_size(const fc_context *ctx, int field, size_t *n) {
  if (!ctx || !n || field < 0 || field >= FC_NUM_FIELDS) return FC_ERR_ARG;
  *n = ctx->field_n[field];
  return FC_OK;
}

int fc_upload(fc_context *ctx, int field, const double *host, size_t n) {
  if (!ctx || !host) return FC_ERR_ARG;
  FC_CHECK(check_field(ctx, field, n, "fc_upload"));
  FC_CUDA(cudaMemcpyAsync(ctx->field[field], host, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  return FC_OK;
}

int fc_download(fc_context *ctx, int field, double *host, size_t n) {
  if (!ctx || !host) return FC_ERR_ARG;
  FC_CHECK(check_field(ctx, field, n, "fc_download"));
  FC_CUDA(cudaMemcpyAsync(host, ctx->field[field], sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  return FC_OK;
}

int fc_fill(fc_context *ctx, int field, double value) {
  if (!ctx) return FC_ERR_ARG;
  FC_CHECK(check_field(ctx, field, 0, "fc_fill"));
  const size_t n = ctx->field_n[field];
  if (n == 0) return FC_OK;
  k_fill_double<<<fc_blocks(n, 256), 256, 0, ctx->stream>>>(ctx->field[field], value, n);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

int fc_copy(fc_context *ctx, int src_field, int dst_field) {
  if (!ctx) return FC_ERR_ARG;
  FC_CHECK(check_field(ctx, src_field, 0, "fc_copy"));
  FC_CHECK(check_field(ctx, dst_field, 0, "fc_copy"));
  const size_t n = ctx->field_n[src_field] < ctx->field_n[dst_field] ? ctx->field_n[src_field] : ctx->field_n[dst_field];
  FC_CUDA(cudaMemcpyAsync(ctx->field[dst_field], ctx->field[src_field], sizeof(double) * n, cudaMemcpyDeviceToDevice,
                          ctx->stream));
  return FC_OK;
}

int fc_set_spmv_sampling(fc_context *ctx, int max_samples) {
  if (!ctx || max_samples < 0 || max_samples > 4096) return FC_ERR_ARG;
  for (auto &e : ctx->spmv_ev) cudaEventDestroy(e);
  ctx->spmv_ev.assign((size_t)2 * max_samples, nullptr);
  for (auto &e : ctx->spmv_ev) FC_CUDA(cudaEventCreate(&e));
  ctx->spmv_sampled = 0;
  return FC_OK;
}

int fc_spmv(fc_context *ctx, int x_field, int y_field) {
  if (!ctx) return FC_ERR_ARG;
  if (!ctx->has_csr) FC_FAIL(FC_ERR_ARG, "fc_spmv: no CSR pattern");
  FC_CHECK(check_field(ctx, x_field, (size_t)ctx->n, "fc_spmv"));
  FC_CHECK(check_field(ctx, y_field, (size_t)ctx->n, "fc_spmv"));
  if (x_field == y_field) FC_FAIL(FC_ERR_ARG, "fc_spmv: x and y must differ");
  if (ctx->npro > 0) FC_CHECK(fc_halo_exchange(ctx, ctx->field[x_field]));
  return fc_launch_spmv(ctx, ctx->field[FC_A], ctx->field[x_field], ctx->field[y_field]);
}

int fc_time_spmv(fc_context *ctx, int x_field, int y_field, int reps, double *mean_ms) {
  if (!ctx || !mean_ms || reps < 1) return FC_ERR_ARG;
  FC_CHECK(fc_spmv(ctx, x_field, y_field));  // argument checks + warm-up
  FC_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  for (int r = 0; r < reps; ++r)
    FC_CHECK(fc_launch_spmv(ctx, ctx->field[FC_A], ctx->field[x_field], ctx->field[y_field]));
  FC_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  FC_CUDA(cudaEventSynchronize(ctx->ev[1]));
  float ms = 0.f;
  FC_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  *mean_ms = (double)ms / reps;
  ctx->tm.spmv_ms = *mean_ms;
  return FC_OK;
}

int fc_grad_gauss(fc_context *ctx, int phi_field, int grad_field, int nigrad) {
  if (!ctx) return FC_ERR_ARG;
  FC_CHECK(check_field(ctx, phi_field, (size_t)ctx->NT, "fc_grad_gauss"));
  FC_CHECK(check_field(ctx, grad_field, 3 * (size_t)ctx->n, "fc_grad_gauss"));
  return fc_grad_gauss_dev(ctx, ctx->field[phi_field], ctx->field[grad_field], nigrad);
}

int fc_grad_gauss_corrected(fc_context *ctx, int phi_field, int grad_field, int zero_seed) {
  if (!ctx) return FC_ERR_ARG;
  FC_CHECK(check_field(ctx, phi_field, (size_t)ctx->NT, "fc_grad_gauss_corrected"));
  FC_CHECK(check_field(ctx, grad_field, 3 * (size_t)ctx->n, "fc_grad_gauss_corrected"));
  return fc_grad_gauss_corrected_dev(ctx, ctx->field[phi_field], ctx->field[grad_field], zero_seed);
}

int fc_bpres(fc_context *ctx, int p_field, int istage) {
  if (!ctx) return FC_ERR_ARG;
  FC_CHECK(check_field(ctx, p_field, (size_t)ctx->NT, "fc_bpres"));
  return fc_bpres_dev(ctx, ctx->field[p_field], ctx->field[FC_DPDXI], istage);
}

int fc_laplacian(fc_context *ctx, int mu_field, int phi_field) {
  if (!ctx) return FC_ERR_ARG;
  FC_CHECK(check_field(ctx, mu_field, (size_t)ctx->NP, "fc_laplacian"));
  FC_CHECK(check_field(ctx, phi_field, (size_t)ctx->NT, "fc_laplacian"));
  return fc_laplacian_dev(ctx, ctx->field[mu_field], ctx->field[phi_field]);
}

int fc_solve(fc_context *ctx, int solver, int fi_field, const fc_solver_opts *o, fc_solver_report *rep) {
  if (!ctx || !o || !rep) return FC_ERR_ARG;
  FC_CUDA(cudaSetDevice(ctx->device));
  FC_CHECK(check_field(ctx, fi_field, (size_t)ctx->n + ctx->npro, "fc_solve"));
  return fc_solve_device(ctx, solver, ctx->field[fi_field], o, rep, nullptr);
}

static int solve_host_impl(fc_context *ctx, int solver, const double *a, const double *su, double *fi, double *res,
                           const fc_solver_opts *o, fc_solver_report *rep, double *hist) {
  const size_t n = ctx->n, nt = ctx->has_mesh ? (size_t)ctx->NT : n;
  cudaStream_t st = ctx->stream;
  FC_CUDA(cudaMemcpyAsync(ctx->field[FC_A], a, sizeof(double) * (size_t)ctx->nnz, cudaMemcpyHostToDevice, st));
  FC_CUDA(cudaMemcpyAsync(ctx->field[FC_SU], su, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  FC_CUDA(cudaMemcpyAsync(ctx->field[FC_SCRATCH_T], fi, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
  FC_CHECK(fc_solve_device(ctx, solver, ctx->field[FC_SCRATCH_T], o, rep, hist));
  FC_CUDA(cudaMemcpyAsync(fi, ctx->field[FC_SCRATCH_T], sizeof(double) * (n + ctx->npro), cudaMemcpyDeviceToHost, st));
  if (res) FC_CUDA(cudaMemcpyAsync(res, ctx->field[FC_RES], sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  FC_CUDA(cudaStreamSynchronize(st));
  return FC_OK;
}

int fc_solve_host(fc_context *ctx, int solver, const double *a, const double *su, double *fi, double *res,
                  const fc_solver_opts *o, fc_solver_report *rep) {
  if (!ctx || !a || !su || !fi || !o || !rep) return FC_ERR_ARG;
  if (!ctx->has_csr) FC_FAIL(FC_ERR_ARG, "fc_solve_host: no CSR pattern");
  FC_CUDA(cudaSetDevice(ctx->device));
  return solve_host_impl(ctx, solver, a, su, fi, res, o, rep, nullptr);
}

int fc_solve_csr(fc_context *ctx, int solver, int numCells, int nnz, const int *ioffset, const int *ja,
                 const int *diag, const double *a, const double *su, double *fi, const fc_solver_opts *o,
                 fc_solver_report *rep, double *hist) {
  if (!ctx || !ioffset || !ja || !diag || !a || !su || !fi || !o || !rep) return FC_ERR_ARG;
  if (ctx->has_mesh) FC_FAIL(FC_ERR_ARG, "fc_solve_csr: context is bound to a mesh; use fc_solve_host");
  if (numCells < 1 || nnz < numCells) FC_FAIL(FC_ERR_ARG, "fc_solve_csr: bad sizes");
  // the sweeps and the SpMV index with these arrays unchecked (1-based, LIS_linear_solver_library.f95:106-119)
  if (ioffset[0] != 1 || ioffset[numCells] != nnz + 1) FC_FAIL(FC_ERR_ARG, "fc_solve_csr: ioffset must run from 1 to nnz+1");
  for (int i = 0; i < numCells; ++i) {
    if (ioffset[i + 1] < ioffset[i]) FC_FAIL(FC_ERR_ARG, "fc_solve_csr: ioffset decreases at row " + std::to_string(i + 1));
    if (diag[i] < ioffset[i] || diag[i] >= ioffset[i + 1] || ja[diag[i] - 1] != i + 1)
      FC_FAIL(FC_ERR_ARG, "fc_solve_csr: diag(" + std::to_string(i + 1) + ") does not point at the diagonal of its row");
  }
  for (int k = 0; k < nnz; ++k)
    if (ja[k] < 1 || ja[k] > numCells) FC_FAIL(FC_ERR_ARG, "fc_solve_csr: ja(" + std::to_string(k + 1) + ") is outside 1..numCells");
  FC_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->csr_external || ctx->n != numCells || ctx->nnz != nnz) {
    ctx->n = numCells; ctx->nnz = nnz; ctx->npro = 0; ctx->NP = numCells; ctx->NT = numCells; ctx->F = 0; ctx->NF = 0;
    FC_CHECK(alloc_field(ctx, FC_A, (size_t)nnz + 2));
    ctx->field_n[FC_A] = nnz;
    FC_CHECK(alloc_field(ctx, FC_SU, (size_t)numCells));
    FC_CHECK(alloc_field(ctx, FC_RES, (size_t)numCells));
    FC_CHECK(alloc_field(ctx, FC_SCRATCH_T, (size_t)numCells));
    ctx->scratch_n = 0;
    ctx->csr_external = true;
  }
  // the pattern may change between calls with equal sizes: always re-adopt it
  FC_CHECK(upload_index(ctx, &ctx->ioffset, ioffset, (size_t)numCells + 1));
  FC_CHECK(upload_index(ctx, &ctx->ja, ja, (size_t)nnz));
  FC_CHECK(upload_index(ctx, &ctx->diag, diag, (size_t)numCells));
  ctx->has_levels = false;
  FC_CHECK(fc_csr_post(ctx));
  return solve_host_impl(ctx, solver, a, su, fi, nullptr, o, rep, hist);
}

int fc_calcp_assemble(fc_context *ctx, const fc_calcp_opts *o) {
  if (!ctx || !o) return FC_ERR_ARG;
  FC_CUDA(cudaSetDevice(ctx->device));
  return fc_calcp_assemble_dev(ctx, o);
}

int fc_calcp_correct(fc_context *ctx, const fc_calcp_opts *o, int ipcorr, fc_calcp_report *rep) {
  if (!ctx || !o) return FC_ERR_ARG;
  FC_CUDA(cudaSetDevice(ctx->device));
  FC_CHECK(fc_calcp_correct_dev(ctx, o, ipcorr));
  if (ipcorr == o->npcor) {
    fc_calcp_report tmp;
    FC_CHECK(fc_calcp_finish_dev(ctx, rep ? rep : &tmp));
  }
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  return FC_OK;
}

int fc_calcp(fc_context *ctx, const fc_calcp_opts *o, fc_calcp_report *rep) {
  if (!ctx || !o || !rep) return FC_ERR_ARG;
  FC_CUDA(cudaSetDevice(ctx->device));
  return fc_calcp_dev(ctx, o, rep);
}

int fc_calcp_host(fc_context *ctx, const fc_calcp_opts *o, double *u, double *v, double *w, double *p, double *pp,
                  const double *apu, const double *apv, const double *apw, double *flmass, fc_calcp_report *rep) {
  if (!ctx || !o || !rep || !u || !v || !w || !p || !apu || !apv || !apw) return FC_ERR_ARG;
  if (!ctx->has_mesh) FC_FAIL(FC_ERR_ARG, "fc_calcp_host: no mesh");
  FC_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t NT = ctx->NT, NP = ctx->NP;
  const struct { int f; const double *h; size_t n; } up[] = {{FC_U, u, NT}, {FC_V, v, NT}, {FC_W, w, NT}, {FC_P, p, NT},
                                                            {FC_APU, apu, NP}, {FC_APV, apv, NP}, {FC_APW, apw, NP}};
  for (auto &t : up) FC_CUDA(cudaMemcpyAsync(ctx->field[t.f], t.h, sizeof(double) * t.n, cudaMemcpyHostToDevice, st));
  FC_CHECK(fc_calcp_dev(ctx, o, rep));
  const struct { int f; double *h; size_t n; } dn[] = {{FC_U, u, NT}, {FC_V, v, NT}, {FC_W, w, NT}, {FC_P, p, NT},
                                                      {FC_PP, pp, NT}, {FC_FLMASS, flmass, (size_t)ctx->F}};
  for (auto &t : dn)
    if (t.h) FC_CUDA(cudaMemcpyAsync(t.h, ctx->field[t.f], sizeof(double) * t.n, cudaMemcpyDeviceToHost, st));
  FC_CUDA(cudaStreamSynchronize(st));
  return FC_OK;
}

int fc_calcuvw_assemble(fc_context *ctx, const fc_calcuvw_opts *o) {
  if (!ctx || !o) return FC_ERR_ARG;
  FC_CUDA(cudaSetDevice(ctx->device));
  return fc_calcuvw_assemble_dev(ctx, o);
}

int fc_calcuvw_component(fc_context *ctx, const fc_calcuvw_opts *o, int comp, fc_solver_report *rep) {
  if (!ctx || !o || !rep) return FC_ERR_ARG;
  FC_CUDA(cudaSetDevice(ctx->device));
  return fc_calcuvw_component_dev(ctx, o, comp, rep);
}

int fc_calcuvw(fc_context *ctx, const fc_calcuvw_opts *o, fc_calcuvw_report *rep) {
  if (!ctx || !o || !rep) return FC_ERR_ARG;
  FC_CUDA(cudaSetDevice(ctx->device));
  return fc_calcuvw_dev(ctx, o, rep);
}

int fc_calcuvw_host(fc_context *ctx, const fc_calcuvw_opts *o, double *u, double *v, double *w, double *p,
                    const double *vis, const double *flmass, double *apu, double *apv, double *apw,
                    fc_calcuvw_report *rep) {
  if (!ctx || !o || !rep || !u || !v || !w || !p || !vis || !flmass) return FC_ERR_ARG;
  if (!ctx->has_mesh) FC_FAIL(FC_ERR_ARG, "fc_calcuvw_host: no mesh");
  FC_CUDA(cudaSetDevice(ctx->device));
  FC_CHECK(fc_momentum_fields(ctx));
  cudaStream_t st = ctx->stream;
  const size_t NT = ctx->NT, n = ctx->n;
  const struct { int f; const double *h; size_t n; } up[] = {{FC_U, u, NT}, {FC_V, v, NT}, {FC_W, w, NT}, {FC_P, p, NT},
                                                            {FC_VIS, vis, NT}, {FC_FLMASS, flmass, (size_t)ctx->F}};
  for (auto &t : up) FC_CUDA(cudaMemcpyAsync(ctx->field[t.f], t.h, sizeof(double) * t.n, cudaMemcpyHostToDevice, st));
  FC_CHECK(fc_calcuvw_dev(ctx, o, rep));
  const struct { int f; double *h; size_t n; } dn[] = {{FC_U, u, NT}, {FC_V, v, NT}, {FC_W, w, NT}, {FC_P, p, NT},
                                                      {FC_APU, apu, n}, {FC_APV, apv, n}, {FC_APW, apw, n}};
  for (auto &t : dn)
    if (t.h) FC_CUDA(cudaMemcpyAsync(t.h, ctx->field[t.f], sizeof(double) * t.n, cudaMemcpyDeviceToHost, st));
  FC_CUDA(cudaStreamSynchronize(st));
  return FC_OK;
}

int fc_set_gradient(fc_context *ctx, int method, int limiter, double small) {
  if (!ctx) return FC_ERR_ARG;
  FC_CUDA(cudaSetDevice(ctx->device));
  return fc_set_gradient_dev(ctx, method, limiter, small);
}

int fc_grad(fc_context *ctx, int phi_field, int grad_field, int nigrad) {
  if (!ctx) return FC_ERR_ARG;
  FC_CUDA(cudaSetDevice(ctx->device));
  FC_CHECK(check_field(ctx, phi_field, (size_t)ctx->NT, "fc_grad"));
  FC_CHECK(check_field(ctx, grad_field, 3 * (size_t)ctx->NP, "fc_grad"));
  if (nigrad < 1) FC_FAIL(FC_ERR_ARG, "fc_grad: nigrad < 1");
  return fc_grad_dev(ctx, ctx->field[phi_field], ctx->field[grad_field], nigrad);
}

int fc_dpcg(fc_context *ctx, int fi_field, const fc_solver_opts *o, fc_solver_report *rep) {
  return fc_solve(ctx, FC_DPCG, fi_field, o, rep);
}
int fc_iccg(fc_context *ctx, int fi_field, const fc_solver_opts *o, fc_solver_report *rep) {
  return fc_solve(ctx, FC_ICCG, fi_field, o, rep);
}
int fc_bicgstab(fc_context *ctx, int fi_field, const fc_solver_opts *o, fc_solver_report *rep) {
  return fc_solve(ctx, FC_BICGSTAB, fi_field, o, rep);
}

int fc_piso(fc_context *ctx, const fc_piso_opts *o, fc_piso_report *rep) {
  if (!ctx || !o || !rep) return FC_ERR_ARG;
  FC_CUDA(cudaSetDevice(ctx->device));
  return fc_piso_dev(ctx, o, rep);
}

int fc_exchange(fc_context *ctx, int field) {
  if (!ctx) return FC_ERR_ARG;
  FC_CHECK(check_field(ctx, field, (size_t)ctx->n + ctx->npro, "fc_exchange"));
  return fc_halo_exchange(ctx, ctx->field[field]);
}

int fc_global_sum(fc_context *ctx, double *value) {
  if (!ctx || !value) return FC_ERR_ARG;
  if (ctx->nranks == 1) return FC_OK;
  FC_CUDA(cudaMemcpyAsync(&ctx->sc->aux[0], value, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  FC_CHECK(fc_allreduce_scalars(ctx, &ctx->sc->aux[0], 1));
  FC_CUDA(cudaMemcpyAsync(value, &ctx->sc->aux[0], sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  return FC_OK;
}

int fc_set_tuning(fc_context *ctx, int key, int value) {
  if (!ctx) return FC_ERR_ARG;
  switch (key) {
    case FC_TUNE_SPMV_KERNEL: if (value < 0 || value > 2) return FC_ERR_ARG; ctx->tune_spmv = value; break;
    case FC_TUNE_DPCG_PERSISTENT: if (value < 0 || value > 1) return FC_ERR_ARG; ctx->tune_persist = value; break;
    case FC_TUNE_PIPE_GEOMETRY: if (value < 0 || value > 4) return FC_ERR_ARG; ctx->tune_pipe = value; break;
    case FC_TUNE_CTAS_PER_SM: if (value < 0 || value > 8) return FC_ERR_ARG; ctx->tune_ctas_per_sm = value; break;
    case FC_TUNE_SWEEP_P2P: if (value < 0 || value > 1) return FC_ERR_ARG; ctx->tune_sweep_p2p = value; break;
    case FC_TUNE_FUSED_GRAD: if (value < 0 || value > 1) return FC_ERR_ARG; ctx->tune_fused_grad = value; break;
    case FC_TUNE_DPCG_FUSED: if (value < 0 || value > 2) return FC_ERR_ARG; ctx->tune_dpcg_fused = value; break;
    case FC_TUNE_FACE_OCC: if (value < 2 || value > 4) return FC_ERR_ARG; ctx->tune_face_occ = value; break;
    case FC_TUNE_JA_CODED: if (value < 0 || value > 2) return FC_ERR_ARG; ctx->tune_ja_coded = value; break;
    case FC_TUNE_DPCG_EAGER: if (value < 0 || value > 2) return FC_ERR_ARG; ctx->tune_dpcg_eager = value; break;
    case FC_TUNE_X_PREFETCH: if (value < 0 || value > 2) return FC_ERR_ARG; ctx->tune_x_prefetch = value; break;
    case FC_TUNE_MAT_KEEP: if (value < -1 || value > 100) return FC_ERR_ARG; ctx->tune_mat_keep = value; break;
    case FC_TUNE_L2_KEEP: if (value < 0 || value > 2) return FC_ERR_ARG; ctx->tune_l2_keep = value; break;
    case FC_TUNE_SWEEP_CHECK: if (value < 0 || value > 1) return FC_ERR_ARG; ctx->tune_sweep_check = value; break;
    case FC_TUNE_TILE_CTAS: if (value < 2 || value > 6) return FC_ERR_ARG; ctx->tune_tile_ctas = value; break;
    case FC_TUNE_SWEEP_TILED: if (value < 0 || value > 5) return FC_ERR_ARG; ctx->tune_sweep_tiled = value; break;
    default: FC_FAIL(FC_ERR_ARG, "fc_set_tuning: unknown key");
  }
  return FC_OK;
}

const char *fc_sweep_schedule_info(fc_context *ctx) {
  if (!ctx) return "";
  if (!ctx->has_levels) { ctx->sweep_info = "no sweep has run on this pattern yet"; return ctx->sweep_info.c_str(); }
  ctx->sweep_info = "level schedule: " + std::to_string(ctx->lower.nlev) + " / " + std::to_string(ctx->upper.nlev) +
                    " row levels (lower / upper triangle)";
  if (ctx->tune_sweep_tiled) {
    if (ctx->tiles_ok) ctx->sweep_info += "; in use: tiled schedule, " + ctx->tiles_info;
    else ctx->sweep_info += "; tiled schedule requested but not used: " + (ctx->tiles_tried ? ctx->tiles_why : std::string("not built yet"));
  }
  return ctx->sweep_info.c_str();
}

int fc_get_timings(const fc_context *ctx, fc_timings *t) {
  if (!ctx || !t) return FC_ERR_ARG;
  *t = ctx->tm;
  t->launches = ctx->launches;
  t->sweep_tiles = (ctx->tune_sweep_tiled && ctx->tiles_ok) ? ctx->tile_lower.nblocks : 0;
  t->column_offsets = ctx->coded_ok ? ctx->ndict : 0;
  return FC_OK;
}

}  // extern "C"
