// Peer-to-peer (NVLink / NVSwitch) communication for the Krylov loop of the cell-partitioned build.
//
// src-parallel runs one blocking exchange and three MPI_ALLREDUCEs of a single double per DPCG
// iteration (src-parallel/dpcg.f90:75,102,114,141).  On 8 B200s an iteration of the 216^3 case is
// ~50 us of HBM traffic per GPU, so three NCCL collectives (~20 us each) plus their scalar kernels cap
// the parallel efficiency near 30 %.  Here the ranks map each other's communication arena with CUDA
// IPC and the kernels talk through it directly:
//   * halo of the search direction: the pack kernel stores every boundary value straight into the
//     neighbour's halo slots of pk / zk (remote stores over NVLink) and raises a per-connection flag;
//     the SpMV only waits for the flag when it reaches a row with processor faces, so the interior
//     rows overlap the transfer;
//   * reductions: the finishing thread of a rank's reduction writes its partial sums into every
//     rank's mailbox, the next kernel adds the mailboxes in rank order (fc_reduce.cuh) -- no
//     collective call, no extra launch, bit-identical scalars on every rank.
// NCCL remains the bootstrap-free fallback and carries the (few) assembly exchanges.
#include "fc_reduce.cuh"

namespace {

struct peer_blob {            // what a rank publishes to its peers (FC_P2P_BLOB_BYTES)
  cudaIpcMemHandle_t ipc;     // 64 bytes
  int rank, n, npro, nconn;
  int nbr_rank[FC_MAX_CONN];
  int nbr_off[FC_MAX_CONN + 1];
  unsigned long long off_mail, off_hflag, off_pk, off_zk;
};
static_assert(sizeof(peer_blob) <= FC_P2P_BLOB_BYTES, "peer blob too large");

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// buffer(i) = phi(bufind(i)) (exchange.f90:48-50) written directly into the neighbour's halo slots;
// the CTA that finishes last raises the arrival flags
__global__ void __launch_bounds__(256)
k_pack_p2p(int npro, const int *__restrict__ bufind, const double *__restrict__ x, const fc_p2p_dev *P, int which,
           unsigned long long seq, const fc_scalars *sc, unsigned int *ticket) {
  __shared__ bool s_last;
  if (((volatile const fc_scalars *)sc)->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npro) {
    int c = 0;
    while (i >= P->conn_off[c + 1]) ++c;
    double *dst = which ? P->peer_zk[c] : P->peer_pk[c];
    dst[i - P->conn_off[c]] = x[bufind[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < P->nconn) {
      __threadfence_system();
      fc_st_release_sys(P->peer_hflag[threadIdx.x], seq);
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

}  // namespace

void fc_p2p_close(fc_context *ctx) {
  for (void *p : ctx->peer_base)
    if (p) cudaIpcCloseMemHandle(p);
  ctx->peer_base.clear();
  if (ctx->p2p_dev) { cudaFree(ctx->p2p_dev); ctx->p2p_dev = nullptr; }
  if (ctx->arena) {
    // pk / zk live inside the arena while P2P is on
    ctx->pk = ctx->zk = nullptr;
    ctx->scratch_n = 0;
    cudaFree(ctx->arena);
    ctx->arena = nullptr;
  }
  ctx->p2p = false;
}

int fc_p2p_pack(fc_context *ctx, double *x) {
  const int which = (x == ctx->zk) ? 1 : 0;
  if (x != ctx->pk && x != ctx->zk) FC_FAIL(FC_ERR_ARG, "P2P halo is only wired for the Krylov vectors");
  ctx->halo_wait = ++ctx->halo_seq;
  k_pack_p2p<<<fc_blocks(ctx->npro, 256), 256, 0, ctx->stream>>>(ctx->npro, ctx->bufind, x, ctx->p2p_dev, which,
                                                               ctx->halo_wait, ctx->sc, &ctx->sc->ticket[3]);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

extern "C" int fc_comm_p2p_blob(fc_context *ctx, char *blob) {
  if (!ctx || !blob) return FC_ERR_ARG;
  if (!ctx->has_csr || !ctx->has_mesh) FC_FAIL(FC_ERR_ARG, "fc_comm_p2p_blob: call fc_set_mesh and fc_create_csr first");
  if (ctx->nranks > FC_MAX_RANKS || (int)ctx->nbr_rank.size() > FC_MAX_CONN)
    FC_FAIL(FC_ERR_UNSUPPORTED, "fc_comm_p2p_blob: too many ranks / connections for the P2P tables");
  FC_CUDA(cudaSetDevice(ctx->device));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  fc_p2p_close(ctx);
  const size_t np = (size_t)ctx->n + ctx->npro;
  peer_blob b;
  memset(&b, 0, sizeof(b));
  b.off_mail = 0;
  b.off_hflag = align256(sizeof(fc_mail) * FC_MAIL_SLOTS * FC_MAX_RANKS);
  b.off_pk = align256(b.off_hflag + sizeof(unsigned long long) * FC_MAX_CONN);
  b.off_zk = align256(b.off_pk + sizeof(double) * np);
  ctx->arena_bytes = align256(b.off_zk + sizeof(double) * np);
  ctx->arena_hflag_off = b.off_hflag;
  FC_CUDA(cudaMalloc(&ctx->arena, ctx->arena_bytes));
  FC_CUDA(cudaMemset(ctx->arena, 0, ctx->arena_bytes));
  // the Krylov vectors whose halos the neighbours write move into the shared arena
  FC_CHECK(fc_alloc_solver_scratch(ctx));
  cudaFree(ctx->pk);
  cudaFree(ctx->zk);
  ctx->pk = (double *)((char *)ctx->arena + b.off_pk);
  ctx->zk = (double *)((char *)ctx->arena + b.off_zk);
  FC_CUDA(cudaIpcGetMemHandle(&b.ipc, ctx->arena));
  b.rank = ctx->rank; b.n = ctx->n; b.npro = ctx->npro; b.nconn = (int)ctx->nbr_rank.size();
  for (int c = 0; c < b.nconn; ++c) b.nbr_rank[c] = ctx->nbr_rank[c];
  for (int c = 0; c <= b.nconn; ++c) b.nbr_off[c] = ctx->nbr_off[c];
  memset(blob, 0, FC_P2P_BLOB_BYTES);
  memcpy(blob, &b, sizeof(b));
  return FC_OK;
}

extern "C" int fc_comm_p2p_open(fc_context *ctx, const char *blobs, int nranks) {
  if (!ctx || !blobs) return FC_ERR_ARG;
  if (!ctx->arena) FC_FAIL(FC_ERR_ARG, "fc_comm_p2p_open: call fc_comm_p2p_blob first");
  if (nranks != ctx->nranks) FC_FAIL(FC_ERR_ARG, "fc_comm_p2p_open: nranks differs from fc_comm_init");
  FC_CUDA(cudaSetDevice(ctx->device));
  std::vector<peer_blob> B(nranks);
  for (int r = 0; r < nranks; ++r) memcpy(&B[r], blobs + (size_t)r * FC_P2P_BLOB_BYTES, sizeof(peer_blob));
  if (B[ctx->rank].rank != ctx->rank) FC_FAIL(FC_ERR_ARG, "fc_comm_p2p_open: blobs are not ordered by rank");
  ctx->peer_base.assign(nranks, nullptr);
  std::vector<char *> base(nranks, nullptr);
  for (int r = 0; r < nranks; ++r) {
    if (r == ctx->rank) { base[r] = (char *)ctx->arena; continue; }
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, B[r].ipc, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      ctx->err = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e);
      cudaGetLastError();
      fc_p2p_close(ctx);
      return FC_ERR_CUDA;
    }
    ctx->peer_base[r] = p;
    base[r] = (char *)p;
  }
  fc_p2p_dev h;
  memset(&h, 0, sizeof(h));
  const peer_blob &me = B[ctx->rank];
  h.rank = ctx->rank; h.nranks = nranks; h.nconn = me.nconn;
  h.mail = (fc_mail *)(base[ctx->rank] + me.off_mail);
  h.hflag = (unsigned long long *)(base[ctx->rank] + me.off_hflag);
  for (int r = 0; r < nranks; ++r) h.peer_mail[r] = (fc_mail *)(base[r] + B[r].off_mail);
  for (int c = 0; c <= me.nconn; ++c) h.conn_off[c] = me.nbr_off[c];
  for (int c = 0; c < me.nconn; ++c) {
    const int q = me.nbr_rank[c];
    int cq = -1;   // the neighbour's connection that points back at me (exchange.f90: mirrored slices pair up)
    for (int k = 0; k < B[q].nconn; ++k)
      if (B[q].nbr_rank[k] == ctx->rank) cq = k;
    if (cq < 0 || B[q].nbr_off[cq + 1] - B[q].nbr_off[cq] != me.nbr_off[c + 1] - me.nbr_off[c]) {
      fc_p2p_close(ctx);
      FC_FAIL(FC_ERR_ARG, "fc_comm_p2p_open: connection tables of two ranks do not mirror each other");
    }
    const size_t halo = (size_t)B[q].n + B[q].nbr_off[cq];
    h.peer_pk[c] = (double *)(base[q] + B[q].off_pk) + halo;
    h.peer_zk[c] = (double *)(base[q] + B[q].off_zk) + halo;
    h.peer_hflag[c] = (unsigned long long *)(base[q] + B[q].off_hflag) + cq;
  }
  FC_CHECK(fc_dev_alloc(ctx, &ctx->p2p_dev, 1));
  FC_CUDA(cudaMemcpy(ctx->p2p_dev, &h, sizeof(h), cudaMemcpyHostToDevice));
  ctx->p2p = true;
  ctx->pending = {0, 0, 0};
  return FC_OK;
}
