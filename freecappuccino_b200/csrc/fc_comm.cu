// Communication of the cell-partitioned build: the NCCL equivalents of src-parallel's
// exchange (exchange.f90:3-92) and global_sum (global_sum_mpi.f90:4-37).
//
// One rank per GPU.  A halo exchange is a device pack (buffer(i) = phi(bufind(i)),
// exchange.f90:48-50) followed by ONE ncclGroup holding a send/recv pair per neighbouring
// rank; the receive lands directly in the halo slots phi(iProcStart+i) (:86-90), so there
// is no unpack kernel.  Scalar sums are an in-place ncclAllReduce on the device-resident
// reduction block; everything is enqueued on the context's stream.
#include <dlfcn.h>

#include "fc_internal.cuh"

// NCCL is bound at run time, on the first communicator call, and never at library load: a host
// process may already hold a libnccl (torch bundles its own 2.28 next to the system 2.27) and two
// copies with one soname must not race for the symbol namespace.  An already-loaded libnccl.so.2 is
// reused (RTLD_NOLOAD); otherwise the system library is opened.
namespace fcnccl {
#define FC_NCCL_FUNCS(X)                                                                         \
  X(ncclGetUniqueId) X(ncclCommInitRank) X(ncclCommDestroy) X(ncclSend) X(ncclRecv) X(ncclAllReduce) \
  X(ncclGroupStart) X(ncclGroupEnd) X(ncclGetErrorString)
#define X(f) static decltype(&::f) f = nullptr;
FC_NCCL_FUNCS(X)
#undef X
static bool bind() {
  static int state = 0;  // 0 untried, 1 ok, -1 failed
  if (state) return state > 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  bool ok = h != nullptr;
#define X(f) if (ok) { f = (decltype(f))dlsym(h, #f); ok = f != nullptr; }
  FC_NCCL_FUNCS(X)
#undef X
  state = ok ? 1 : -1;
  return ok;
}
}  // namespace fcnccl
#define ncclGetUniqueId fcnccl::ncclGetUniqueId
#define ncclCommInitRank fcnccl::ncclCommInitRank
#define ncclCommDestroy fcnccl::ncclCommDestroy
#define ncclSend fcnccl::ncclSend
#define ncclRecv fcnccl::ncclRecv
#define ncclAllReduce fcnccl::ncclAllReduce
#define ncclGroupStart fcnccl::ncclGroupStart
#define ncclGroupEnd fcnccl::ncclGroupEnd
#define ncclGetErrorString fcnccl::ncclGetErrorString

void fc_comm_destroy(fc_context *ctx) {
  if (ctx->comm && fcnccl::bind()) ncclCommDestroy(ctx->comm);
  ctx->comm = nullptr;
}

namespace {
__global__ void k_pack(int npro, const int *__restrict__ bufind, const double *__restrict__ phi,
                       double *__restrict__ buf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npro) buf[i] = phi[bufind[i]];
}
__global__ void k_pack3(int npro, const int *__restrict__ bufind, const double *__restrict__ g,
                        double *__restrict__ buf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 3 * npro) buf[t] = g[3 * (size_t)bufind[t / 3] + (t % 3)];
}
}  // namespace

// exchange( dPhidxi(1,:) ), (2,:), (3,:) of src-parallel/gradients.f90:157-159 as ONE exchange: the
// halo part of the interleaved (3,numPCells) array is contiguous, so the receive needs no unpack
int fc_halo_exchange3(fc_context *ctx, double *grad) {
  if (ctx->npro == 0) return FC_OK;
  if (!ctx->comm) FC_FAIL(FC_ERR_ARG, "halo exchange on a partitioned mesh needs fc_comm_init");
  k_pack3<<<fc_blocks(3 * (size_t)ctx->npro, 256), 256, 0, ctx->stream>>>(ctx->npro, ctx->bufind, grad, ctx->sendbuf);
  FC_LAUNCH_CHECK();
  FC_NCCL(ncclGroupStart());
  for (size_t c = 0; c < ctx->nbr_rank.size(); ++c) {
    const int off = ctx->nbr_off[c], len = ctx->nbr_off[c + 1] - off;
    FC_NCCL(ncclSend(ctx->sendbuf + 3 * (size_t)off, 3 * (size_t)len, ncclDouble, ctx->nbr_rank[c], ctx->comm,
                     ctx->stream));
    FC_NCCL(ncclRecv(grad + 3 * ((size_t)ctx->n + off), 3 * (size_t)len, ncclDouble, ctx->nbr_rank[c], ctx->comm,
                     ctx->stream));
  }
  FC_NCCL(ncclGroupEnd());
  return FC_OK;
}

int fc_halo_exchange(fc_context *ctx, double *phi) {
  if (ctx->npro == 0) return FC_OK;
  if (!ctx->comm) FC_FAIL(FC_ERR_ARG, "halo exchange on a partitioned mesh needs fc_comm_init");
  k_pack<<<fc_blocks(ctx->npro, 256), 256, 0, ctx->stream>>>(ctx->npro, ctx->bufind, phi, ctx->sendbuf);
  FC_LAUNCH_CHECK();
  FC_NCCL(ncclGroupStart());
  for (size_t c = 0; c < ctx->nbr_rank.size(); ++c) {
    const int off = ctx->nbr_off[c], len = ctx->nbr_off[c + 1] - off;
    FC_NCCL(ncclSend(ctx->sendbuf + off, (size_t)len, ncclDouble, ctx->nbr_rank[c], ctx->comm, ctx->stream));
    FC_NCCL(ncclRecv(phi + ctx->n + off, (size_t)len, ncclDouble, ctx->nbr_rank[c], ctx->comm, ctx->stream));
  }
  FC_NCCL(ncclGroupEnd());
  return FC_OK;
}

int fc_allreduce_scalars(fc_context *ctx, double *dev, int count) {
  if (ctx->nranks == 1) return FC_OK;
  if (!ctx->comm) FC_FAIL(FC_ERR_ARG, "all-reduce needs fc_comm_init");
  FC_NCCL(ncclAllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, ctx->comm, ctx->stream));
  return FC_OK;
}

// element-wise maximum over the ranks (global_max of src-parallel; a minimum travels negated)
int fc_allreduce_max(fc_context *ctx, double *dev, int count) {
  if (ctx->nranks == 1) return FC_OK;
  if (!ctx->comm) FC_FAIL(FC_ERR_ARG, "all-reduce needs fc_comm_init");
  FC_NCCL(ncclAllReduce(dev, dev, (size_t)count, ncclDouble, ncclMax, ctx->comm, ctx->stream));
  return FC_OK;
}

extern "C" int fc_comm_unique_id(char id128[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (!fcnccl::bind()) return FC_ERR_NCCL;
  if (ncclGetUniqueId(&id) != ncclSuccess) return FC_ERR_NCCL;
  memcpy(id128, &id, 128);
  return FC_OK;
}

extern "C" int fc_comm_init(fc_context *ctx, int rank, int nranks, const char id128[128]) {
  if (!ctx) return FC_ERR_ARG;
  if (nranks < 1 || rank < 0 || rank >= nranks) FC_FAIL(FC_ERR_ARG, "fc_comm_init: bad rank / nranks");
  FC_CUDA(cudaSetDevice(ctx->device));
  fc_comm_destroy(ctx);
  // a mesh set before the communicator: its neighbour table is checked here (a self or out-of-range peer would
  // hang the first ncclSend / ncclRecv pair)
  for (size_t c = 0; c < ctx->nbr_rank.size(); ++c)
    if (nranks > 1 && (ctx->nbr_rank[c] < 0 || ctx->nbr_rank[c] >= nranks || ctx->nbr_rank[c] == rank))
      FC_FAIL(FC_ERR_ARG, "fc_comm_init: neighbProcNo(" + std::to_string(c + 1) + ") = " + std::to_string(ctx->nbr_rank[c]) +
                              " of the mesh is not another rank of the communicator");
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return FC_OK;
  if (!fcnccl::bind()) FC_FAIL(FC_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : ""));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  FC_NCCL(ncclCommInitRank(&ctx->comm, nranks, id, rank));
  return FC_OK;
}
