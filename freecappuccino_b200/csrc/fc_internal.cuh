// Internal declarations of libfcapp_cuda (not installed; the public ABI is include/fcapp.h).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fcapp.h"

// Number of SMs the persistent / grid-stride kernels are sized for (B200).
constexpr int FC_SMS = 148;
constexpr int FC_MAX_DEVICES = 64;               // per-device caches of kernel attributes (one process may hold several contexts)
constexpr int FC_RED_BLOCK = 512;               // threads of the streaming vector kernels
constexpr int FC_RED_GRID = FC_SMS * 4;         // 4 resident CTAs of 512 threads per SM
constexpr int FC_MAX_RED = 4;                   // scalars reduced by one kernel

// Scalars of a Krylov solve, resident on the device so that no iteration
// needs a host round trip (dpcg.f90:33, bicgstab.f90:26-28).
struct fc_scalars {
  double res0, resl;
  double sk, s0, pkapk;                 // dpcg / iccg
  double bet, beto, alf, gam, om;       // bicgstab (bicgstab.f90:105-113)
  double ukreso, svkres, svkvk;
  double sor, small;
  double red[FC_MAX_RED];               // raw (rank-local, then all-reduced) sums of the last reduction
  double aux[4];
  int iters, nsw, done, pad;
  unsigned int ticket[4];               // "last block finalises" counters
  unsigned long long applied;           // P2P mode: sequence number of the last reduction folded into the scalars
};

// ---- peer-to-peer (NVLink) communication state, see fc_p2p.cu ----
constexpr int FC_MAX_RANKS = 16, FC_MAX_CONN = 16, FC_MAIL_SLOTS = 4;
// One rank's contribution to one reduction, written by that rank straight into every rank's copy.  Every FP64
// value travels as two 8-byte words {32 bits of the value | 32-bit sequence number}: an 8-byte store is atomic,
// so a word whose sequence half matches is complete and no fence or separate flag is needed between data and
// "ready" (one NVLink flight per reduction instead of store - fence - flag).
struct fc_mail {
  unsigned long long w[2 * FC_MAX_RED];
};
struct fc_p2p_dev {                     // device-resident tables read by the kernels
  int rank, nranks, nconn, pad;
  fc_mail *mail;                                   // my mailboxes [FC_MAIL_SLOTS][FC_MAX_RANKS]
  fc_mail *peer_mail[FC_MAX_RANKS];                // every rank's mailbox array (peer-mapped; [rank] = mine)
  unsigned long long *hflag;                       // my halo-arrival flags, one per connection
  unsigned long long *peer_hflag[FC_MAX_CONN];     // the flag in neighbour c's arena that I raise
  double *peer_pk[FC_MAX_CONN], *peer_zk[FC_MAX_CONN];  // halo slots of neighbour c's pk / zk that I fill
  int conn_off[FC_MAX_CONN + 1];                   // my processor faces [conn_off[c], conn_off[c+1]) go to c
};
struct fc_sync {                        // per-launch synchronisation descriptor of a Krylov kernel
  fc_p2p_dev *p2p;                      // nullptr: single rank, or NCCL mode
  unsigned long long wait_seq;          // reduction to fold into the scalars before the kernel body (0: none)
  int wait_step, wait_count;
  unsigned long long post_seq;          // sequence number of the reduction this kernel produces
  int local, pad;                       // 1: single rank, the finalising thread runs the scalar step itself
  double *hist;
};

// grid barrier and phase clocks of the persistent DPCG kernel (fc_dpcg_persist.cu)
struct fc_persist_state {
  unsigned int count;                   // arrivals at the current barrier
  unsigned int pad0;
  unsigned long long gen;               // barriers completed so far
  unsigned long long t_mark;            // globaltimer at the last barrier release
  unsigned long long t_phase[4];        // accumulated ns: p-update, SpMV, x/r update, (spare)
  unsigned long long t_total;           // ns from kernel start to the last release
  unsigned long long t_start;
  unsigned long long t_mail;            // ns the reducing CTA waited for the other ranks' partial sums
  unsigned long long pad1[5];
};

struct fc_levels {                      // level schedule of the strict lower / upper triangle
  int nlev = 0;
  int nslots = 0;                       // rows padded so that a block never straddles two levels
  int *rows = nullptr;                  // [nslots] row id or -1 (padding), level-major, ascending inside a level
  int *blk_level = nullptr;             // [nslots / TRI_BLOCK] level of every block
  int *lev_blocks_before = nullptr;     // [nlev+1] number of blocks in levels < L
  unsigned int *done = nullptr;         // [nlev] blocks finished per level (monotone over sweeps)
  unsigned int *ready = nullptr;        // [nlev] number of the last sweep whose level is complete
  unsigned int *ticket = nullptr;       // dynamic block id
  unsigned long long epoch = 0;         // sweeps run so far
  // point-to-point mode (FC_TUNE_SWEEP_P2P): a block waits for the flags of the blocks it gathers from instead of
  // for the whole previous level
  int nblocks = 0;
  int *prod = nullptr;                  // [nblocks * FC_TRI_MAXP] producer blocks
  int *prod_cnt = nullptr;              // [nblocks]
  unsigned int *flag = nullptr;         // [nblocks] number of the last sweep whose rows the block has published
  bool p2p_ok = false;                  // every block has <= FC_TRI_MAXP producers
  // tiled mode (FC_TUNE_SWEEP_TILED, fc_tile_schedule.hpp): a block = one spatial tile of FC_TILE slots, `nlev` etc.
  // count tile levels, and inside the tile the rows are walked by local level
  int4 *meta_rm = nullptr;              // [nslots] the same in ascending row order: row, slot | level << 16, [s, e) (k_tile_walk)
  int4 *meta = nullptr;                 // [nslots] per slot: row (-1 = padding), local level, triangle range [s, e) in a / tja
  int *blk_nlev = nullptr;              // [nblocks] local levels of the tile
};
constexpr int FC_TRI_MAXP = 16;

struct fc_context {
  int device = 0;
  int sms = FC_SMS;                     // SM count of `device` (queried in fc_create)
  int l2_bytes = 0;                     // L2 size of `device`
  double *hist = nullptr;               // residual history of a solve (fc_solve_csr's `hist`), grown on demand
  size_t hist_cap = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  long long launches = 0;

  // ---- mesh (0-based on the device) ----
  bool has_mesh = false;
  fc_mesh_desc m{};                     // scalar members only are meaningful
  int n = 0, F = 0, NF = 0, NT = 0, npro = 0, nnz = 0, NP = 0;
  int *owner = nullptr, *neigh = nullptr;
  double *xc = nullptr, *yc = nullptr, *zc = nullptr, *vol = nullptr;
  double *arx = nullptr, *ary = nullptr, *arz = nullptr, *xf = nullptr, *yf = nullptr, *zf = nullptr;
  double *facint = nullptr, *fpro = nullptr;
  // cell -> face map in the reference's loop order (inner faces ascending, processor
  // faces, then inlet, outlet, symmetry, wall, prOutlet): grad_gauss.f90:53-103
  int *c2f_off = nullptr;               // [n+1]
  int *c2f_face = nullptr;              // face id | (cell is the neighbour side) << 31
  int *c2f_other = nullptr;             // other cell (inner), halo / boundary slot otherwise
  int *c2f_pos = nullptr;               // CSR position of a(cell,other) for inner faces, else -1
  int c2f_len = 0;

  // ---- CSR pattern (0-based on the device) ----
  bool has_csr = false;
  bool csr_external = false;            // adopted through fc_solve_csr (no mesh)
  bool csr_dup = false;                 // duplicate cell pairs: `a` must be zeroed before assembly
  int *ioffset = nullptr, *ja = nullptr, *diag = nullptr, *icj = nullptr, *jci = nullptr;
  int *tpos = nullptr;                  // position of a(j,i) for every lower a(i,j) (bicgstab.f90:72-75)
  int spmv_max_chunk = 0;               // max nnz of a 256-row block
  unsigned char *jcode = nullptr;       // [nnz] one-byte column codes, ja[k] = row + jdict[jcode[k]] (fc_codes_build)
  int *jdict = nullptr;                 // [256] ascending column offsets
  int ndict = 0;
  bool coded_ok = false;                // the pattern has <= 256 distinct column offsets: jcode / jdict are valid

  // ---- fields ----
  double *field[FC_NUM_FIELDS] = {};
  size_t field_n[FC_NUM_FIELDS] = {};

  // ---- solver scratch ----
  double *pk = nullptr, *zk = nullptr, *dd = nullptr, *reso = nullptr, *uk = nullptr, *vk = nullptr;
  double *adiag = nullptr;              // a(diag(i)) compacted once per solve
  double *tt = nullptr;                 // forward-sweep result of the preconditioner
  double *coef = nullptr;               // per-face coefficient cap = can [F + npro]
  double *facev = nullptr;              // per-face scratch [NF]
  double *gtmp = nullptr;               // previous-pass gradient (3,numCells)
  double *gtmp3 = nullptr;              // the same for the fused u, v, w pass (FC_TUNE_FUSED_GRAD, nigrad > 1)
  // gradient scheme of the `grad` dispatcher (fc_set_gradient): 0 gauss, 1 lstsq, 2 lstsq_qr, 3 lstsq_dm; limiter 0..3
  int grad_method = 0, grad_limiter = 0;
  double grad_small = 0.0;
  double *dmat = nullptr;               // (9,numCells) of grad_lsq / grad_lsq_dm
  double *dmatqr = nullptr;             // (3,6,numCells) R^-1 Q^T of grad_lsq_qr
  double *hcoef = nullptr;              // PISO: h = a, the momentum matrix backed up before the correctors [nnz]
  double *uvw_face = nullptr;           // momentum predictor: can, cap, sup, svp, swp, fie per inner face [6 F]
  double *partials = nullptr;           // [FC_MAX_RED * FC_RED_GRID]
  fc_scalars *sc = nullptr;             // device
  fc_scalars *sc_host = nullptr;        // pinned
  size_t scratch_n = 0;
  fc_levels lower, upper;
  bool has_levels = false;
  fc_levels tile_lower, tile_upper;     // tiled schedule of the same sweeps (only when tune_sweep_tiled)
  int *tja = nullptr;                   // [nnz] column, or -(slot+1) when the column's row sits in the same tile
  bool tiles_tried = false, tiles_ok = false, tiles_pre8 = false, tiles_pre3 = false;
  std::string tiles_why;                // why the mesh got no tiling
  std::string tiles_info;               // tiles, tile levels, local levels of the tiling in use
  std::string sweep_info;               // fc_sweep_schedule_info's answer

  // ---- communication ----
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  std::vector<int> nbr_rank, nbr_off;   // neighbProcNo, neighbProcOffset (0-based offsets)
  int *bufind = nullptr;                // owner cell of every processor face (exchange.f90:48-50)
  double *sendbuf = nullptr;
  int *strip_off = nullptr, *strip_idx = nullptr;  // per-row processor faces (apr strip of the SpMV)
  unsigned char *strip_any32 = nullptr;            // [ceil(n/32)] 1 when one of the 32 rows has processor faces
  // peer-to-peer mode (fc_p2p.cu): reductions and the Krylov halo go over mapped peer memory
  bool p2p = false;
  void *arena = nullptr;                // IPC-shared allocation: mailboxes, flags, pk, zk
  size_t arena_bytes = 0, arena_hflag_off = 0;
  fc_p2p_dev *p2p_dev = nullptr;
  std::vector<void *> peer_base;        // opened peer arenas
  unsigned long long red_seq = 0, halo_seq = 0, halo_wait = 0;
  struct { unsigned long long seq; int step, count; } pending = {0, 0, 0};

  // ---- tuning (fc_set_tuning) ----
  int tune_spmv = 2;                    // 0: CSR-stream kernel, 1: TMA-staged pipeline, 2: by size (measured cross-over)
  int tune_persist = 1;                 // 1: DPCG as one persistent cooperative kernel
  int tune_pipe = 1;                    // staging geometry of the TMA pipeline (threads, capacity, stages)
  int tune_ctas_per_sm = 0;             // persistent kernel: CTAs per SM (0 = as many as fit)
  int tune_sweep_p2p = 0;               // triangular sweeps: 1 = point-to-point block flags instead of level counters
  int first_batch[3] = {8, 8, 8};       // per solver: iterations enqueued before the first look at `done`
  int tune_dpcg_fused = 0;              // persistent DPCG: fused-p scheme (0 never [default: measured slower, profiles/r02_fused_p.txt], 1 always, 2 on partitioned meshes)
  int tune_face_occ = 3;                // k_calcp_faces: CTAs per SM its registers must allow (2: 102 registers, 3: 80, 4: 64 + spills)
  int tune_mat_keep = -1;               // persistent DPCG: percent of the matrix chunks kept in L2 (evict_last); -1 = by size
  int tune_dpcg_eager = 2;              // persistent DPCG: x update behind the beta reduction + q hand-over (0 off, 1 on, 2 when the vectors fit the L2)
  int tune_x_prefetch = 0;              // persistent DPCG: prefetch the next chunk's far x gathers (0 off, 1 into L1, 2 into L2)
  int tune_ja_coded = 2;                // persistent DPCG: one-byte column codes instead of `ja` where the pattern allows (0 off, 1 on, 2 from 2 M rows)
  int tune_l2_keep = 2;                 // persistent DPCG: Krylov vectors evict_last in L2 (0 off, 1 on, 2 when they fit)
  int tune_fused_grad = 1;              // 1: the three velocity gradients of calcuvw / calcp in one kernel per pass
  int tune_sweep_check = 0;             // debugging: every tiled sweep is repeated with the level schedule and compared
  double *vf_armed = nullptr;           // value-as-flag sweeps: the vector currently known to be all "unset"
  double *sweep_chk = nullptr;          // [n + 2] scratch of that comparison (+ two counters)
  int tune_tile_ctas = 2;               // tiled sweeps: CTAs per SM the kernel is compiled for (2 or 3)
  int tune_sweep_tiled = 4;             // triangular sweeps: 1 = two-level tiled schedule where the mesh allows it
  fc_persist_state *persist = nullptr;  // device: grid barrier + phase clocks of the persistent kernel
  fc_persist_state *persist_host = nullptr;

  // ---- timing ----
  cudaEvent_t ev[4] = {};
  std::vector<cudaEvent_t> spmv_ev;     // 2 * max_samples events bracketing SpMV launches of a solve
  int spmv_sampled = 0;
  fc_timings tm{};
};

#define FC_CUDA(call)                                                                           \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +   \
                 std::to_string(__LINE__) + ")";                                                \
      return FC_ERR_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

#define FC_NCCL(call)                                                                           \
  do {                                                                                          \
    ncclResult_t r_ = (call);                                                                   \
    if (r_ != ncclSuccess) {                                                                    \
      ctx->err = std::string(#call) + ": " + ncclGetErrorString(r_);                            \
      return FC_ERR_NCCL;                                                                       \
    }                                                                                           \
  } while (0)

#define FC_CHECK(expr)                                                                          \
  do {                                                                                          \
    int s_ = (expr);                                                                            \
    if (s_ != FC_OK) return s_;                                                                 \
  } while (0)

#define FC_FAIL(code, msg)                                                                      \
  do {                                                                                          \
    ctx->err = (msg);                                                                           \
    return (code);                                                                              \
  } while (0)

#define FC_LAUNCH_CHECK()                                                                       \
  do {                                                                                          \
    ctx->launches++;                                                                            \
    FC_CUDA(cudaGetLastError());                                                                \
  } while (0)

template <class T>
static inline int fc_dev_alloc(fc_context *ctx, T **p, size_t count) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (count == 0) count = 1;
  // 64 spare bytes: the bulk copies of the SpMV pipeline read whole 16-byte groups
  FC_CUDA(cudaMalloc((void **)p, count * sizeof(T) + 64));
  return FC_OK;
}

static inline int fc_blocks(size_t n, int bs) { return (int)((n + bs - 1) / bs); }

// ---- cross-file entry points ----
int fc_csr_build(fc_context *ctx);                                   // fc_csr.cu
int fc_codes_build(fc_context *ctx);                                 // one-byte column codes of the pattern
int fc_csr_post(fc_context *ctx);                                    // transposed positions, spmv chunking
int fc_c2f_build(fc_context *ctx);                                   // fc_csr.cu
int fc_levels_build(fc_context *ctx);                                // fc_trisolve.cu
void fc_levels_free(fc_levels &L);
int fc_levels_reset(fc_context *ctx);
int fc_precond_factor(fc_context *ctx, int kind, const double *a, double *d, double padd);
int fc_precond_apply(fc_context *ctx, const double *a, const double *d, const double *r, double *t, double *z,
                     double small);
int fc_alloc_solver_scratch(fc_context *ctx);                        // fc_krylov.cu
int fc_launch_spmv(fc_context *ctx, const double *a, const double *x, double *y);       // fc_spmv.cu
int fc_launch_spmv_dots(fc_context *ctx, const double *a, const double *x, double *y, const double *w, int two,
                        int step, const fc_sync &sy);
int fc_launch_residual(fc_context *ctx, const double *a, const double *su, const double *x, double *res,
                       double *adiag, const fc_sync &sy);
void fc_comm_destroy(fc_context *ctx);                               // fc_comm.cu
int fc_halo_exchange(fc_context *ctx, double *phi);
int fc_halo_exchange3(fc_context *ctx, double *grad);                // interleaved (3,numPCells) field
int fc_strip_build(fc_context *ctx);
int fc_p2p_pack(fc_context *ctx, double *x);                          // fc_p2p.cu: halo of pk / zk by peer stores
void fc_p2p_close(fc_context *ctx);                                 // fc_csr.cu: per-row processor faces
int fc_allreduce_scalars(fc_context *ctx, double *dev, int count);
int fc_allreduce_max(fc_context *ctx, double *dev, int count);   // element-wise maximum over the ranks
int fc_dpcg_persistent(fc_context *ctx, double *fi, const fc_solver_opts *o, fc_solver_report *rep, double *hist,
                       bool *handled);                                                   // fc_dpcg_persist.cu
int fc_solve_device(fc_context *ctx, int solver, double *fi, const fc_solver_opts *o, fc_solver_report *rep,
                    double *hist);
int fc_momentum_fields(fc_context *ctx);                            // fc_capi.cu: FC_VIS.. + face scratch, first use
int fc_calcuvw_assemble_dev(fc_context *ctx, const fc_calcuvw_opts *o);                   // fc_momentum.cu
int fc_calcuvw_component_dev(fc_context *ctx, const fc_calcuvw_opts *o, int comp, fc_solver_report *rep);
int fc_calcuvw_dev(fc_context *ctx, const fc_calcuvw_opts *o, fc_calcuvw_report *rep);
int fc_piso_dev(fc_context *ctx, const fc_piso_opts *o, fc_piso_report *rep);             // fc_assemble.cu
int fc_grad_dev(fc_context *ctx, double *phi, double *grad, int nigrad);                   // fc_gradients.cu
int fc_grad_uvw_dev(fc_context *ctx, int nigrad, bool with_p_stage1);                                              // fc_gradients.cu: grad(U), grad(V), grad(W)
int fc_limit_gradient_dev(fc_context *ctx, const double *phi, double *grad);
int fc_set_gradient_dev(fc_context *ctx, int method, int limiter, double small);
