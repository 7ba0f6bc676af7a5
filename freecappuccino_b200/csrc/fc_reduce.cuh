// Deterministic grid reduction ("last block finalises") and the scalar steps of the
// Krylov recurrences.  No floating-point atomics: every block writes its partial sum,
// the block that draws the last ticket adds the partials in a fixed order, so a given
// (grid, block) configuration always produces the same bits.
#pragma once
#include "fc_internal.cuh"
#include "fc_tma.cuh"

// Scalar recurrences executed once per reduction, either by the finalising thread
// (single GPU) or by k_scalar_step after the NCCL all-reduce (src-parallel: every
// sum is followed by global_sum, dpcg.f90:75,102,141,154 of src-parallel).
enum fc_step {
  STEP_NONE = 0,
  STEP_RES0,        // red[0] = sum|res|            -> res0, resl; reset recurrences (dpcg.f90:64-79)
  STEP_RES0_SK,     // + red[1] = sum res*z          -> sk (first Jacobi inner product fused in)
  STEP_SK,          // red[0] = sum res*z            -> sk (dpcg.f90:90)
  STEP_PKAPK,       // red[0] = sum pk*A pk          -> pkapk (dpcg.f90:119)
  STEP_CG_UPDATE,   // red[0] = sum|res|             -> resl, s0 = sk, ++iters, convergence (dpcg.f90:130-142)
  STEP_CG_UPDATE_SK,// + red[1] = next sum res*z     -> sk
  STEP_BET,         // red[0] = res.reso             -> bet, om, beto (bicgstab.f90:105-113)
  STEP_UKRESO,      // red[0] = uk.reso              -> gam (bicgstab.f90:152-157)
  STEP_VK,          // red[0] = vk.res, red[1]=vk.vk -> alf (bicgstab.f90:200-206)
  STEP_BI_UPDATE    // red[0] = sum|res|             -> resl, ++iters, convergence (bicgstab.f90:217-225)
};

__device__ __forceinline__ void fc_scalar_step(fc_scalars *sc, int step, double *hist) {
  switch (step) {
    case STEP_RES0:
    case STEP_RES0_SK:
      sc->res0 = sc->red[0];
      sc->resl = sc->red[0];
      sc->s0 = (double)1.e20f;  // s0=1.e20 is a default-real literal (dpcg.f90:79)
      sc->iters = 0;
      sc->done = 0;
      sc->alf = 1.0; sc->beto = 1.0; sc->gam = 1.0;  // bicgstab.f90:84-86
      if (step == STEP_RES0_SK) sc->sk = sc->red[1];
      break;
    case STEP_SK:
      sc->sk = sc->red[0];
      break;
    case STEP_PKAPK:
      sc->pkapk = sc->red[0];
      break;
    case STEP_CG_UPDATE:
    case STEP_CG_UPDATE_SK: {
      sc->resl = sc->red[0];
      sc->s0 = sc->sk;
      if (step == STEP_CG_UPDATE_SK) sc->sk = sc->red[1];
      int it = ++sc->iters;
      if (hist) hist[it - 1] = sc->resl;
      double rsm = sc->resl / (sc->res0 + sc->small);
      if (rsm < sc->sor || it >= sc->nsw) sc->done = 1;
    } break;
    case STEP_BET:
      sc->bet = sc->red[0];
      sc->om = sc->bet * sc->gam / (sc->alf * sc->beto + sc->small);
      sc->beto = sc->bet;
      break;
    case STEP_UKRESO:
      sc->ukreso = sc->red[0];
      sc->gam = sc->bet / sc->ukreso;
      break;
    case STEP_VK:
      sc->svkres = sc->red[0];
      sc->svkvk = sc->red[1];
      sc->alf = sc->svkres / (sc->svkvk + sc->small);
      break;
    case STEP_BI_UPDATE: {
      sc->resl = sc->red[0];
      int it = ++sc->iters;
      if (hist) hist[it - 1] = sc->resl;
      double rsm = sc->resl / (sc->res0 + sc->small);
      if (rsm < sc->sor || it >= sc->nsw) sc->done = 1;
    } break;
    default:
      break;
  }
}

__device__ __forceinline__ double fc_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NR values; result valid in thread 0.  blockDim.x must be a multiple of 32, <= 1024.
template <int NR>
__device__ __forceinline__ void fc_block_sum(double (&v)[NR], double *smem /* [NR*32] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double w = fc_warp_sum(v[r]);
    if (lane == 0) smem[r * 32 + wid] = w;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double w = lane < nw ? smem[r * 32 + lane] : 0.0;
      v[r] = fc_warp_sum(w);
    }
  }
  __syncthreads();
}

// Grid-wide deterministic sum.  Every thread passes its private partial sums; after the
// call, thread 0 of exactly one block (the last to arrive) gets `true` with the totals
// in v[].  `partials` holds NR * gridDim.x doubles; `ticket` must be zero on entry and is
// reset on exit.
template <int NR>
__device__ __forceinline__ bool fc_grid_sum(double (&v)[NR], double *partials, unsigned int *ticket,
                                            double *smem /* [NR*32] */) {
  __shared__ bool s_last;
  fc_block_sum<NR>(v, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) partials[r * gridDim.x + blockIdx.x] = v[r];
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double w = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
      w += ((volatile double *)partials)[r * gridDim.x + b];
    v[r] = w;
  }
  fc_block_sum<NR>(v, smem);
  if (threadIdx.x == 0) *ticket = 0u;
  return threadIdx.x == 0;
}


// ---------------------------------------------------------------------------------------------
// Peer-to-peer reductions (multi-GPU without a collective call): the thread that finishes a
// rank's reduction writes the partial sums into EVERY rank's mailbox over NVLink (fc_mail_post);
// the next kernel starts by adding the mailboxes in rank order -- the same order on every rank,
// so all ranks hold bit-identical scalars -- and runs the scalar step (fc_kernel_begin).  This
// replaces ncclAllReduce + a scalar kernel (2 launches, ~20 us) by a few remote 8-byte stores.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long fc_ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fc_st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long fc_ld_acquire_gpu(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fc_st_release_gpu(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void fc_st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long fc_ld_relaxed_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// post `count` partial sums of reduction `seq` to every rank (my own copy included)
__device__ __forceinline__ void fc_mail_post(const fc_p2p_dev *P, unsigned long long seq, const double *v, int count) {
  const int slot = (int)(seq % FC_MAIL_SLOTS);
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  for (int i = 0; i < count; ++i) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v[i]);
    const unsigned long long lo = (bits & 0xffffffffull) | tag, hi = (bits >> 32) | tag;
    for (int q = 0; q < P->nranks; ++q) {
      unsigned long long *dst = P->peer_mail[q][slot * FC_MAX_RANKS + P->rank].w;
      fc_st_relaxed_sys(dst + 2 * i, lo);
      fc_st_relaxed_sys(dst + 2 * i + 1, hi);
    }
  }
}

// wait for every rank's partial sums of reduction `seq` and add them in rank order (global_sum of src-parallel;
// the same order on every rank, so all ranks hold bit-identical totals)
template <int NMAX>
__device__ __forceinline__ void fc_mail_collect(const fc_p2p_dev *P, unsigned long long seq, int count, double (&t)[NMAX]) {
  const fc_mail *box = P->mail + (seq % FC_MAIL_SLOTS) * FC_MAX_RANKS;
  const unsigned long long tag = seq & 0xffffffffull;
#pragma unroll
  for (int i = 0; i < NMAX; ++i) t[i] = 0.0;
  fc_spin_guard g;
  for (int r = 0; r < P->nranks; ++r) {
#pragma unroll
    for (int i = 0; i < NMAX; ++i) {
      if (i < count) {
        unsigned long long lo, hi;
        while (((lo = fc_ld_relaxed_sys(box[r].w + 2 * i)) >> 32) != tag) g.tick();
        while (((hi = fc_ld_relaxed_sys(box[r].w + 2 * i + 1)) >> 32) != tag) g.tick();
        t[i] = t[i] + __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
      }
    }
  }
}

// Kernel prologue.  Returns false when the solve has converged (the kernel must return at once).
// Block 0 folds the pending reduction into the scalars, the other blocks wait for it.
__device__ __forceinline__ bool fc_kernel_begin(fc_scalars *sc, const fc_sync &sy) {
  if (sy.wait_seq) {
    if (threadIdx.x == 0) {
      if (blockIdx.x == 0) {
        if (!((volatile fc_scalars *)sc)->done) {
          double t[FC_MAX_RED];
          fc_mail_collect<FC_MAX_RED>(sy.p2p, sy.wait_seq, sy.wait_count, t);
          for (int i = 0; i < sy.wait_count; ++i) sc->red[i] = t[i];
          fc_scalar_step(sc, sy.wait_step, sy.hist);
        }
        fc_st_release_gpu(&sc->applied, sy.wait_seq);
      } else {
        fc_spin_guard g;
        while (fc_ld_acquire_gpu(&sc->applied) < sy.wait_seq) g.tick();
      }
    }
    __syncthreads();
  }
  return !((volatile fc_scalars *)sc)->done;
}

// Hand the finished reduction on: peer mailboxes (P2P), the scalar step itself (single rank), or
// just red[] for the NCCL all-reduce that follows on the stream.
template <int NR>
__device__ __forceinline__ void fc_reduction_done(fc_scalars *sc, const fc_sync &sy, const double (&v)[NR], int step) {
  if (sy.p2p) {
    fc_mail_post(sy.p2p, sy.post_seq, v, NR);
  } else {
#pragma unroll
    for (int r = 0; r < NR; ++r) sc->red[r] = v[r];
    if (sy.local) fc_scalar_step(sc, step, sy.hist);
  }
}
