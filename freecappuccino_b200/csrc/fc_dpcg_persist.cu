// dpcg(fi,ifi) (src/dpcg.f90:3-154, src-parallel/dpcg.f90) as ONE persistent cooperative kernel.
//
// The multi-kernel path (fc_krylov.cu) pays a kernel boundary per vector operation: three launches
// per iteration plus, on several GPUs, a pack kernel and waits in the prologues.  On 8 GPUs an
// iteration of the 216^3 case is ~50 us of memory traffic per GPU, so those boundaries -- not the
// traffic -- set the time.  Here the whole solve is one launch of 148 x k co-resident CTAs:
//
//   * every CTA owns a fixed, contiguous range of rows for ALL phases (residual, p-update, SpMV, x/r
//     update): a row is only ever touched by one SM, consecutive chunks share their x neighbours in
//     L1, and when the partition is small enough the vectors stay in L2 while the matrix streams past
//     them (evict-first).  Ranges are equal shares of a cost that counts a row with processor faces
//     twice (fc_row_range), so the CTAs on a partition boundary finish with the others;
//   * the SpMV is the TMA pipeline of fc_spmv_pipe.cuh; the first chunks of the next product are
//     requested before the reductions, so the copy engine works through the barriers;
//   * three grid barriers per iteration (dpcg.f90:85-142): after the p-update (the product gathers
//     other CTAs' rows), and one per inner product.  The CTA that arrives last adds the per-CTA
//     partial sums in a fixed order (deterministic, no float atomics), runs the scalar recurrence
//     (alpha, beta, convergence test) and releases the others;
//   * several GPUs (P2P mode, fc_p2p.cu): the p-update stores the boundary values of p straight
//     into the neighbours' halo slots over NVLink and the barrier's last CTA raises their arrival
//     flags; only rows with processor faces wait for a flag.  An inner product's last CTA posts the
//     rank's partial sums to every rank's mailbox and adds all mailboxes in rank order (global_sum
//     of src-parallel, bit-identical on all ranks) before it releases its own grid.
// Every wait is bounded (fc_spin_guard): a lost peer surfaces as a CUDA error, not as a hung GPU.
#include <cooperative_groups.h>

#include "fc_spmv_pipe.cuh"

namespace {

struct persist_args {
  fc_spmv_mat M;
  const double *su;
  double *fi, *pk, *zk, *res, *adiag;
  const int *diag;
  double padd, tol;
  fc_strip st;
  const int *bufind;
  fc_scalars *sc;
  fc_persist_state *ps;
  double *partials;
  fc_p2p_dev *p2p;
  unsigned long long red_seq0, halo_seq0;
  double *hist;
  int l2keep;   // the Krylov vectors fit the L2: mark them evict_last (the matrix stream is evict_first)
  int xprefetch;     // FC_TUNE_X_PREFETCH
  int eager;         // small partitions: fi += alf*pk runs behind the beta reduction, q = res/a_ii is handed to the p-update
  int mat_keep256;   // share (of 256) of the matrix chunks that stay in the L2 from one product to the next
  // "fused p" scheme (FUSED kernels): q = res / (a_ii + padd) (in the arena: the neighbours store its halo), the product's
  // result y, and the second direction buffer (pk / pk2 alternate as pold / pnew)
  double *q, *y, *pk2;
};

enum { PH_PUPDATE = 0, PH_SPMV = 1, PH_UPDATE = 2, PH_SETUP = 3 };

struct sync_ctx {
  unsigned long long my_gen;
  bool *s_last;
  double *s_red;
};

// thread 0 of a CTA that is not the last: wait for the release of barrier `my_gen`
__device__ __forceinline__ void wait_release(fc_persist_state *ps, unsigned long long my_gen) {
  fc_spin_guard g;
  while (fc_ld_acquire_gpu(&ps->gen) <= my_gen) g.tick();
}

// last CTA, thread 0: book-keeping of the phase clocks, then open the barrier
__device__ __forceinline__ void release_barrier(fc_persist_state *ps, unsigned long long my_gen, int phase,
                                                unsigned long long t_arrive) {
  ps->t_phase[phase] += t_arrive - ps->t_mark;
  const unsigned long long now = fc_globaltimer();
  ps->t_mark = now;
  ps->t_total = now - ps->t_start;
  ps->count = 0u;
  __threadfence();
  fc_st_release_gpu(&ps->gen, my_gen + 1ull);
}

// plain grid barrier; `hseq` != 0: the last CTA raises the neighbours' halo-arrival flags first
__device__ __forceinline__ void grid_barrier(const persist_args &A, sync_ctx &S, int phase, unsigned long long hseq,
                                             bool sent_remote) {
  __syncthreads();
  if (threadIdx.x == 0) {
    // one fence per CTA covers the block's stores (ordered before it by the block barrier); CTAs that stored
    // halo values into a neighbour's memory fence at system scope
    if (sent_remote) __threadfence_system();
    else __threadfence();
    const unsigned t = atomicAdd(&A.ps->count, 1u);
    if (t == gridDim.x - 1) {
      const unsigned long long t_arrive = fc_globaltimer();
      if (hseq && A.p2p) {   // every CTA's halo stores are ordered before its arrival: one fence, then the flags
        __threadfence_system();
        for (int c = 0; c < A.p2p->nconn; ++c) fc_st_relaxed_sys(A.p2p->peer_hflag[c], hseq);
      } else {
        __threadfence();
      }
      release_barrier(A.ps, S.my_gen, phase, t_arrive);
    } else {
      wait_release(A.ps, S.my_gen);
    }
    __threadfence();
  }
  S.my_gen++;
  __syncthreads();
}

// grid barrier that carries a reduction: v[] = per-thread partial sums on entry.  The last CTA
// finishes the sum, exchanges it with the other ranks and runs the scalar step `step`.
// `hseq` != 0 (fused-p scheme): the CTAs stored halo values into the neighbours' memory before this barrier; the last
// CTA raises the neighbours' halo-arrival flags before it releases its own grid (`sent_remote`: this CTA did store).
struct no_shadow { __device__ __forceinline__ void operator()() const {} };

// `shadow`: work of the whole CTA that does not depend on the reduction's result.  A CTA runs it between its arrival
// and its wait for the release, i.e. behind the reduction and the exchange with the other ranks; the CTA that arrives
// last -- and does the reduction -- runs it after it has released the others.
template <int NR, class Shadow = no_shadow>
__device__ __forceinline__ void grid_reduce(const persist_args &A, sync_ctx &S, double (&v)[NR], int step, int phase,
                                            unsigned long long seq, unsigned long long hseq = 0ull,
                                            bool sent_remote = false, Shadow shadow = Shadow()) {
  const int G = gridDim.x;
  __shared__ unsigned long long s_tarr;
  fc_block_sum<NR>(v, S.s_red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) A.partials[r * G + blockIdx.x] = v[r];
    if (sent_remote) __threadfence_system();
    else __threadfence();
    const unsigned t = atomicAdd(&A.ps->count, 1u);
    *S.s_last = (t == (unsigned)G - 1u);
    if (*S.s_last) s_tarr = fc_globaltimer();
  }
  __syncthreads();
  const bool last = *S.s_last;
  if (!last) shadow();
  if (last) {
    __threadfence();
    if (hseq && A.p2p && threadIdx.x == 0) {   // every CTA's halo stores are ordered before its arrival
      __threadfence_system();
      for (int c = 0; c < A.p2p->nconn; ++c) fc_st_relaxed_sys(A.p2p->peer_hflag[c], hseq);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double w = 0.0;
      for (int b = threadIdx.x; b < G; b += blockDim.x) w += __ldcg(A.partials + r * G + b);
      v[r] = w;
    }
    fc_block_sum<NR>(v, S.s_red);
    if (A.p2p) {
      // all-reduce over the ranks through the mailboxes, one thread per 8-byte word: thread t posts word
      // t % (2 NR) of my sums to rank t / (2 NR) and polls the same word of that rank's contribution to me,
      // so the exchange costs one NVLink flight, not one per word
      __shared__ double s_v[NR];
      __shared__ unsigned long long s_mail[FC_MAX_RANKS * 2 * NR];
      __shared__ unsigned long long s_tpost;
      const fc_p2p_dev *P = A.p2p;
      if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) s_v[r] = v[r];
        s_tpost = fc_globaltimer();
      }
      __syncthreads();
      const int nw = P->nranks * 2 * NR;
      for (int t = threadIdx.x; t < nw; t += blockDim.x) {
        const int q = t / (2 * NR), w = t % (2 * NR);
        const int slot = (int)(seq % FC_MAIL_SLOTS);
        const unsigned long long tag = seq & 0xffffffffull;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(s_v[w >> 1]);
        const unsigned long long word = ((w & 1) ? (bits >> 32) : (bits & 0xffffffffull)) | (tag << 32);
        fc_st_relaxed_sys(P->peer_mail[q][slot * FC_MAX_RANKS + P->rank].w + w, word);
        const unsigned long long *src = P->mail[slot * FC_MAX_RANKS + q].w + w;
        unsigned long long x;
        fc_spin_guard g;
        while (((x = fc_ld_relaxed_sys(src)) >> 32) != tag) g.tick();
        s_mail[t] = x;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NR; ++i) v[i] = 0.0;
        for (int r = 0; r < P->nranks; ++r) {   // rank order: identical sums on every rank (global_sum)
#pragma unroll
          for (int i = 0; i < NR; ++i) {
            const unsigned long long lo = s_mail[r * 2 * NR + 2 * i], hi = s_mail[r * 2 * NR + 2 * i + 1];
            v[i] = v[i] + __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
          }
        }
        A.ps->t_mail += fc_globaltimer() - s_tpost;
      }
    }
    if (threadIdx.x == 0) {
      fc_scalars *sc = A.sc;
#pragma unroll
      for (int r = 0; r < NR; ++r) sc->red[r] = v[r];
      fc_scalar_step(sc, step, A.hist);
      if (step == STEP_RES0_SK && A.tol >= 0.0 && sc->res0 < A.tol) sc->done = 1;   // dpcg.f90:66-70
      if (step == STEP_RES0_SK && sc->nsw <= 0) sc->done = 1;   // `do l=1,ns` with ns = 0: no iteration, fi untouched
      release_barrier(A.ps, S.my_gen, phase, s_tarr);
    }
    shadow();
  } else if (threadIdx.x == 0) {
    wait_release(A.ps, S.my_gen);
  }
  if (threadIdx.x == 0) __threadfence();
  S.my_gen++;
  __syncthreads();
}

template <int T, int CAP, int S, bool STRIP, bool FUSED, bool CODED>
__global__ void __launch_bounds__(T, 768 / T)   // <= 85 registers: three 256-thread CTAs per SM
k_dpcg_persist(persist_args A) {
  extern __shared__ __align__(128) unsigned char fc_smem_raw[];
  __shared__ double s_red[64];
  __shared__ bool s_last;
  // connection table of the halo stores (P2P mode)
  __shared__ int s_conn_off[FC_MAX_CONN + 1];
  __shared__ double *s_peer_pk[FC_MAX_CONN];
  const int tid = threadIdx.x;
  const int n = A.M.n;
  fc_scalars *sc = A.sc;
  sync_ctx SY{fc_ld_acquire_gpu(&A.ps->gen), &s_last, s_red};

  fc_spmv_pipe<T, CAP, S, CODED> pipe;
  auto *sm = reinterpret_cast<fc_spmv_smem<T, CAP, S, CODED> *>(fc_smem_raw);
  int rbeg, rend;
  fc_row_range(n, STRIP ? A.st.off : nullptr, &rbeg, &rend);
  pipe.init(sm, A.M.ioffset, rbeg, rend, STRIP ? A.st.off : nullptr, A.M.dict);
  // L2 policy of the vector traffic (pk, zk, res, adiag): evict_last when they fit the L2 next to the matrix stream
  const unsigned long long vpol = A.l2keep ? fc_policy_evict_last() : fc_policy_evict_normal();
  pipe.pol_y = vpol;
  pipe.keep256 = A.mat_keep256;
  pipe.xprefetch = A.xprefetch;
  pipe.pol_keep = fc_policy_evict_last();
  pipe.prefetch(A.M);
  const bool p2p = STRIP && A.p2p != nullptr;
  if (p2p) {
    if (tid <= A.p2p->nconn) s_conn_off[tid] = A.p2p->conn_off[tid];
    if (tid < A.p2p->nconn) s_peer_pk[tid] = FUSED ? A.p2p->peer_zk[tid] : A.p2p->peer_pk[tid];   // FUSED: q lives in zk
    __syncthreads();
  }
  const int nch = pipe.nch;
  fc_strip st = A.st;
  st.hseq = 0ull;   // the residual uses fi's halo as it is (src-parallel/dpcg.f90:58-66)
  unsigned long long red_seq = A.red_seq0, halo_seq = A.halo_seq0;

  // my cells on a processor boundary: their values of `src` go straight into the neighbours' halo slots
  auto send_halo = [&](const double *src) {
    __syncthreads();
    for (int j = 0; j < nch; j += 2) {   // two chunks at a time: the load chains of the two rows overlap
      int row[2], q0[2], q1[2], f[2];
      double v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        row[u] = rbeg + (j + u) * T + tid;
        const bool act = j + u < nch && sm->cs[j + u] && row[u] < rend;
        q0[u] = act ? A.st.off[row[u]] : 0;
        q1[u] = act ? A.st.off[row[u] + 1] : 0;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (q1[u] > q0[u]) {
          f[u] = A.st.idx[q0[u]];
          v[u] = src[row[u]];
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        for (int q = q0[u]; q < q1[u]; ++q) {
          const int ff = q == q0[u] ? f[u] : A.st.idx[q];
          int c = 0;
          while (ff >= s_conn_off[c + 1]) ++c;
          s_peer_pk[c][ff - s_conn_off[c]] = v[u];
        }
      }
    }
  };

  if (FUSED) {
    // ================= "fused p" scheme: two phases and two grid barriers per iteration =================
    // The direction vector is never swept on its own.  The x/r update leaves q = res/(a_ii[+small]) behind (it needs
    // that quotient for the next sk anyway) and sends the boundary values of q to the neighbours; the product then
    // gathers p(j) = q(j) + bet*pold(j) -- dpcg.f90:95-100 evaluated where it is used, same operands and rounding,
    // so every iterate is bit-identical to the three-phase kernel -- writes the row's own p to the other direction
    // buffer and carries the deferred x update.  A rank forms the p of its halo cells itself from the q it received
    // and the pold it formed an iteration earlier.  One phase (48 n bytes, a division per row) and one grid barrier
    // less per iteration: what limits the strong scaling at ~1.3 M cells per GPU.
    {
      fc_spmv_vec V{A.fi, A.res, A.su, nullptr, A.diag, A.adiag, A.padd};
      V.qout = A.q;
      double v[2] = {0.0, 0.0};
      pipe.template sweep<FC_MODE_RESID_SK, STRIP>(A.M, V, st, v[0], v[1]);
      pipe.prefetch(A.M);
      if (p2p && pipe.cta_strip) send_halo(A.q);
      ++halo_seq;
      grid_reduce<2>(A, SY, v, STEP_RES0_SK, PH_SETUP, ++red_seq, p2p ? halo_seq : 0ull, p2p && pipe.cta_strip);
    }
    double *pold = A.pk, *pnew = A.pk2;
    bool first = true;
    while (!__ldcg(&sc->done)) {
      // ---- y = A p, p = q + bet*pold ; pkapk = p.y ; fi += alf_prev*pold   (dpcg.f90:95-124) ----
      {
        if (p2p) {
          st.hseq = halo_seq;
          pipe.halo_pending = pipe.cta_strip;   // only CTAs that own rows with processor faces wait for the q halo
        }
        fc_spmv_vec V{pold, A.y, nullptr, nullptr, nullptr, nullptr, 0.0};
        V.q = A.q;
        V.pnew = pnew;
        V.fi = A.fi;
        V.bet = __ldcg(&sc->sk) / __ldcg(&sc->s0);
        V.alfp = first ? 0.0 : __ldcg(&sc->s0) / __ldcg(&sc->pkapk);
        first = false;
        double v[1] = {0.0};
        double unused = 0.0;
        pipe.template sweep<FC_MODE_DOT_FUSED, STRIP>(A.M, V, st, v[0], unused);
        pipe.prefetch(A.M);
        grid_reduce<1>(A, SY, v, STEP_PKAPK, PH_SPMV, ++red_seq);
      }
      // ---- res -= alf*y ; q = res/(a_ii[+small]) ; resl = sum|res| ; sk = sum res*q   (dpcg.f90:125-142, :85-90) ----
      {
        const double alf = __ldcg(&sc->sk) / __ldcg(&sc->pkapk);
        double a0 = 0.0, a1 = 0.0;
        int i = rbeg + tid;
        for (; i + 3 * T < rend; i += 4 * T) {
          const double r0 = fc_ld_pol(A.res + i, vpol), r1 = fc_ld_pol(A.res + i + T, vpol),
                       r2 = fc_ld_pol(A.res + i + 2 * T, vpol), r3 = fc_ld_pol(A.res + i + 3 * T, vpol);
          const double z0 = fc_ld_pol(A.y + i, vpol), z1 = fc_ld_pol(A.y + i + T, vpol),
                       z2 = fc_ld_pol(A.y + i + 2 * T, vpol), z3 = fc_ld_pol(A.y + i + 3 * T, vpol);
          const double d0 = fc_ld_pol(A.adiag + i, vpol), d1 = fc_ld_pol(A.adiag + i + T, vpol),
                       d2 = fc_ld_pol(A.adiag + i + 2 * T, vpol), d3 = fc_ld_pol(A.adiag + i + 3 * T, vpol);
          const double n0 = r0 - alf * z0, n1 = r1 - alf * z1, n2 = r2 - alf * z2, n3 = r3 - alf * z3;
          const double q0 = n0 / (d0 + A.padd), q1 = n1 / (d1 + A.padd), q2 = n2 / (d2 + A.padd), q3 = n3 / (d3 + A.padd);
          fc_st_pol(A.res + i, n0, vpol);
          fc_st_pol(A.res + i + T, n1, vpol);
          fc_st_pol(A.res + i + 2 * T, n2, vpol);
          fc_st_pol(A.res + i + 3 * T, n3, vpol);
          fc_st_pol(A.q + i, q0, vpol);
          fc_st_pol(A.q + i + T, q1, vpol);
          fc_st_pol(A.q + i + 2 * T, q2, vpol);
          fc_st_pol(A.q + i + 3 * T, q3, vpol);
          a0 += fabs(n0); a1 += n0 * q0;
          a0 += fabs(n1); a1 += n1 * q1;
          a0 += fabs(n2); a1 += n2 * q2;
          a0 += fabs(n3); a1 += n3 * q3;
        }
        for (; i < rend; i += T) {
          const double n0 = fc_ld_pol(A.res + i, vpol) - alf * fc_ld_pol(A.y + i, vpol);
          const double q0 = n0 / (fc_ld_pol(A.adiag + i, vpol) + A.padd);
          fc_st_pol(A.res + i, n0, vpol);
          fc_st_pol(A.q + i, q0, vpol);
          a0 += fabs(n0); a1 += n0 * q0;
        }
        if (p2p && pipe.cta_strip) send_halo(A.q);
        ++halo_seq;
        double v[2] = {a0, a1};
        grid_reduce<2>(A, SY, v, STEP_CG_UPDATE_SK, PH_UPDATE, ++red_seq, p2p ? halo_seq : 0ull, p2p && pipe.cta_strip);
      }
      double *t = pold; pold = pnew; pnew = t;
    }
    if (!first) {   // the x update of the last iteration (dpcg.f90:121-124); pold = the direction of that iteration
      const double alf = __ldcg(&sc->s0) / __ldcg(&sc->pkapk);
      for (int i = rbeg + tid; i < rend; i += T) A.fi[i] = A.fi[i] + alf * pold[i];
    }
    pipe.drain();
    return;
  }

  // ---- res = su - A fi ; res0 = sum|res| ; sk = sum res*res/(a_ii[+small])   (dpcg.f90:51-90) ----
  {
    fc_spmv_vec V{A.fi, A.res, A.su, nullptr, A.diag, A.adiag, A.padd};
    if (A.eager) V.qout = A.zk;
    double v[2] = {0.0, 0.0};
    pipe.template sweep<FC_MODE_RESID_SK, STRIP>(A.M, V, st, v[0], v[1]);
    pipe.prefetch(A.M);
    grid_reduce<2>(A, SY, v, STEP_RES0_SK, PH_SETUP, ++red_seq);
  }

  // ================= small partitions (A.eager): the vectors live in the L2 and an iteration is a few tens of
  // microseconds, a third of it barriers and reductions.  Two changes, both with the operands and the rounding of
  // dpcg.f90:95-142 (bit-identical iterates):
  //   * the x/r update leaves q = res/(a_ii[+small]) in zk (it needs that quotient for the next sk anyway), so the
  //     p-update reads q and pk only -- no division, 24 instead of 48 bytes per row;
  //   * fi += alf*pk does not ride in the p-update but runs BEHIND the beta reduction: every CTA does its share
  //     between its arrival at the barrier and its wait for the release, while the last CTA adds the partial sums
  //     and exchanges them with the other ranks.
  if (A.eager) {
    while (!__ldcg(&sc->done)) {
      {
        const double bet = __ldcg(&sc->sk) / __ldcg(&sc->s0);
        int i = rbeg + tid;
        for (; i + 3 * T < rend; i += 4 * T) {
          const double q0 = fc_ld_pol(A.zk + i, vpol), q1 = fc_ld_pol(A.zk + i + T, vpol),
                       q2 = fc_ld_pol(A.zk + i + 2 * T, vpol), q3 = fc_ld_pol(A.zk + i + 3 * T, vpol);
          const double p0 = fc_ld_pol(A.pk + i, vpol), p1 = fc_ld_pol(A.pk + i + T, vpol),
                       p2 = fc_ld_pol(A.pk + i + 2 * T, vpol), p3 = fc_ld_pol(A.pk + i + 3 * T, vpol);
          fc_st_pol(A.pk + i, q0 + bet * p0, vpol);
          fc_st_pol(A.pk + i + T, q1 + bet * p1, vpol);
          fc_st_pol(A.pk + i + 2 * T, q2 + bet * p2, vpol);
          fc_st_pol(A.pk + i + 3 * T, q3 + bet * p3, vpol);
        }
        for (; i < rend; i += T) fc_st_pol(A.pk + i, fc_ld_pol(A.zk + i, vpol) + bet * fc_ld_pol(A.pk + i, vpol), vpol);
        if (p2p && pipe.cta_strip) send_halo(A.pk);
      }
      ++halo_seq;
      grid_barrier(A, SY, PH_PUPDATE, p2p ? halo_seq : 0ull, p2p && pipe.cta_strip);
      {
        if (p2p) {
          st.hseq = halo_seq;
          pipe.halo_pending = pipe.cta_strip;
        }
        fc_spmv_vec V{A.pk, A.zk, nullptr, A.pk, nullptr, nullptr, 0.0};
        double v[1] = {0.0};
        double unused = 0.0;
        pipe.template sweep<FC_MODE_DOT, STRIP>(A.M, V, st, v[0], unused);
        pipe.prefetch(A.M);
        grid_reduce<1>(A, SY, v, STEP_PKAPK, PH_SPMV, ++red_seq);
      }
      {
        const double alf = __ldcg(&sc->sk) / __ldcg(&sc->pkapk);
        double a0 = 0.0, a1 = 0.0;
        int i = rbeg + tid;
        for (; i + 3 * T < rend; i += 4 * T) {
          const double r0 = fc_ld_pol(A.res + i, vpol), r1 = fc_ld_pol(A.res + i + T, vpol),
                       r2 = fc_ld_pol(A.res + i + 2 * T, vpol), r3 = fc_ld_pol(A.res + i + 3 * T, vpol);
          const double z0 = fc_ld_pol(A.zk + i, vpol), z1 = fc_ld_pol(A.zk + i + T, vpol),
                       z2 = fc_ld_pol(A.zk + i + 2 * T, vpol), z3 = fc_ld_pol(A.zk + i + 3 * T, vpol);
          const double d0 = fc_ld_pol(A.adiag + i, vpol), d1 = fc_ld_pol(A.adiag + i + T, vpol),
                       d2 = fc_ld_pol(A.adiag + i + 2 * T, vpol), d3 = fc_ld_pol(A.adiag + i + 3 * T, vpol);
          const double n0 = r0 - alf * z0, n1 = r1 - alf * z1, n2 = r2 - alf * z2, n3 = r3 - alf * z3;
          const double q0 = n0 / (d0 + A.padd), q1 = n1 / (d1 + A.padd), q2 = n2 / (d2 + A.padd), q3 = n3 / (d3 + A.padd);
          fc_st_pol(A.res + i, n0, vpol);
          fc_st_pol(A.res + i + T, n1, vpol);
          fc_st_pol(A.res + i + 2 * T, n2, vpol);
          fc_st_pol(A.res + i + 3 * T, n3, vpol);
          fc_st_pol(A.zk + i, q0, vpol);
          fc_st_pol(A.zk + i + T, q1, vpol);
          fc_st_pol(A.zk + i + 2 * T, q2, vpol);
          fc_st_pol(A.zk + i + 3 * T, q3, vpol);
          a0 += fabs(n0); a1 += n0 * q0;
          a0 += fabs(n1); a1 += n1 * q1;
          a0 += fabs(n2); a1 += n2 * q2;
          a0 += fabs(n3); a1 += n3 * q3;
        }
        for (; i < rend; i += T) {
          const double n0 = fc_ld_pol(A.res + i, vpol) - alf * fc_ld_pol(A.zk + i, vpol);
          const double q0 = n0 / (fc_ld_pol(A.adiag + i, vpol) + A.padd);
          fc_st_pol(A.res + i, n0, vpol);
          fc_st_pol(A.zk + i, q0, vpol);
          a0 += fabs(n0); a1 += n0 * q0;
        }
        double v[2] = {a0, a1};
        auto x_update = [&]() {   // fi += alf*pk (dpcg.f90:121-124) behind the reduction
          int k = rbeg + tid;
          for (; k + 3 * T < rend; k += 4 * T) {
            const double p0 = fc_ld_pol(A.pk + k, vpol), p1 = fc_ld_pol(A.pk + k + T, vpol),
                         p2 = fc_ld_pol(A.pk + k + 2 * T, vpol), p3 = fc_ld_pol(A.pk + k + 3 * T, vpol);
            const double f0 = A.fi[k], f1 = A.fi[k + T], f2 = A.fi[k + 2 * T], f3 = A.fi[k + 3 * T];
            A.fi[k] = f0 + alf * p0;
            A.fi[k + T] = f1 + alf * p1;
            A.fi[k + 2 * T] = f2 + alf * p2;
            A.fi[k + 3 * T] = f3 + alf * p3;
          }
          for (; k < rend; k += T) A.fi[k] = A.fi[k] + alf * fc_ld_pol(A.pk + k, vpol);
        };
        grid_reduce<2>(A, SY, v, STEP_CG_UPDATE_SK, PH_UPDATE, ++red_seq, 0ull, false, x_update);
      }
    }
    pipe.drain();
    return;
  }

  bool first = true;
  while (!__ldcg(&sc->done)) {
    // ---- pk = res/(a_ii[+small]) + bet*pk   (dpcg.f90:95-100); exchange(pk) (src-parallel/dpcg.f90:114): the
    //      owner of a cell on a processor boundary stores its new value straight into the neighbour's halo ----
    {
      // The x update of the PREVIOUS iteration, fi += alf*pk (dpcg.f90:121-124), rides along: this phase reads
      // pk anyway, so deferring it saves one pass over pk per iteration.  alf = sk_prev / pkapk, and the last
      // reduction left sk_prev in s0 (dpcg.f90:139), so the value -- and fi -- are bit-identical.
      const double bet = __ldcg(&sc->sk) / __ldcg(&sc->s0);
      const double alfp = first ? 0.0 : __ldcg(&sc->s0) / __ldcg(&sc->pkapk);
      first = false;
      int i = rbeg + tid;
      for (; i + 3 * T < rend; i += 4 * T) {
        const double r0 = fc_ld_pol(A.res + i, vpol), r1 = fc_ld_pol(A.res + i + T, vpol),
                     r2 = fc_ld_pol(A.res + i + 2 * T, vpol), r3 = fc_ld_pol(A.res + i + 3 * T, vpol);
        const double d0 = fc_ld_pol(A.adiag + i, vpol), d1 = fc_ld_pol(A.adiag + i + T, vpol),
                     d2 = fc_ld_pol(A.adiag + i + 2 * T, vpol), d3 = fc_ld_pol(A.adiag + i + 3 * T, vpol);
        const double p0 = fc_ld_pol(A.pk + i, vpol), p1 = fc_ld_pol(A.pk + i + T, vpol),
                     p2 = fc_ld_pol(A.pk + i + 2 * T, vpol), p3 = fc_ld_pol(A.pk + i + 3 * T, vpol);
        const double f0 = A.fi[i], f1 = A.fi[i + T], f2 = A.fi[i + 2 * T], f3 = A.fi[i + 3 * T];
        A.fi[i] = f0 + alfp * p0;
        A.fi[i + T] = f1 + alfp * p1;
        A.fi[i + 2 * T] = f2 + alfp * p2;
        A.fi[i + 3 * T] = f3 + alfp * p3;
        fc_st_pol(A.pk + i, r0 / (d0 + A.padd) + bet * p0, vpol);
        fc_st_pol(A.pk + i + T, r1 / (d1 + A.padd) + bet * p1, vpol);
        fc_st_pol(A.pk + i + 2 * T, r2 / (d2 + A.padd) + bet * p2, vpol);
        fc_st_pol(A.pk + i + 3 * T, r3 / (d3 + A.padd) + bet * p3, vpol);
      }
      if (i < rend) {   // up to three rows left: one predicated batch instead of three dependent round trips
        const bool b1 = i + T < rend, b2 = i + 2 * T < rend;
        const double r0 = fc_ld_pol(A.res + i, vpol), d0 = fc_ld_pol(A.adiag + i, vpol), p0 = fc_ld_pol(A.pk + i, vpol),
                     f0 = A.fi[i];
        const double r1 = b1 ? fc_ld_pol(A.res + i + T, vpol) : 0.0, d1 = b1 ? fc_ld_pol(A.adiag + i + T, vpol) : 1.0,
                     p1 = b1 ? fc_ld_pol(A.pk + i + T, vpol) : 0.0;
        const double f1 = b1 ? A.fi[i + T] : 0.0;
        const double r2 = b2 ? fc_ld_pol(A.res + i + 2 * T, vpol) : 0.0, d2 = b2 ? fc_ld_pol(A.adiag + i + 2 * T, vpol) : 1.0;
        const double p2 = b2 ? fc_ld_pol(A.pk + i + 2 * T, vpol) : 0.0, f2 = b2 ? A.fi[i + 2 * T] : 0.0;
        A.fi[i] = f0 + alfp * p0;
        fc_st_pol(A.pk + i, r0 / (d0 + A.padd) + bet * p0, vpol);
        if (b1) {
          A.fi[i + T] = f1 + alfp * p1;
          fc_st_pol(A.pk + i + T, r1 / (d1 + A.padd) + bet * p1, vpol);
        }
        if (b2) {
          A.fi[i + 2 * T] = f2 + alfp * p2;
          fc_st_pol(A.pk + i + 2 * T, r2 / (d2 + A.padd) + bet * p2, vpol);
        }
      }
      if (p2p && pipe.cta_strip) send_halo(A.pk);   // exchange(pk): straight into the neighbours' halo slots
    }
    ++halo_seq;
    grid_barrier(A, SY, PH_PUPDATE, p2p ? halo_seq : 0ull, p2p && pipe.cta_strip);

    // ---- zk = A pk ; pkapk = pk.zk   (dpcg.f90:105-119) ----
    {
      if (p2p) {
        st.hseq = halo_seq;
        pipe.halo_pending = pipe.cta_strip;   // only CTAs that own rows with processor faces wait
      }
      fc_spmv_vec V{A.pk, A.zk, nullptr, A.pk, nullptr, nullptr, 0.0};
      double v[1] = {0.0};
      double unused = 0.0;
      pipe.template sweep<FC_MODE_DOT, STRIP>(A.M, V, st, v[0], unused);
      pipe.prefetch(A.M);   // the next product's first chunks load behind the two reductions
      grid_reduce<1>(A, SY, v, STEP_PKAPK, PH_SPMV, ++red_seq);
    }

    // ---- res -= alf*zk ; resl = sum|res| ; next sk   (dpcg.f90:125-142; fi += alf*pk is deferred, see above) ----
    {
      const double alf = __ldcg(&sc->sk) / __ldcg(&sc->pkapk);
      double a0 = 0.0, a1 = 0.0;
      int i = rbeg + tid;
      for (; i + 3 * T < rend; i += 4 * T) {
        const double r0 = fc_ld_pol(A.res + i, vpol), r1 = fc_ld_pol(A.res + i + T, vpol),
                     r2 = fc_ld_pol(A.res + i + 2 * T, vpol), r3 = fc_ld_pol(A.res + i + 3 * T, vpol);
        const double z0 = fc_ld_pol(A.zk + i, vpol), z1 = fc_ld_pol(A.zk + i + T, vpol),
                     z2 = fc_ld_pol(A.zk + i + 2 * T, vpol), z3 = fc_ld_pol(A.zk + i + 3 * T, vpol);
        const double d0 = fc_ld_pol(A.adiag + i, vpol), d1 = fc_ld_pol(A.adiag + i + T, vpol),
                     d2 = fc_ld_pol(A.adiag + i + 2 * T, vpol), d3 = fc_ld_pol(A.adiag + i + 3 * T, vpol);
        const double n0 = r0 - alf * z0, n1 = r1 - alf * z1, n2 = r2 - alf * z2, n3 = r3 - alf * z3;
        fc_st_pol(A.res + i, n0, vpol);
        fc_st_pol(A.res + i + T, n1, vpol);
        fc_st_pol(A.res + i + 2 * T, n2, vpol);
        fc_st_pol(A.res + i + 3 * T, n3, vpol);
        a0 += fabs(n0); a1 += n0 * (n0 / (d0 + A.padd));
        a0 += fabs(n1); a1 += n1 * (n1 / (d1 + A.padd));
        a0 += fabs(n2); a1 += n2 * (n2 / (d2 + A.padd));
        a0 += fabs(n3); a1 += n3 * (n3 / (d3 + A.padd));
      }
      if (i < rend) {
        const bool b1 = i + T < rend, b2 = i + 2 * T < rend;
        const double r0 = fc_ld_pol(A.res + i, vpol), z0 = fc_ld_pol(A.zk + i, vpol), d0 = fc_ld_pol(A.adiag + i, vpol);
        const double r1 = b1 ? fc_ld_pol(A.res + i + T, vpol) : 0.0, z1 = b1 ? fc_ld_pol(A.zk + i + T, vpol) : 0.0,
                     d1 = b1 ? fc_ld_pol(A.adiag + i + T, vpol) : 1.0;
        const double r2 = b2 ? fc_ld_pol(A.res + i + 2 * T, vpol) : 0.0, z2 = b2 ? fc_ld_pol(A.zk + i + 2 * T, vpol) : 0.0;
        const double d2 = b2 ? fc_ld_pol(A.adiag + i + 2 * T, vpol) : 1.0;
        const double n0 = r0 - alf * z0;
        fc_st_pol(A.res + i, n0, vpol);
        a0 += fabs(n0); a1 += n0 * (n0 / (d0 + A.padd));
        if (b1) {
          const double n1 = r1 - alf * z1;
          fc_st_pol(A.res + i + T, n1, vpol);
          a0 += fabs(n1); a1 += n1 * (n1 / (d1 + A.padd));
        }
        if (b2) {
          const double n2 = r2 - alf * z2;
          fc_st_pol(A.res + i + 2 * T, n2, vpol);
          a0 += fabs(n2); a1 += n2 * (n2 / (d2 + A.padd));
        }
      }
      double v[2] = {a0, a1};
      grid_reduce<2>(A, SY, v, STEP_CG_UPDATE_SK, PH_UPDATE, ++red_seq);
    }
  }
  if (!first) {   // the x update of the last iteration (dpcg.f90:121-124)
    const double alf = __ldcg(&sc->s0) / __ldcg(&sc->pkapk);
    for (int i = rbeg + tid; i < rend; i += T) A.fi[i] = A.fi[i] + alf * A.pk[i];
  }
  pipe.drain();   // no CTA may exit with bulk copies in flight
}

template <int T, int CAP, int S, bool STRIP, bool FUSED, bool CODED>
int launch_persist(fc_context *ctx, persist_args &A, bool *ok) {
  auto kern = k_dpcg_persist<T, CAP, S, STRIP, FUSED, CODED>;
  const size_t smem = sizeof(fc_spmv_smem<T, CAP, S, CODED>);
  // per instantiation and per device (the shared-memory opt-in is a per-device function attribute)
  static int per_sm_dev[FC_MAX_DEVICES];
  static bool per_sm_set[FC_MAX_DEVICES];
  const int dev = ctx->device >= 0 && ctx->device < FC_MAX_DEVICES ? ctx->device : 0;
  if (!per_sm_set[dev] || dev != ctx->device) {
    int v = 0;
    FC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, T, smem));
    per_sm_dev[dev] = v < 1 ? 0 : v;
    per_sm_set[dev] = true;
  }
  int use = per_sm_dev[dev];
  if (ctx->tune_ctas_per_sm > 0 && ctx->tune_ctas_per_sm < use) use = ctx->tune_ctas_per_sm;
  if (use == 0) { *ok = false; return FC_OK; }
  const int sms = ctx->sms;
  const long long groups = ((long long)A.M.n + 31) / 32;
  long long grid = (long long)sms * use;
  if (grid > groups) grid = groups;
  if (grid < 1) grid = 1;
  // largest row range of a CTA (one without processor faces) must fit the chunk table
  const long long rows_per_cta = (((long long)A.M.n + ctx->npro + grid - 1) / grid / 32 + 2) * 32;
  if ((rows_per_cta + T - 1) / T > FC_KB_MAX || grid * FC_MAX_RED > (long long)FC_MAX_RED * 2048) {
    *ok = false;
    return FC_OK;
  }
  void *args[] = {(void *)&A};
  FC_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)grid), dim3(T), args, smem, ctx->stream));
  ctx->launches++;
  ctx->tm.persist_grid = (int)grid;
  *ok = true;
  return FC_OK;
}

__global__ void k_persist_begin(fc_persist_state *ps) {
  ps->count = 0u;
  for (int i = 0; i < 4; ++i) ps->t_phase[i] = 0ull;
  const unsigned long long now = fc_globaltimer();
  ps->t_start = now;
  ps->t_mark = now;
  ps->t_total = 0ull;
  ps->t_mail = 0ull;
}

}  // namespace

// `handled` = false: the caller must run the multi-kernel path (NCCL mode, a CSR pattern whose rows
// do not fit the staging buffers, no cooperative launch).
int fc_dpcg_persistent(fc_context *ctx, double *fi, const fc_solver_opts *o, fc_solver_report *rep, double *hist,
                       bool *handled) {
  *handled = false;
  if (!ctx->tune_persist) return FC_OK;
  if (ctx->nranks > 1 && !ctx->p2p) return FC_OK;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
  if (!coop) return FC_OK;
  const int n = ctx->n;
  cudaStream_t st = ctx->stream;
  persist_args A{};
  A.M = fc_spmv_mat{n, ctx->ioffset, ctx->ja, ctx->field[FC_A]};
  A.su = ctx->field[FC_SU];
  A.fi = fi; A.pk = ctx->pk; A.zk = ctx->zk; A.res = ctx->field[FC_RES]; A.adiag = ctx->adiag;
  A.diag = ctx->diag;
  A.padd = o->parallel ? o->small : 0.0;
  A.tol = o->tol;
  A.st = fc_strip{ctx->strip_off, ctx->strip_idx, ctx->field[FC_APR], ctx->strip_any32, n, nullptr, 0ull, 0};
  if (ctx->p2p) {
    A.st.hflag = (const unsigned long long *)((const char *)ctx->arena + ctx->arena_hflag_off);
    A.st.nconn = (int)ctx->nbr_rank.size();
  }
  A.bufind = ctx->bufind;
  A.sc = ctx->sc;
  A.ps = ctx->persist;
  A.partials = ctx->partials;
  A.p2p = ctx->p2p ? ctx->p2p_dev : nullptr;
  A.red_seq0 = ctx->red_seq;
  A.halo_seq0 = ctx->halo_seq;
  A.hist = hist;
  // four vectors of (n + npro) doubles against the 126 MB L2, which the matrix stream shares
  A.l2keep = ctx->tune_l2_keep == 1 || (ctx->tune_l2_keep == 2 && 32.0 * ((double)n + ctx->npro) <= 56e6);
  {
    int keep = ctx->tune_mat_keep;
    if (keep < 0) keep = 0;
    A.mat_keep256 = keep * 256 / 100;
    A.xprefetch = ctx->tune_x_prefetch;
    // measured (profiles/r02_dpcg_eager.jsonl): -4.6 % per iteration at 1.26 M rows, -4.7 % at 2.5 M (the vectors are
    // mostly L2-resident and the barriers a tenth of the iteration), +1.5 % at 10 M (there the deferred x update saves
    // a pass over pk that no barrier can hide)
    A.eager = ctx->tune_dpcg_eager == 1 || (ctx->tune_dpcg_eager == 2 && (double)n + ctx->npro <= 3.2e6);
  }

  // fused-p scheme (FC_TUNE_DPCG_FUSED: 0 never, 1 always, [2] on partitioned meshes, where an iteration is short and
  // the phase / barrier it saves is a fifth of it; on one GPU at 10 M cells the extra gathers cost more than it saves)
  const bool fused = ctx->tune_dpcg_fused == 1 || (ctx->tune_dpcg_fused == 2 && ctx->nranks > 1);
  A.q = ctx->zk; A.y = ctx->uk; A.pk2 = ctx->reso;
  // one-byte column codes instead of `ja` (FC_TUNE_JA_CODED) when the pattern has them; the fused-p option keeps `ja`
  // (measured, profiles/r02_ja_coded.jsonl: -2 % per iteration at 10 M rows, -6 % at 2.5 M, +0.7 % at 1.26 M where
  // the product is bound by the latency of a chunk, not by its bytes)
  const bool coded = !fused && ctx->coded_ok && A.M.ja == ctx->ja &&
                     (ctx->tune_ja_coded == 1 || (ctx->tune_ja_coded == 2 && n >= 2000000 && ctx->nranks == 1));
  // (by default on one rank only: the coded kernel with processor strips compiles from the same template and is
  // selectable with FC_TUNE_JA_CODED = 1, but the round's GPU time ended before it had run on several GPUs)
  if (coded) { A.M.jc = ctx->jcode; A.M.dict = ctx->jdict; }
  FC_CUDA(cudaMemsetAsync(ctx->pk, 0, sizeof(double) * ((size_t)n + ctx->npro), st));
  if (fused) FC_CUDA(cudaMemsetAsync(ctx->reso, 0, sizeof(double) * ((size_t)n + ctx->npro), st));
  k_persist_begin<<<1, 1, 0, st>>>(ctx->persist);
  FC_LAUNCH_CHECK();
  const bool strip = ctx->npro > 0;
  bool ok = false;
#define FC_PERSIST(T, CAP, S)                                                                            \
  do {                                                                                                   \
    if (strip && fused) FC_CHECK((launch_persist<T, CAP, S, true, true, false>(ctx, A, &ok)));           \
    else if (fused)     FC_CHECK((launch_persist<T, CAP, S, false, true, false>(ctx, A, &ok)));          \
    else if (strip && coded) FC_CHECK((launch_persist<T, CAP, S, true, false, true>(ctx, A, &ok)));      \
    else if (strip)     FC_CHECK((launch_persist<T, CAP, S, true, false, false>(ctx, A, &ok)));          \
    else if (coded)     FC_CHECK((launch_persist<T, CAP, S, false, false, true>(ctx, A, &ok)));          \
    else                FC_CHECK((launch_persist<T, CAP, S, false, false, false>(ctx, A, &ok)));         \
  } while (0)
#define FC_PERSIST_CODED(T, CAP, S)                                                                      \
  do {                                                                                                   \
    if (strip) FC_CHECK((launch_persist<T, CAP, S, true, false, true>(ctx, A, &ok)));                    \
    else       FC_CHECK((launch_persist<T, CAP, S, false, false, true>(ctx, A, &ok)));                   \
  } while (0)
  // One-byte codes shrink a 256-row chunk of a hexahedral mesh from 21.5 to 16.1 KB.  The pipeline is bound by the
  // bytes it keeps in flight (one chunk per CTA with two stages: ~1.5 us of loaded HBM latency), so the coded kernel
  // takes a third stage -- two chunks in flight per CTA, still three CTAs per SM -- with stages of 1824 non-zeros
  // (256 rows x 7 + the alignment slack of the copies).
  const int pipe_sel = (ctx->tune_pipe == 4 && !(coded && ctx->spmv_max_chunk <= 1792)) ? 1 : ctx->tune_pipe;
  if (ctx->spmv_max_chunk <= 2000) {
    switch (pipe_sel) {
      case 4: FC_PERSIST_CODED(256, 1824, 3); break;
      case 1: FC_PERSIST(256, 2304, 2); break;
      case 2: FC_PERSIST(256, 2048, 2); break;
      case 3: FC_PERSIST(128, 1024, 2); break;
      default: FC_PERSIST(256, 2304, 3); break;
    }
  } else {
    FC_PERSIST(256, 4096, 2);
  }
#undef FC_PERSIST
#undef FC_PERSIST_CODED
  if (!ok) return FC_OK;
  FC_CUDA(cudaMemcpyAsync(ctx->sc_host, ctx->sc, sizeof(fc_scalars), cudaMemcpyDeviceToHost, st));
  FC_CUDA(cudaMemcpyAsync(ctx->persist_host, ctx->persist, sizeof(fc_persist_state), cudaMemcpyDeviceToHost, st));
  FC_CUDA(cudaStreamSynchronize(st));
  rep->res0 = ctx->sc_host->res0;
  rep->resl = ctx->sc_host->resl;
  rep->iters = ctx->sc_host->iters;
  // the sequence numbers the kernel consumed (identical on every rank)
  ctx->red_seq += 1ull + 2ull * (unsigned long long)rep->iters;
  ctx->halo_seq += (unsigned long long)rep->iters + (fused ? 1ull : 0ull);   // fused: q's halo also after the residual
  const fc_persist_state &ps = *ctx->persist_host;
  ctx->tm.persist_ms = 1e-6 * (double)ps.t_total;
  ctx->tm.persist_pupdate_ms = 1e-6 * (double)ps.t_phase[PH_PUPDATE];
  ctx->tm.persist_spmv_ms = 1e-6 * (double)ps.t_phase[PH_SPMV];
  ctx->tm.persist_update_ms = 1e-6 * (double)ps.t_phase[PH_UPDATE];
  ctx->tm.persist_mail_ms = 1e-6 * (double)ps.t_mail;
  ctx->tm.persist_iters = rep->iters;
  ctx->tm.persist_index_bytes = coded ? 1 : 4;
  if (rep->iters > 0) {
    ctx->tm.spmv_ms = ctx->tm.persist_spmv_ms / rep->iters;
    ctx->tm.spmv_samples = rep->iters;
  }
  *handled = true;
  return FC_OK;
}
