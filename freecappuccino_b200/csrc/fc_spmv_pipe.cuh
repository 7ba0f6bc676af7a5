// CSR SpMV as a TMA-fed shared-memory pipeline (sm_100a), shared by the stand-alone SpMV kernels
// (fc_spmv.cu) and the persistent DPCG kernel (fc_dpcg_persist.cu).
//
// A CTA owns a contiguous range of rows (fc_row_range: equal shares, rows with processor faces
// weighted up so that the CTAs on a partition boundary get fewer of them) and walks it in chunks
// of T rows; consecutive chunks of a CTA share most of their x neighbours, which keeps the gathers
// in the SM's L1.  For every chunk one thread asks the copy engine for three 1-D bulk copies (cp.async.bulk): the chunk's T+1 row
// offsets and its contiguous slices of `a` (FP64) and `ja` (int32), 12 B per non-zero, marked
// L2 evict-first because the matrix is streamed once per product.  The copies complete on an
// mbarrier per stage; S stages keep S-1 chunks in flight while the CTA works on one, so the HBM
// stream never drains at the block-level synchronisation that frees a stage.  The arithmetic is
// thread = row: thread t reads its row's a / ja from shared memory (stride = row length: odd on
// hexahedra, conflict-free), gathers x through L1/L2 -- the 32 lanes of a warp ask for the c-th
// neighbours of 32 consecutive cells, which a finite-volume numbering keeps (nearly) contiguous, so a
// warp gather costs one or two L1 lines instead of the ~8 of a non-zero-per-thread mapping -- and adds
// the products left to right: the order of the reference's loop (dpcg.f90:105-110), so y is
// bit-identical to the Fortran result.  No thread ever loads the matrix from global memory.
// Algorithmic HBM traffic: 12 B per non-zero + 20 B per row (ioffset 4, x 8, y 8).
#pragma once
#include "fc_reduce.cuh"
#include "fc_tma.cuh"

constexpr int FC_KB_MAX = 512;   // chunks per CTA whose boundaries are cached in shared memory

// DOT: w.y ; DOT2: w.y and y.y ; RESID: y = su - A x, sum|y| ; RESID_SK: + sum y*y/(a_ii + padd)
// DOT_FUSED (persistent DPCG, "fused p" scheme): the direction vector is never swept on its own -- the product gathers
//   p(j) = q(j) + bet * pold(j)   with q = res / (a_ii + padd) left behind by the x/r update,
// the expression of dpcg.f90:95-100 evaluated where it is used (same operands, same rounding: bit-identical); the row's
// own p goes to `pnew`, the deferred x update fi += alfp * pold rides along, red[0] = p.y.
enum { FC_MODE_SPMV = 0, FC_MODE_DOT = 1, FC_MODE_RESID = 2, FC_MODE_DOT2 = 3, FC_MODE_RESID_SK = 4, FC_MODE_DOT_FUSED = 5 };

struct fc_strip {              // processor-boundary coupling kept outside the CSR (src-parallel `apr`)
  const int *off;              // [n+1] per-row range into idx, or nullptr on a single rank
  const int *idx;              // processor-face index i (0-based, ascending per row)
  const double *apr;           // [npro]
  const unsigned char *any32;  // [ceil(n/32)] 1 when one of the 32 rows has processor faces
  int halo0;                   // x[halo0 + i] = value on the other rank
  // P2P mode: the neighbours store the halo of x themselves and raise hflag[c] to hseq
  const unsigned long long *hflag;
  unsigned long long hseq;
  int nconn;
};

struct fc_spmv_mat {
  int n;
  const int *ioffset, *ja;
  const double *a;
  // CODED pipelines: the column indices as one-byte codes, ja[k] = row + dict[jc[k]] (fc_csr.cu, fc_codes_build):
  // a finite-volume numbering has few distinct column offsets (seven on a structured hexahedral block), so the
  // index stream shrinks from 4 to 1 byte per non-zero and a product moves 9 instead of 12 bytes per non-zero
  const unsigned char *jc = nullptr;
  const int *dict = nullptr;   // [256], ascending offsets
};

struct fc_spmv_vec {
  const double *x;
  double *y;
  const double *su;      // RESID: y = su - A x
  const double *w;       // DOT / DOT2: red[0] = w.y
  const int *diag;       // RESID: adiag[r] = a[diag[r]]
  double *adiag;
  double padd;           // RESID_SK: `small` of the parallel preconditioner (src-parallel/dpcg.f90:96)
  // DOT_FUSED (x = pold); RESID_SK also stores q = res / (a_ii + padd) when `qout` is set
  const double *q = nullptr;
  double *qout = nullptr, *pnew = nullptr, *fi = nullptr;
  double bet = 0.0, alfp = 0.0;
};

// Rows [rbeg, rend) of CTA b out of G: equal shares of cost(r) = r + FC_STRIP_WEIGHT * (processor faces of
// rows < r), cut at multiples of 32 rows.  A row with a processor face costs about twice an interior one
// (strip chain in the product, remote store in the p-update), so the CTAs on a partition boundary get
// about half as many rows and finish with the others.  Every CTA computes the same cuts.
constexpr long long FC_STRIP_WEIGHT = 1;
__device__ __forceinline__ int fc_row_cut(int n, const int *soff, long long b, long long G) {
  const long long groups = ((long long)n + 31) / 32;
  if (!soff) {
    const long long r = (groups * b / G) * 32;
    return (int)(r < n ? r : n);
  }
  const long long total = (long long)n + FC_STRIP_WEIGHT * soff[n];
  const long long target = total / G * b + (total % G) * b / G;
  long long lo = 0, hi = groups;   // smallest group g with cost(32 g) >= target
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    const long long r = mid * 32 < n ? mid * 32 : n;
    if (r + FC_STRIP_WEIGHT * soff[r] >= target) hi = mid;
    else lo = mid + 1;
  }
  const long long r = lo * 32;
  return (int)(r < n ? r : n);
}
__device__ __forceinline__ void fc_row_range(int n, const int *soff, int *rbeg, int *rend) {
  *rbeg = fc_row_cut(n, soff, blockIdx.x, gridDim.x);
  *rend = blockIdx.x + 1 == gridDim.x ? n : fc_row_cut(n, soff, blockIdx.x + 1, gridDim.x);
}

template <int T, int CAP, int S, bool CODED = false>
struct fc_spmv_smem {
  double a[S][CAP];
  // column indices of the staged non-zeros: CAP int32, or (CODED) CAP + 32 one-byte codes -- a code copy starts and
  // ends on a multiple of 16 non-zeros, the copy of `a` on a multiple of 4 (at most 12 + 12 codes more), so the two
  // stages have different bases
  alignas(16) unsigned char jraw[S][CODED ? CAP + 32 : CAP * 4];
  int off[S][T + 4];
  unsigned long long full[S];
  int kb[FC_KB_MAX + 1];                // first non-zero of every chunk this CTA owns (kb[j + 1]: one past its last)
  unsigned char cs[FC_KB_MAX];          // chunk contains rows with processor faces
  int dict[CODED ? 256 : 4];            // CODED: column offset of every code
};

// N consecutive non-zeros of one row from the staged chunk, v (+|-)= a[c + u] * x[column], left to right
template <int N, bool RES, bool CODED>
__device__ __forceinline__ double fc_row_part(const double *pa, const int *pj, const unsigned char *pc, const int *dict,
                                              int c, int r, const double *x, double v) {
  int id[N];
  double av[N], xv[N];
#pragma unroll
  for (int u = 0; u < N; ++u) {
    id[u] = CODED ? r + dict[pc[c + u]] : pj[c + u];
    av[u] = pa[c + u];
  }
#pragma unroll
  for (int u = 0; u < N; ++u) xv[u] = x[id[u]];
#pragma unroll
  for (int u = 0; u < N; ++u) {
    const double t = av[u] * xv[u];
    v = RES ? v - t : v + t;
  }
  return v;
}

// the same with x already in registers (gathered one chunk ahead)
template <int N, bool RES>
__device__ __forceinline__ double fc_row_part_x(const double *pa, int c, const double (&xq)[8], double v) {
#pragma unroll
  for (int u = 0; u < N; ++u) {
    const double t = pa[c + u] * xq[u];
    v = RES ? v - t : v + t;
  }
  return v;
}

template <int T, int CAP, int S, bool CODED = false>
struct fc_spmv_pipe {
  static_assert(CAP % 16 == 0 && T % 32 == 0, "staging sizes must keep the bulk copies 16-byte aligned");
  using smem_t = fc_spmv_smem<T, CAP, S, CODED>;
  smem_t *sm;
  int rbeg, rend, nch;           // rows owned by this CTA, number of chunks
  unsigned issued, consumed;     // running chunk counters over the kernel's life (uniform over the CTA)
  int next;                      // next chunk of the current sweep to hand to the copy engine
  bool halo_pending;             // set by the caller before a sweep whose halo arrives by peer stores
  bool cta_strip;                // this CTA owns rows with processor faces
  unsigned long long pol;
  unsigned long long pol_y;      // L2 policy of the result vector's stores (evict_normal unless the caller overrides it)
  // A Krylov loop walks the same chunks once per iteration.  keep256 / 256 of the CTA's chunks, spread evenly over its
  // range so that L2 hits and HBM misses alternate through the sweep, carry pol_keep (evict_last) instead of the
  // streaming policy: as much of the matrix as the L2 can hold next to the vectors stays there from one product to
  // the next, and only the rest streams from HBM (0 = the whole matrix streams).
  int keep256;
  unsigned long long pol_keep;
  // 1 / 2: while a chunk is computed, the x values the NEXT chunk will gather are prefetched into L1 / L2 (its column
  // indices are already in shared memory when its copies have landed).  Without it every chunk exposes one full
  // round trip of its far gathers (the +-nx*ny neighbours come from L2 or, at 10 M cells, from HBM) with only three
  // chunks per SM to hide it behind.
  int xprefetch;
  __device__ __forceinline__ bool kept(int j) const { return (((j + 1) * keep256) >> 8) != ((j * keep256) >> 8); }

  __device__ __forceinline__ int row0(int j) const { return rbeg + j * T; }

  // Once per kernel.  [rbeg_, rend_) starts on a multiple of 4 rows; `soff` = per-row processor-face
  // offsets (fc_strip::off) or nullptr.  Ends with a block barrier.
  __device__ __forceinline__ void init(smem_t *s, const int *ioffset, int rbeg_, int rend_, const int *soff,
                                       const int *dict = nullptr) {
    sm = s;
    if (CODED) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) sm->dict[i] = dict[i];
    }
    rbeg = rbeg_;
    rend = rend_ > rbeg_ ? rend_ : rbeg_;
    nch = (rend - rbeg + T - 1) / T;
    issued = consumed = 0u;
    next = 0;
    halo_pending = false;
    pol = fc_policy_evict_first();
    pol_y = fc_policy_evict_normal();
    keep256 = 0;
    xprefetch = 0;
    pol_keep = pol;
    int any = 0;
    for (int j = threadIdx.x; j < nch; j += blockDim.x) {
      const int r0 = row0(j);
      const int r1 = min(rend, r0 + T);
      sm->kb[j] = ioffset[r0];
      if (j + 1 == nch) sm->kb[nch] = ioffset[r1];
      const unsigned char c = (soff && soff[r1] > soff[r0]) ? 1 : 0;
      sm->cs[j] = c;
      any |= c;
    }
    if (threadIdx.x == 0) {
      for (int s2 = 0; s2 < S; ++s2) fc_mbar_init(&sm->full[s2], 1u);
      fc_mbar_init_fence();
    }
    cta_strip = __syncthreads_or(any) != 0;
  }

  // all threads call; thread 0 talks to the copy engine
  __device__ __forceinline__ void issue_one(const fc_spmv_mat &M) {
    const int j = next++;
    const unsigned st = issued % S;
    issued++;
    if (threadIdx.x == 0) {
      const int r0 = row0(j);
      const int nr = min(T, rend - r0);
      const int k0 = sm->kb[j], k1 = sm->kb[j + 1];
      const int ka = k0 & ~3, cnt = ((k1 + 3) & ~3) - ka;
      const int kc = k0 & ~15, cntc = ((k1 + 15) & ~15) - kc;   // CODED: the code copy, 16-byte granules
      const unsigned off_bytes = ((unsigned)(nr + 1) * 4u + 15u) & ~15u;
      const bool staged = cnt <= CAP && cnt > 0;
      const unsigned jbytes = CODED ? (unsigned)cntc : (unsigned)cnt * 4u;
      fc_mbar_expect_tx(&sm->full[st], off_bytes + (staged ? (unsigned)cnt * 8u + jbytes : 0u));
      fc_bulk_g2s(sm->off[st], M.ioffset + r0, off_bytes, &sm->full[st]);
      if (staged) {
        const unsigned long long pl = kept(j) ? pol_keep : pol;
        fc_bulk_g2s(sm->a[st], M.a + ka, (unsigned)cnt * 8u, &sm->full[st], pl);
        if (CODED) fc_bulk_g2s(sm->jraw[st], M.jc + kc, jbytes, &sm->full[st], pl);
        else       fc_bulk_g2s(sm->jraw[st], M.ja + ka, jbytes, &sm->full[st], pl);
      }
    }
  }

  // start of a sweep: fill the pipeline (may be called long before sweep(); the matrix does not
  // depend on x, so a Krylov loop prefetches the next product's first chunks behind its barriers)
  __device__ __forceinline__ void prefetch(const fc_spmv_mat &M) {
    while (next < nch && next < S) issue_one(M);
  }

  // wait for copies that will never be consumed (a CTA must not exit with copies in flight)
  __device__ __forceinline__ void drain() {
    while (consumed < issued) {
      fc_mbar_wait(&sm->full[consumed % S], (consumed / S) & 1u);
      consumed++;
    }
    next = 0;
  }

  // y = A x over the CTA's rows (prefetch() must have been called for this sweep)
  template <int MODE, bool STRIP>
  __device__ __forceinline__ void sweep(const fc_spmv_mat &M, const fc_spmv_vec &V, const fc_strip &st, double &acc,
                                        double &acc2) {
    const int tid = threadIdx.x;
    constexpr bool RES = (MODE == FC_MODE_RESID || MODE == FC_MODE_RESID_SK);
    // Pipelines with three stages gather x one chunk ahead: while chunk j is computed, the x values of the first
    // eight non-zeros of this thread's row of chunk j+1 (every non-zero of a hexahedral row) are already in flight
    // into registers -- the column indices of chunk j+1 are in shared memory one stage early.  Without it every
    // chunk exposes a full round trip of its far gathers (ncu: half of the stall samples of the product are this
    // scoreboard wait, a quarter the block barrier behind it, 3 % the wait for the matrix).
    constexpr bool PIPEG = S >= 3 && MODE != FC_MODE_DOT_FUSED;
    auto gather = [&](int jj, unsigned sgx, double (&xb)[8]) -> int {
      const int rr0 = row0(jj);
      const int q0 = sm->kb[jj], q1 = sm->kb[jj + 1];
      const int qa = q0 & ~3;
      if (tid >= min(T, rend - rr0) || (((q1 + 3) & ~3) - qa) > CAP) return 0;
      const int s1 = sm->off[sgx][tid] - qa, e1 = sm->off[sgx][tid + 1] - qa;
      const int *pj1 = reinterpret_cast<const int *>(sm->jraw[sgx]);
      const unsigned char *pc1 = sm->jraw[sgx] + (qa - (q0 & ~15));
      const int n1 = min(e1 - s1, 8);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (u < n1) xb[u] = V.x[CODED ? rr0 + tid + sm->dict[pc1[s1 + u]] : pj1[s1 + u]];
      return n1;
    };
    // one chunk; xc / nc: the x values gathered ahead for it, xn / nn: where the next chunk's go.  The caller
    // alternates two register buffers (no copies: a copy would wait for the loads it is meant to hide)
    auto chunk = [&](const int j, double (&xc)[8], int &nc, double (&xn)[8], int &nn) {
      const unsigned sg = consumed % S, par = (consumed / S) & 1u;
      consumed++;
      const int r0 = row0(j);
      const int nr = min(T, rend - r0);
      const int r = r0 + tid;
      // ---- processor-boundary strip of the row, requested BEFORE the wait for the matrix chunk: the chain
      //      strip_off -> strip_idx -> (apr, halo of x) then overlaps the chunk's own loads.  Only chunks that
      //      contain rows with processor faces pay for it. ----
      int sq0 = 0, sq1 = 0;
      double sap = 0.0, sxh = 0.0;
      if (STRIP && sm->cs[j]) {
        if (halo_pending) {
          // P2P mode: the neighbours store the halo of x themselves.  The CTA waits for their flags once per
          // sweep, at its first chunk with processor faces; everything before overlaps the transfer.  One
          // thread polls (an acquire at system scope is expensive), the others follow through the barrier.
          if (tid == 0) {
            fc_spin_guard gd;
            for (int c = 0; c < st.nconn; ++c)
              while (fc_ld_relaxed_sys(st.hflag + c) < st.hseq) gd.tick();
            __threadfence_system();   // acquire: the halo values precede the flags
          }
          __syncthreads();
          halo_pending = false;
        }
        if (tid < nr) {
          sq0 = st.off[r];
          sq1 = st.off[r + 1];
          if (sq1 > sq0) {
            const int i = st.idx[sq0];
            sap = st.apr[i];
            if (MODE == FC_MODE_DOT_FUSED) {   // the neighbour sent q; its p is formed here, like every other p
              sxh = __ldcg(V.q + st.halo0 + i) + V.bet * V.x[st.halo0 + i];
              V.pnew[st.halo0 + i] = sxh;
            } else {
              sxh = __ldcg(V.x + st.halo0 + i);
            }
          }
        }
      }
      fc_mbar_wait(&sm->full[sg], par);
      if (PIPEG) {
        if (j == 0) nc = gather(0, sg, xc);   // the first chunk of a sweep pays its round trip
        nn = 0;
        if (j + 1 < nch) {
          const unsigned sg1 = consumed % S, par1 = (consumed / S) & 1u;
          fc_mbar_wait(&sm->full[sg1], par1);   // requested two chunks ago
          nn = gather(j + 1, sg1, xn);
        }
      }
      if (!PIPEG && xprefetch && j + 1 < nch) {
        // the next chunk's stage: tested once, never waited for
        const unsigned sg1 = consumed % S, par1 = (consumed / S) & 1u;
        const int nr1 = min(T, rend - (r0 + T));
        const int q0 = sm->kb[j + 1], q1 = sm->kb[j + 2];
        const int qa = q0 & ~3;
        if (tid < nr1 && (((q1 + 3) & ~3) - qa) <= CAP && fc_mbar_try_wait(&sm->full[sg1], par1)) {
          const int *pj1 = reinterpret_cast<const int *>(sm->jraw[sg1]);
          const unsigned char *pc1 = sm->jraw[sg1] + (qa - (q0 & ~15));
          const int s1 = sm->off[sg1][tid] - qa, e1 = sm->off[sg1][tid + 1] - qa;
          const int r1 = r + T;
          for (int c = s1; c < e1; ++c) {
            const int id = CODED ? r1 + sm->dict[pc1[c]] : pj1[c];
            const int far = id - r1;
            if (far > T || far < -T) {   // the near neighbours are this chunk's own gathers
              if (xprefetch == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(V.x + id));
              else                asm volatile("prefetch.global.L2 [%0];" ::"l"(V.x + id));
            }
          }
        }
      }
      const int k0 = sm->kb[j], k1 = sm->kb[j + 1];
      const int ka = k0 & ~3;
      const bool staged = (((k1 + 3) & ~3) - ka) <= CAP;
      const double *pa = sm->a[sg];
      const int *pj = reinterpret_cast<const int *>(sm->jraw[sg]);
      const unsigned char *pc = sm->jraw[sg] + (ka - (k0 & ~15));   // CODED: pc[c] belongs to pa[c]
      const int *po = sm->off[sg];
      if (tid < nr) {
        const int s = po[tid], e = po[tid + 1];
        double v = RES ? V.su[r] : 0.0;
        if (staged) {
          // thread = row: the 32 lanes of a warp read x at 32 consecutive rows' c-th neighbours, which on a
          // finite-volume numbering are (nearly) contiguous -- one or two L1 lines per warp gather
          if (MODE != FC_MODE_DOT_FUSED && (CODED || PIPEG)) {
            // Straight-line bodies for every row length (fc_row_part<N>): no predicates, no selects.  Full batches
            // of eight first, then one body of exactly the remaining length; same operands, same left-to-right
            // order.  Measured at 216^3 (profiles/r02_spmv_variants.txt): with one-byte codes the product phase of
            // the persistent kernel takes 162-166 us this way and 172-176 us with the predicated batches below; with
            // `ja` it is the other way round (185 vs 168 us), so each index format keeps the loop that suits it.
            int c = s - ka;
            const int ce = e - ka;
            if (PIPEG) {
              switch (nc) {
                case 1: v = fc_row_part_x<1, RES>(pa, c, xc, v); break;
                case 2: v = fc_row_part_x<2, RES>(pa, c, xc, v); break;
                case 3: v = fc_row_part_x<3, RES>(pa, c, xc, v); break;
                case 4: v = fc_row_part_x<4, RES>(pa, c, xc, v); break;
                case 5: v = fc_row_part_x<5, RES>(pa, c, xc, v); break;
                case 6: v = fc_row_part_x<6, RES>(pa, c, xc, v); break;
                case 7: v = fc_row_part_x<7, RES>(pa, c, xc, v); break;
                case 8: v = fc_row_part_x<8, RES>(pa, c, xc, v); break;
                default: break;
              }
              c += nc;
            }
            for (; ce - c >= 8; c += 8) v = fc_row_part<8, RES, CODED>(pa, pj, pc, sm->dict, c, r, V.x, v);
            switch (ce - c) {
              case 1: v = fc_row_part<1, RES, CODED>(pa, pj, pc, sm->dict, c, r, V.x, v); break;
              case 2: v = fc_row_part<2, RES, CODED>(pa, pj, pc, sm->dict, c, r, V.x, v); break;
              case 3: v = fc_row_part<3, RES, CODED>(pa, pj, pc, sm->dict, c, r, V.x, v); break;
              case 4: v = fc_row_part<4, RES, CODED>(pa, pj, pc, sm->dict, c, r, V.x, v); break;
              case 5: v = fc_row_part<5, RES, CODED>(pa, pj, pc, sm->dict, c, r, V.x, v); break;
              case 6: v = fc_row_part<6, RES, CODED>(pa, pj, pc, sm->dict, c, r, V.x, v); break;
              case 7: v = fc_row_part<7, RES, CODED>(pa, pj, pc, sm->dict, c, r, V.x, v); break;
              default: break;
            }
          } else {
          constexpr int U = 8;
          for (int c = s - ka; c < e - ka; c += U) {
            int id[U];
            double av[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const bool in = c + u < e - ka;
              id[u] = in ? (CODED ? r + sm->dict[pc[c + u]] : pj[c + u]) : r;
              av[u] = in ? pa[c + u] : 0.0;
            }
            if (MODE == FC_MODE_DOT_FUSED) {
              double qv[U];
#pragma unroll
              for (int u = 0; u < U; ++u) { qv[u] = V.q[id[u]]; xv[u] = V.x[id[u]]; }
#pragma unroll
              for (int u = 0; u < U; ++u) xv[u] = qv[u] + V.bet * xv[u];
            } else {
#pragma unroll
              for (int u = 0; u < U; ++u) xv[u] = V.x[id[u]];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (c + u < e - ka) {
                const double t = av[u] * xv[u];
                v = RES ? v - t : v + t;
              }
            }
          }
          }
        } else {  // a chunk too long for the staging buffers: straight from global memory
          for (int k = s; k < e; ++k) {
            const int jc = M.ja[k];
            const double xj = MODE == FC_MODE_DOT_FUSED ? V.q[jc] + V.bet * V.x[jc] : V.x[jc];
            double t = M.a[k] * xj;
            v = RES ? v - t : v + t;
          }
        }
        if (STRIP) {
          if (sq1 > sq0) {   // `apr` strip after the CSR part of the row (src-parallel/dpcg.f90:132-136)
            double t = sap * sxh;
            v = RES ? v - t : v + t;
            for (int q = sq0 + 1; q < sq1; ++q) {
              const int i = st.idx[q];
              double xh;
              if (MODE == FC_MODE_DOT_FUSED) {
                xh = __ldcg(V.q + st.halo0 + i) + V.bet * V.x[st.halo0 + i];
                V.pnew[st.halo0 + i] = xh;
              } else {
                xh = __ldcg(V.x + st.halo0 + i);
              }
              t = st.apr[i] * xh;
              v = RES ? v - t : v + t;
            }
          }
        }
        fc_st_pol(V.y + r, v, pol_y);
        if (MODE == FC_MODE_DOT || MODE == FC_MODE_DOT2) acc += V.w[r] * v;
        if (MODE == FC_MODE_DOT_FUSED) {
          const double po = V.x[r];
          const double pn = V.q[r] + V.bet * po;
          fc_st_pol(V.pnew + r, pn, pol_y);
          V.fi[r] = V.fi[r] + V.alfp * po;   // x update of the previous iteration (dpcg.f90:121-124), deferred
          acc += pn * v;
        }
        if (MODE == FC_MODE_DOT2) acc2 += v * v;
        if (RES) {
          acc += fabs(v);
          const double ad = staged ? pa[V.diag[r] - ka] : M.a[V.diag[r]];
          V.adiag[r] = ad;
          if (MODE == FC_MODE_RESID_SK) {
            const double qv = v / (ad + V.padd);
            acc2 += v * qv;
            if (V.qout) V.qout[r] = qv;
          }
        }
      }
      __syncthreads();
      if (next < nch) {
        if (tid == 0) fc_fence_proxy_async();
        issue_one(M);
      }
    };
    double xa[8], xb[8];
    int na = 0, nb = 0;
    if (PIPEG) {
      for (int j = 0; j < nch; j += 2) {
        chunk(j, xa, na, xb, nb);
        if (j + 1 < nch) chunk(j + 1, xb, nb, xa, na);
      }
    } else {
      for (int j = 0; j < nch; ++j) chunk(j, xa, na, xb, nb);
    }
    next = 0;
  }
};
