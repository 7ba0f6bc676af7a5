// Level-scheduled triangular sweeps for the DIC / DILU preconditioners.
//
// Reference: iccg.f90:77-83 (d_i = 1/(a_ii - sum_{k<i} a_ik^2 d_k)), :94-111 (forward sweep,
// z = z/(d+small), backward sweep); bicgstab.f90:68-79 (DILU diagonal), :117-136, :167-185.
// L and U are A's own strict triangles in NATURAL ordering -- reordering would change the
// preconditioner and with it the iteration counts, so the rows are only *scheduled* by
// dependency level, never renumbered.  One launch runs a whole sweep: CTAs draw a ticket,
// take the next TRI_BLOCK rows of the level-major row list, wait until every CTA of the
// previous level has published its rows (one counter per level, no grid-wide barrier) and
// then add their row left to right exactly like the Fortran loop, so z is bit-identical.
#include <cub/device/device_radix_sort.cuh>

#include "fc_internal.cuh"
#include "fc_tile_schedule.hpp"
#include "fc_tma.cuh"

constexpr int TRI_BLOCK = 128;

namespace {

__global__ void k_level_relax(const int *__restrict__ ioffset, const int *__restrict__ ja,
                              const int *__restrict__ diag, int n, int lower, int *level, int *changed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lv = level[i], nl = lv;
  int s = lower ? ioffset[i] : diag[i] + 1;
  int e = lower ? diag[i] : ioffset[i + 1];
  for (int k = s; k < e; ++k) {
    int j = ja[k];
    if (j < n) nl = max(nl, level[j] + 1);
  }
  if (nl != lv) {
    level[i] = nl;
    *changed = 1;
  }
}

__global__ void k_iota(int *p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

__global__ void k_level_hist(const int *__restrict__ level, int n, int *hist) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&hist[level[i]], 1);
}

// scatter the level-sorted rows into the padded slot list
__global__ void k_level_place(const int *__restrict__ sorted_level, const int *__restrict__ sorted_row, int n,
                              const int *__restrict__ lev_start, const int *__restrict__ lev_slot, int *rows) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lv = sorted_level[i];
  rows[lev_slot[lv] + (i - lev_start[lv])] = sorted_row[i];
}

__global__ void k_fill(int *p, int v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- point-to-point mode: which blocks does a block gather from? ----
__global__ void k_slot_of_row(const int *__restrict__ rows, int nslots, int *slot_of_row) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nslots && rows[s] >= 0) slot_of_row[rows[s]] = s;
}

// one CTA per block: the set of blocks holding the rows its own rows depend on (lock-free insertion into a small
// shared table; more than FC_TRI_MAXP distinct producers raises `overflow` and the level counters stay in use)
__global__ void __launch_bounds__(TRI_BLOCK)
k_block_producers(const int *__restrict__ rows, const int *__restrict__ slot_of_row, const int *__restrict__ ioffset,
                  const int *__restrict__ ja, const int *__restrict__ diag, int n, int lower, int *prod, int *prod_cnt,
                  int *overflow) {
  __shared__ int s_set[FC_TRI_MAXP];
  __shared__ int s_over;
  const int b = blockIdx.x;
  if (threadIdx.x < FC_TRI_MAXP) s_set[threadIdx.x] = -1;
  if (threadIdx.x == 0) s_over = 0;
  __syncthreads();
  const int row = rows[b * TRI_BLOCK + threadIdx.x];
  if (row >= 0) {
    const int s = lower ? ioffset[row] : diag[row] + 1;
    const int e = lower ? diag[row] : ioffset[row + 1];
    for (int k = s; k < e; ++k) {
      const int j = ja[k];
      if (j >= n) continue;
      const int pb = slot_of_row[j] / TRI_BLOCK;
      bool placed = false;
      for (int i = 0; i < FC_TRI_MAXP && !placed; ++i) {
        const int old = atomicCAS(&s_set[i], -1, pb);
        placed = (old == -1 || old == pb);
      }
      if (!placed) s_over = 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int i = 0; i < FC_TRI_MAXP; ++i)
      if (s_set[i] >= 0) prod[b * FC_TRI_MAXP + c++] = s_set[i];
    prod_cnt[b] = c;
    if (s_over) atomicExch(overflow, 1);
  }
}

int build_one(fc_context *ctx, fc_levels &L, int lower) {
  const int n = ctx->n, B = 256;
  int *level = nullptr, *changed = nullptr;
  FC_CHECK(fc_dev_alloc(ctx, &level, (size_t)n));
  FC_CHECK(fc_dev_alloc(ctx, &changed, 1));
  FC_CUDA(cudaMemsetAsync(level, 0, sizeof(int) * (size_t)n, ctx->stream));
  // Jacobi relaxation of level(i) = 1 + max level(dependencies): converges in nlev passes
  for (int pass = 0;; ++pass) {
    FC_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), ctx->stream));
    for (int rep = 0; rep < 16; ++rep) {
      k_level_relax<<<fc_blocks(n, B), B, 0, ctx->stream>>>(ctx->ioffset, ctx->ja, ctx->diag, n, lower, level, changed);
      FC_LAUNCH_CHECK();
    }
    int h = 0;
    FC_CUDA(cudaMemcpyAsync(&h, changed, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!h) break;
    if (pass > (n / 16) + 2) FC_FAIL(FC_ERR_ARG, "level schedule did not converge");
  }
  // stable sort of the row ids by level -> ascending rows inside a level
  int *rowid = nullptr, *slevel = nullptr, *srow = nullptr;
  FC_CHECK(fc_dev_alloc(ctx, &rowid, (size_t)n));
  FC_CHECK(fc_dev_alloc(ctx, &slevel, (size_t)n));
  FC_CHECK(fc_dev_alloc(ctx, &srow, (size_t)n));
  k_iota<<<fc_blocks(n, B), B, 0, ctx->stream>>>(rowid, n);
  FC_LAUNCH_CHECK();
  void *tmp = nullptr;
  size_t bytes = 0;
  FC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, level, slevel, rowid, srow, n, 0, 32, ctx->stream));
  FC_CUDA(cudaMalloc(&tmp, bytes ? bytes : 1));
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, bytes, level, slevel, rowid, srow, n, 0, 32, ctx->stream);
  int nlev = 0;
  cudaMemcpyAsync(&nlev, slevel + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(tmp);
  FC_CUDA(e);
  ctx->launches++;
  nlev += 1;
  int *hist = nullptr;
  FC_CHECK(fc_dev_alloc(ctx, &hist, (size_t)nlev));
  FC_CUDA(cudaMemsetAsync(hist, 0, sizeof(int) * (size_t)nlev, ctx->stream));
  k_level_hist<<<fc_blocks(n, B), B, 0, ctx->stream>>>(level, n, hist);
  FC_LAUNCH_CHECK();
  std::vector<int> h(nlev), start(nlev + 1), slot(nlev + 1), blkb(nlev + 1);
  FC_CUDA(cudaMemcpyAsync(h.data(), hist, sizeof(int) * (size_t)nlev, cudaMemcpyDeviceToHost, ctx->stream));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  start[0] = slot[0] = blkb[0] = 0;
  for (int l = 0; l < nlev; ++l) {
    int nb = (h[l] + TRI_BLOCK - 1) / TRI_BLOCK;
    start[l + 1] = start[l] + h[l];
    blkb[l + 1] = blkb[l] + nb;
    slot[l + 1] = slot[l] + nb * TRI_BLOCK;
  }
  const int nblocks = blkb[nlev];
  std::vector<int> blev(nblocks);
  for (int l = 0; l < nlev; ++l)
    for (int b = blkb[l]; b < blkb[l + 1]; ++b) blev[b] = l;
  fc_levels_free(L);
  L.nlev = nlev;
  L.nslots = slot[nlev];
  int *dstart = nullptr, *dslot = nullptr;
  FC_CHECK(fc_dev_alloc(ctx, &dstart, (size_t)nlev + 1));
  FC_CHECK(fc_dev_alloc(ctx, &dslot, (size_t)nlev + 1));
  FC_CHECK(fc_dev_alloc(ctx, &L.rows, (size_t)L.nslots));
  FC_CHECK(fc_dev_alloc(ctx, &L.blk_level, (size_t)nblocks));
  FC_CHECK(fc_dev_alloc(ctx, &L.lev_blocks_before, (size_t)nlev + 1));
  FC_CHECK(fc_dev_alloc(ctx, &L.done, (size_t)nlev));
  FC_CHECK(fc_dev_alloc(ctx, &L.ready, (size_t)nlev));
  FC_CHECK(fc_dev_alloc(ctx, &L.ticket, 1));
  FC_CUDA(cudaMemcpyAsync(dstart, start.data(), sizeof(int) * (nlev + 1), cudaMemcpyHostToDevice, ctx->stream));
  FC_CUDA(cudaMemcpyAsync(dslot, slot.data(), sizeof(int) * (nlev + 1), cudaMemcpyHostToDevice, ctx->stream));
  FC_CUDA(cudaMemcpyAsync(L.blk_level, blev.data(), sizeof(int) * (size_t)nblocks, cudaMemcpyHostToDevice, ctx->stream));
  FC_CUDA(cudaMemcpyAsync(L.lev_blocks_before, blkb.data(), sizeof(int) * (nlev + 1), cudaMemcpyHostToDevice,
                          ctx->stream));
  k_fill<<<fc_blocks(L.nslots, B), B, 0, ctx->stream>>>(L.rows, -1, L.nslots);
  FC_LAUNCH_CHECK();
  k_level_place<<<fc_blocks(n, B), B, 0, ctx->stream>>>(slevel, srow, n, dstart, dslot, L.rows);
  FC_LAUNCH_CHECK();
  FC_CUDA(cudaMemsetAsync(L.done, 0, sizeof(unsigned int) * (size_t)nlev, ctx->stream));
  FC_CUDA(cudaMemsetAsync(L.ready, 0, sizeof(unsigned int) * (size_t)nlev, ctx->stream));
  FC_CUDA(cudaMemsetAsync(L.ticket, 0, sizeof(unsigned int), ctx->stream));
  L.epoch = 0;
  // producer tables of the point-to-point mode
  L.nblocks = nblocks;
  int *slot_of_row = nullptr, *overflow = nullptr, h_over = 0;
  FC_CHECK(fc_dev_alloc(ctx, &slot_of_row, (size_t)n));
  FC_CHECK(fc_dev_alloc(ctx, &overflow, 1));
  FC_CHECK(fc_dev_alloc(ctx, &L.prod, (size_t)nblocks * FC_TRI_MAXP));
  FC_CHECK(fc_dev_alloc(ctx, &L.prod_cnt, (size_t)nblocks));
  FC_CHECK(fc_dev_alloc(ctx, &L.flag, (size_t)nblocks));
  FC_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), ctx->stream));
  FC_CUDA(cudaMemsetAsync(L.flag, 0, sizeof(unsigned int) * (size_t)nblocks, ctx->stream));
  k_slot_of_row<<<fc_blocks(L.nslots, B), B, 0, ctx->stream>>>(L.rows, L.nslots, slot_of_row);
  FC_LAUNCH_CHECK();
  k_block_producers<<<nblocks, TRI_BLOCK, 0, ctx->stream>>>(L.rows, slot_of_row, ctx->ioffset, ctx->ja, ctx->diag, n,
                                                            lower, L.prod, L.prod_cnt, overflow);
  FC_LAUNCH_CHECK();
  FC_CUDA(cudaMemcpyAsync(&h_over, overflow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  L.p2p_ok = (h_over == 0);
  cudaFree(slot_of_row); cudaFree(overflow);
  cudaFree(level); cudaFree(changed); cudaFree(rowid); cudaFree(slevel); cudaFree(srow); cudaFree(hist);
  cudaFree(dstart); cudaFree(dslot);
  return FC_OK;
}

enum { TRI_FWD = 0, TRI_BWD = 1, TRI_DIC = 2, TRI_DIC_PAR = 3, TRI_DILU = 4 };

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned int *p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int atom_add_acq_rel(unsigned int *p, unsigned int v) {
  unsigned int old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

constexpr int TRI_PRE = 4;  // matrix entries of a row fetched before the wait (hex rows have 3 per triangle)

// One sweep.  `in`: r (FWD) or the forward result t (BWD); `out`: t (FWD), z (BWD), d (factor modes).
// Dependency chain per level: poll `ready[lev-1]` -> gather z of earlier rows from L2 -> row sum ->
// store -> CTA barrier -> one acq_rel atomic per CTA; the CTA that completes the level publishes
// ready[lev].  Everything that does not depend on other rows (row bounds, d, r and the first TRI_PRE
// matrix entries) is already in registers when the wait ends.
// P2P = true (FC_TUNE_SWEEP_P2P): instead of the level counter a block waits for the flags of the (at most FC_TRI_MAXP)
// blocks it gathers from and publishes its own flag -- no atomic on the critical path and no barrier across a level.
// Tickets are drawn in block order and a block only depends on blocks with smaller numbers, so a resident block never
// waits for one that has not started; the waits are bounded (fc_spin_guard traps) all the same.
template <int MODE, bool P2P>
__global__ void __launch_bounds__(TRI_BLOCK)
k_tri_sweep(const int *__restrict__ rows, const int *__restrict__ blk_level,
            const int *__restrict__ lev_blocks_before, unsigned int *done, unsigned int *ready, unsigned int *ticket,
            const int *__restrict__ prod, const int *__restrict__ prod_cnt, unsigned int *flag,
            unsigned int ticket_base, unsigned int sweep_no, const int *__restrict__ ioffset,
            const int *__restrict__ ja, const int *__restrict__ diag, const int *__restrict__ tpos,
            const double *__restrict__ a, const double *__restrict__ d, const double *__restrict__ in,
            double *out, double small, double padd, int n, const fc_scalars *sc) {
  __shared__ unsigned int s_b;
  if (sc && sc->done) return;
  if (threadIdx.x == 0) s_b = atomicAdd(ticket, 1u) - ticket_base;
  __syncthreads();
  const unsigned int b = s_b;
  const int lev = blk_level[b];
  const int row = rows[b * TRI_BLOCK + threadIdx.x];
  int s = 0, e = 0;
  double v = 0.0, di = 0.0;
  double pa[TRI_PRE], pt[TRI_PRE];
  int pj[TRI_PRE];
  if (row >= 0) {
    if (MODE == TRI_BWD) { s = diag[row] + 1; e = ioffset[row + 1]; }
    else { s = ioffset[row]; e = diag[row]; }
#pragma unroll
    for (int q = 0; q < TRI_PRE; ++q) {
      const int k = s + q;
      if (k < e) {
        pa[q] = a[k];
        pj[q] = ja[k];
        if (MODE == TRI_DILU) pt[q] = a[tpos[k]];
      }
    }
    if (MODE == TRI_FWD) { v = in[row]; di = d[row]; }
    else if (MODE == TRI_BWD) { di = d[row]; v = in[row] / (di + small); }   // z = z/(d+small), iccg.f90:102
    else v = a[diag[row]];
  }
  if (P2P) {
    const int np = prod_cnt[b];
    if ((int)threadIdx.x < np) {
      const unsigned int *r = flag + prod[b * FC_TRI_MAXP + threadIdx.x];
      fc_spin_guard g;
      while (ld_acquire(r) < sweep_no) g.tick();
    }
    if (np > 0) __syncthreads();
  } else if (lev > 0) {
    if (threadIdx.x == 0) {
      const unsigned int *r = ready + (lev - 1);
      while (ld_acquire(r) < sweep_no) {}
    }
    __syncthreads();
  }
  if (row >= 0) {
    double zq[TRI_PRE];
#pragma unroll
    for (int q = 0; q < TRI_PRE; ++q)
      if (s + q < e) zq[q] = __ldcg(out + pj[q]);   // written by another CTA during this launch: bypass L1
#pragma unroll
    for (int q = 0; q < TRI_PRE; ++q) {
      if (s + q < e) {
        const double ak = pa[q], zj = zq[q];
        if (MODE == TRI_FWD || MODE == TRI_BWD) v = v - ak * zj;
        else if (MODE == TRI_DIC) v = v - (ak * ak) * zj;            // iccg.f90:80
        else if (MODE == TRI_DIC_PAR) v = v - ak * zj * ak;          // src-parallel/iccg.f90:97
        else v = v - ak * zj * pt[q];                                // bicgstab.f90:76
      }
    }
    for (int k = s + TRI_PRE; k < e; ++k) {                          // long rows (polyhedral cells)
      const int j = ja[k];
      const double zj = __ldcg(out + j);
      const double ak = a[k];
      if (MODE == TRI_FWD || MODE == TRI_BWD) v = v - ak * zj;
      else if (MODE == TRI_DIC) v = v - (ak * ak) * zj;
      else if (MODE == TRI_DIC_PAR) v = v - ak * zj * ak;
      else v = v - ak * zj * a[tpos[k]];
    }
    if (MODE == TRI_FWD || MODE == TRI_BWD) out[row] = v * di;
    else out[row] = 1.0 / (v + padd);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (P2P) {
      st_release(flag + b, sweep_no);   // this CTA's rows are ordered before it by the barrier above
    } else {
      // release this CTA's rows; the CTA that completes the level has acquired every other CTA's release
      const unsigned int nb = (unsigned int)(lev_blocks_before[lev + 1] - lev_blocks_before[lev]);
      const unsigned int old = atom_add_acq_rel(done + lev, 1u);
      if (old + 1u == sweep_no * nb) st_release(ready + lev, sweep_no);
    }
  }
}

#ifdef FC_SWEEP_TRACE
__device__ unsigned long long *fct_trace_buf = nullptr;
#define FCT_TRACE(i)                                                                                       \
  do {                                                                                                     \
    if (fct_trace_buf && threadIdx.x == 0 && (b & 15u) == 0u) fct_trace_buf[(size_t)(b >> 4) * 12 + (i)] = fc_globaltimer(); \
  } while (0)
#define FCT_TRACE_AT(t, i)                                                                                 \
  do {                                                                                                     \
    if (fct_trace_buf && threadIdx.x == (t) && (b & 15u) == 0u) fct_trace_buf[(size_t)(b >> 4) * 12 + (i)] = fc_globaltimer(); \
  } while (0)
#endif
static_assert(FC_TILE_MAXP == FC_TRI_MAXP, "one producer-table width");
#include "fc_tile_sweep.cuh"   // k_tile_sweep<MODE, PRE, P2P>

#ifdef FC_SWEEP_TRACE
// Measurement aid (compiled only with -DFC_SWEEP_TRACE, never in the shipped library): the point-to-point tiled
// forward sweep with %globaltimer stamps of every 16th tile -- ticket drawn, descriptors loaded, coefficients loaded,
// producers' flags seen, out-of-tile values loaded, walk done, flag published.
__global__ void __launch_bounds__(FC_TILE, 2)
k_tile_sweep_trace(const int4 *__restrict__ meta, const int *__restrict__ blk_nlev, unsigned int *ticket,
                   const int *__restrict__ prod, const int *__restrict__ prod_cnt, unsigned int *flag,
                   unsigned int ticket_base, unsigned int sweep_no, const int *__restrict__ tja,
                   const double *__restrict__ a, const double *__restrict__ d, const double *__restrict__ in,
                   double *out, unsigned long long *trace) {
  constexpr int PRE = 4;
  __shared__ double s_z[FC_TILE];
  __shared__ unsigned int s_b;
  const unsigned long long t0 = fc_globaltimer();
  if (threadIdx.x == 0) s_b = atomicAdd(ticket, 1u) - ticket_base;
  __syncthreads();
  const unsigned int b = s_b;
  const bool tr = threadIdx.x == 0 && (b & 15u) == 0u;
  unsigned long long *T = trace + (size_t)(b >> 4) * 12;
  if (tr) { T[0] = t0; T[1] = fc_globaltimer(); }
  const int nl = blk_nlev[b];
  const int4 mt = meta[(size_t)b * FC_TILE + threadIdx.x];
  const int row = mt.x, my = mt.y, s = mt.z, e = mt.w;
  double v = 0.0, di = 0.0;
  double pa[PRE], zq[PRE];
  int pj[PRE];
  if (tr) T[2] = fc_globaltimer() + (unsigned long long)(row & 0);
  if (row >= 0) {
#pragma unroll
    for (int q = 0; q < PRE; ++q) {
      const int k = s + q;
      if (k < e) { pa[q] = a[k]; pj[q] = tja[k]; }
    }
    v = in[row]; di = d[row];
  }
  __syncthreads();
  if (tr) T[3] = fc_globaltimer() + (unsigned long long)(__double_as_longlong(v) & 0);
  const int np = prod_cnt[b];
  if ((int)threadIdx.x < np) {
    const unsigned int *r = flag + prod[b * FC_TILE_MAXP + threadIdx.x];
    fc_spin_guard g;
    while (ld_acquire(r) < sweep_no) g.tick();
  }
  if (np > 0) __syncthreads();
  if (tr) T[4] = fc_globaltimer();
  if (row >= 0) {
#pragma unroll
    for (int q = 0; q < PRE; ++q)
      if (s + q < e && pj[q] >= 0) zq[q] = __ldcg(out + pj[q]);
  }
  __syncthreads();
  if (tr) T[5] = fc_globaltimer() + (unsigned long long)(row >= 0 && s < e && pj[0] >= 0 ? (__double_as_longlong(zq[0]) & 0) : 0);
  for (int l = 0; l < nl; ++l) {
    if (my == l) {
#pragma unroll
      for (int q = 0; q < PRE; ++q) {
        if (s + q < e) {
          const double ak = pa[q], zj = pj[q] < 0 ? s_z[-pj[q] - 1] : zq[q];
          v = v - ak * zj;
        }
      }
      const double r = v * di;
      s_z[threadIdx.x] = r;
      out[row] = r;
    }
    __syncthreads();
    if (tr && l == 7) T[8] = fc_globaltimer();
  }
  if (tr) T[6] = fc_globaltimer();
  if (threadIdx.x == 0) st_release(flag + b, sweep_no);
  if (tr) { T[7] = fc_globaltimer(); T[9] = (unsigned long long)nl; T[10] = (unsigned long long)np; T[11] = b; }
}
#endif

// FC_TUNE_SWEEP_CHECK: rows whose bits differ between two sweeps of the same input (count, lowest row)
__global__ void k_sweep_compare(int n, const double *__restrict__ x, const double *__restrict__ y, unsigned int *bad,
                                const fc_scalars *sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (sc && sc->done)) return;   // a solve that has converged skips both sweeps: nothing to compare
  if (__double_as_longlong(x[i]) != __double_as_longlong(y[i])) {
    atomicAdd(bad, 1u);
    atomicMin(bad + 1, (unsigned int)i);
  }
}

constexpr int FC_STAGE_THREADS = 256;   // k_tile_walk: threads that stage a tile ...
constexpr int FC_WALK_THREADS = 64;     // ... and threads that walk it

template <int MODE, int PRE, bool FLAGS, int ST = FC_STAGE_THREADS>
int launch_walk(fc_context *ctx, fc_levels &T, unsigned int tbase, const double *a, const double *d, const double *in,
                double *out, double small, double padd, bool guarded) {
  auto kern = k_tile_walk<MODE, PRE, ST, FC_WALK_THREADS, ST == 256 ? 3 : 4, FLAGS>;
  constexpr size_t smem = fct_walk_layout<MODE, PRE>::bytes;
  static bool set[FC_MAX_DEVICES];   // per instantiation and per device: the opt-in above 48 KB is a per-device attribute
  const int dev = ctx->device >= 0 && ctx->device < FC_MAX_DEVICES ? ctx->device : 0;
  if (!set[dev] || dev != ctx->device) {
    FC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set[dev] = true;
  }
  const fct_handover H{T.blk_level, T.lev_blocks_before, T.prod, T.prod_cnt, T.done, T.ready, T.flag, (unsigned int)T.epoch};
  kern<<<T.nblocks, ST, smem, ctx->stream>>>(T.meta_rm, T.blk_nlev, T.ticket, tbase, ctx->tja, ctx->diag,
                                                           ctx->tpos, a, d, in, out, small, padd,
                                                           guarded ? ctx->sc : nullptr, H);
  return FC_OK;
}

// k_tile_walk_vf: 256 staging / helper threads + 64 walkers, the tile's shared layout plus one counter per local level
template <int MODE, int PRE, int OCC = 3>
int launch_walk_vf(fc_context *ctx, fc_levels &T, unsigned int tbase, const double *a, const double *d, double *in_rw,
                   double *out, double *arm, double small, double padd, bool guarded, bool rearm_in) {
  auto kern = k_tile_walk_vf<MODE, PRE, FC_STAGE_THREADS, FC_WALK_THREADS, OCC>;
  static const unsigned int backoff = getenv("FC_SWEEP_BACKOFF_NS") ? (unsigned int)atoi(getenv("FC_SWEEP_BACKOFF_NS")) : 256u;
  constexpr size_t smem = ((fct_walk_layout<MODE, PRE>::bytes + 15) & ~(size_t)15) + sizeof(int) * (FC_TILE + 2);
  static bool set[FC_MAX_DEVICES];
  const int dev = ctx->device >= 0 && ctx->device < FC_MAX_DEVICES ? ctx->device : 0;
  if (!set[dev] || dev != ctx->device) {
    FC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set[dev] = true;
  }
  kern<<<T.nblocks, FC_STAGE_THREADS + FC_WALK_THREADS, smem, ctx->stream>>>(
      T.meta_rm, T.blk_nlev, T.ticket, tbase, ctx->tja, ctx->diag, ctx->tpos, a, d, in_rw, out, arm, small, padd,
      guarded ? ctx->sc : nullptr, rearm_in, backoff);
  return FC_OK;
}

// `arm` (forward sweep of the value-as-flag mode): the vector the following backward sweep writes, see fc_tile_sweep.cuh
template <int MODE>
int sweep(fc_context *ctx, fc_levels &L, const double *a, const double *d, const double *in, double *out,
          double small, double padd, bool guarded, double *arm = nullptr) {
  if (ctx->tune_sweep_tiled && ctx->tiles_ok) {
    fc_levels &T = (&L == &ctx->lower) ? ctx->tile_lower : ctx->tile_upper;
    const unsigned int tbase = (unsigned int)(T.epoch * (unsigned long long)T.nblocks);
    T.epoch++;
    const bool p2p = ctx->tune_sweep_tiled == 2 && T.p2p_ok;
    if (ctx->tune_sweep_tiled == 4) {
      // k_tile_walk: the tile staged in shared memory by 256 threads, walked by 64; flags of the producer tiles, or the
      // tile-level counters when a tile names more than FC_TILE_MAXP producers
#define FC_WALK4(PRE_)                                                                                             \
  do {                                                                                                             \
    if (!T.p2p_ok) FC_CHECK((launch_walk<MODE, PRE_, false>(ctx, T, tbase, a, d, in, out, small, padd, guarded)));  \
    else if (ctx->tune_tile_ctas == 3)                                                                             \
      FC_CHECK((launch_walk<MODE, PRE_, true, 128>(ctx, T, tbase, a, d, in, out, small, padd, guarded)));           \
    else FC_CHECK((launch_walk<MODE, PRE_, true>(ctx, T, tbase, a, d, in, out, small, padd, guarded)));             \
  } while (0)
#ifdef FC_SWEEP_TRACE
      unsigned long long *trbuf = nullptr;
      const bool tracing = MODE == TRI_FWD && getenv("FC_SWEEP_TRACE_FILE") && T.epoch == 5;
      const size_t ntr = ((size_t)T.nblocks / 16 + 1) * 12;
      if (tracing) {
        FC_CUDA(cudaMalloc((void **)&trbuf, ntr * 8));
        FC_CUDA(cudaMemsetAsync(trbuf, 0, ntr * 8, ctx->stream));
        FC_CUDA(cudaMemcpyToSymbolAsync(fct_trace_buf, &trbuf, sizeof(trbuf), 0, cudaMemcpyHostToDevice, ctx->stream));
      }
#endif
      if (ctx->tiles_pre8) FC_WALK4(8); else if (ctx->tiles_pre3) FC_WALK4(3); else FC_WALK4(4);
#undef FC_WALK4
      FC_LAUNCH_CHECK();
#ifdef FC_SWEEP_TRACE
      if (tracing) {
        std::vector<unsigned long long> h(ntr);
        FC_CUDA(cudaMemcpyAsync(h.data(), trbuf, ntr * 8, cudaMemcpyDeviceToHost, ctx->stream));
        unsigned long long *nullp = nullptr;
        FC_CUDA(cudaMemcpyToSymbolAsync(fct_trace_buf, &nullp, sizeof(nullp), 0, cudaMemcpyHostToDevice, ctx->stream));
        FC_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(trbuf);
        if (FILE *fp = fopen(getenv("FC_SWEEP_TRACE_FILE"), "w")) {
          for (size_t i = 0; i + 12 <= ntr; i += 12) {
            for (int c = 0; c < 12; ++c) fprintf(fp, "%llu ", h[i + c]);
            fprintf(fp, "\n");
          }
          fclose(fp);
        }
      }
#endif
    } else if (ctx->tune_sweep_tiled == 5) {
      // staged walk + value-as-flag hand-over (k_tile_walk_vf); the arming of the vectors is that of mode 3 below
      const bool check = ctx->tune_sweep_check != 0;
      if (MODE == TRI_FWD || MODE == TRI_BWD) {
        if (ctx->vf_armed != out) FC_CUDA(cudaMemsetAsync(out, 0xFF, sizeof(double) * (size_t)ctx->n, ctx->stream));
      } else {
        FC_CUDA(cudaMemsetAsync(out, 0xFF, sizeof(double) * (size_t)ctx->n, ctx->stream));
      }
      ctx->vf_armed = nullptr;
      double *in_rw = const_cast<double *>(in);
      double *armv = MODE == TRI_FWD ? arm : nullptr;
#ifdef FC_SWEEP_TRACE
      unsigned long long *trbuf = nullptr;
      const bool tracing = MODE == TRI_FWD && getenv("FC_SWEEP_TRACE_FILE") && T.epoch == 5;
      const size_t ntr = ((size_t)T.nblocks / 16 + 1) * 12;
      if (tracing) {
        FC_CUDA(cudaMalloc((void **)&trbuf, ntr * 8));
        FC_CUDA(cudaMemsetAsync(trbuf, 0, ntr * 8, ctx->stream));
        FC_CUDA(cudaMemcpyToSymbolAsync(fct_trace_buf, &trbuf, sizeof(trbuf), 0, cudaMemcpyHostToDevice, ctx->stream));
      }
#endif
      if (ctx->tiles_pre8) FC_CHECK((launch_walk_vf<MODE, 8>(ctx, T, tbase, a, d, in_rw, out, armv, small, padd, guarded, !check)));
      else if (ctx->tiles_pre3 && ctx->tune_tile_ctas == 4) FC_CHECK((launch_walk_vf<MODE, 3, 4>(ctx, T, tbase, a, d, in_rw, out, armv, small, padd, guarded, !check)));
      else if (ctx->tiles_pre3 && ctx->tune_tile_ctas == 5) FC_CHECK((launch_walk_vf<MODE, 3, 5>(ctx, T, tbase, a, d, in_rw, out, armv, small, padd, guarded, !check)));
      else if (ctx->tiles_pre3 && ctx->tune_tile_ctas == 6) FC_CHECK((launch_walk_vf<MODE, 3, 6>(ctx, T, tbase, a, d, in_rw, out, armv, small, padd, guarded, !check)));
      else if (ctx->tiles_pre3) FC_CHECK((launch_walk_vf<MODE, 3>(ctx, T, tbase, a, d, in_rw, out, armv, small, padd, guarded, !check)));
      else FC_CHECK((launch_walk_vf<MODE, 4>(ctx, T, tbase, a, d, in_rw, out, armv, small, padd, guarded, !check)));
      FC_LAUNCH_CHECK();
#ifdef FC_SWEEP_TRACE
      if (tracing) {
        std::vector<unsigned long long> h(ntr);
        FC_CUDA(cudaMemcpyAsync(h.data(), trbuf, ntr * 8, cudaMemcpyDeviceToHost, ctx->stream));
        unsigned long long *nullp = nullptr;
        FC_CUDA(cudaMemcpyToSymbolAsync(fct_trace_buf, &nullp, sizeof(nullp), 0, cudaMemcpyHostToDevice, ctx->stream));
        FC_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(trbuf);
        if (FILE *fp = fopen(getenv("FC_SWEEP_TRACE_FILE"), "w")) {
          for (size_t i = 0; i + 12 <= ntr; i += 12) {
            for (int c = 0; c < 12; ++c) fprintf(fp, "%llu ", h[i + c]);
            fprintf(fp, "\n");
          }
          fclose(fp);
        }
      }
#endif
      if (MODE == TRI_FWD && arm) ctx->vf_armed = arm;
      if (MODE == TRI_BWD && !check) ctx->vf_armed = in_rw;
    } else if (ctx->tune_sweep_tiled == 3) {
      // value-as-flag hand-over: `out` must be all "unset" when the kernel starts.  The factor sweeps and a forward
      // sweep whose target has not been re-armed by the previous backward sweep pay one memset; in the steady state
      // of a solve the sweeps re-arm each other's vectors (FWD arms z, BWD re-arms t) and no pass is added.
      const bool check = ctx->tune_sweep_check != 0;
      if (MODE == TRI_FWD || MODE == TRI_BWD) {
        if (ctx->vf_armed != out) FC_CUDA(cudaMemsetAsync(out, 0xFF, sizeof(double) * (size_t)ctx->n, ctx->stream));
      } else {
        FC_CUDA(cudaMemsetAsync(out, 0xFF, sizeof(double) * (size_t)ctx->n, ctx->stream));
      }
      ctx->vf_armed = nullptr;
      double *in_rw = const_cast<double *>(in);
      // the comparison sweep of FC_TUNE_SWEEP_CHECK reads the same input afterwards: do not overwrite it then
      const double vf_small = small;
#define FC_VF_LAUNCH_OCC(PRE_, OCC_)                                                                           \
  k_tile_sweep_vf<MODE, PRE_, OCC_><<<T.nblocks, FC_TILE, 0, ctx->stream>>>(                                    \
      T.meta, T.blk_nlev, T.ticket, tbase, ctx->tja, ctx->diag, ctx->tpos, a, d, in_rw, out,                    \
      MODE == TRI_FWD ? arm : nullptr, vf_small, padd, guarded ? ctx->sc : nullptr, !check)
#define FC_VF_LAUNCH(PRE_)                                       \
  do {                                                           \
    if (ctx->tune_tile_ctas == 3) FC_VF_LAUNCH_OCC(PRE_, 3);     \
    else FC_VF_LAUNCH_OCC(PRE_, 2);                              \
  } while (0)
      if (ctx->tiles_pre8) FC_VF_LAUNCH(8); else FC_VF_LAUNCH(4);
#undef FC_VF_LAUNCH
#undef FC_VF_LAUNCH_OCC
      FC_LAUNCH_CHECK();
      if (MODE == TRI_FWD && arm) ctx->vf_armed = arm;              // z is unset: the backward sweep may start
      if (MODE == TRI_BWD && !check) ctx->vf_armed = in_rw;         // t is unset again: the next forward sweep may start
    } else
#define FC_TILE_LAUNCH_OCC(PRE_, P2P_, OCC_)                                                                         \
  k_tile_sweep<MODE, PRE_, P2P_, OCC_><<<T.nblocks, FC_TILE, 0, ctx->stream>>>(                                       \
      T.meta, T.blk_nlev, T.blk_level, T.lev_blocks_before, T.done, T.ready, T.ticket, T.prod, T.prod_cnt, T.flag,    \
      tbase, (unsigned int)T.epoch, ctx->tja, ctx->diag, ctx->tpos, a, d, in, out, small, padd,                      \
      guarded ? ctx->sc : nullptr)
#define FC_TILE_LAUNCH(PRE_, P2P_)                                            \
  do {                                                                        \
    if (ctx->tune_tile_ctas == 3) FC_TILE_LAUNCH_OCC(PRE_, P2P_, 3);          \
    else FC_TILE_LAUNCH_OCC(PRE_, P2P_, 2);                                   \
  } while (0)
#ifdef FC_SWEEP_TRACE
    if (MODE == TRI_FWD && p2p && !ctx->tiles_pre8 && getenv("FC_SWEEP_TRACE_FILE") && T.epoch == 5) {
      const size_t nt = ((size_t)T.nblocks / 16 + 1) * 12;
      unsigned long long *tr = nullptr;
      FC_CUDA(cudaMalloc((void **)&tr, nt * 8));
      FC_CUDA(cudaMemsetAsync(tr, 0, nt * 8, ctx->stream));
      k_tile_sweep_trace<<<T.nblocks, FC_TILE, 0, ctx->stream>>>(T.meta, T.blk_nlev, T.ticket, T.prod, T.prod_cnt, T.flag,
                                                                 tbase, (unsigned int)T.epoch, ctx->tja, a, d, in, out, tr);
      std::vector<unsigned long long> h(nt);
      FC_CUDA(cudaMemcpyAsync(h.data(), tr, nt * 8, cudaMemcpyDeviceToHost, ctx->stream));
      FC_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(tr);
      if (FILE *fp = fopen(getenv("FC_SWEEP_TRACE_FILE"), "w")) {
        for (size_t i = 0; i + 12 <= nt; i += 12) {
          for (int c = 0; c < 12; ++c) fprintf(fp, "%llu ", h[i + c]);
          fprintf(fp, "\n");
        }
        fclose(fp);
      }
    } else
#endif
    {
    if (ctx->tiles_pre8) { if (p2p) FC_TILE_LAUNCH(8, true); else FC_TILE_LAUNCH(8, false); }
    else                 { if (p2p) FC_TILE_LAUNCH(4, true); else FC_TILE_LAUNCH(4, false); }
    FC_LAUNCH_CHECK();
    }
#undef FC_TILE_LAUNCH
#undef FC_TILE_LAUNCH_OCC
    if (!ctx->tune_sweep_check) return FC_OK;
    // debugging aid: the same sweep once more with the level schedule into a scratch vector, compared bit for bit
    if (!ctx->sweep_chk) FC_CHECK(fc_dev_alloc(ctx, &ctx->sweep_chk, (size_t)ctx->n + 2));
    unsigned int *bad = reinterpret_cast<unsigned int *>(ctx->sweep_chk + ctx->n);
    const unsigned int init[2] = {0u, 0xffffffffu};
    FC_CUDA(cudaMemcpyAsync(bad, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    const int saved = ctx->tune_sweep_tiled;
    ctx->tune_sweep_tiled = 0;
    const int rc = sweep<MODE>(ctx, L, a, d, in, ctx->sweep_chk, small, padd, guarded);
    ctx->tune_sweep_tiled = saved;
    FC_CHECK(rc);
    k_sweep_compare<<<fc_blocks(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->n, out, ctx->sweep_chk, bad,
                                                                        guarded ? ctx->sc : nullptr);
    FC_LAUNCH_CHECK();
    unsigned int h[2] = {0u, 0u};
    FC_CUDA(cudaMemcpyAsync(h, bad, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    FC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h[0] > 0)
      FC_FAIL(FC_ERR_CUDA, "tiled sweep (mode " + std::to_string(MODE) + (&L == &ctx->lower ? ", lower" : ", upper") +
                               " triangle, sweep " + std::to_string(T.epoch) + ") differs from the level sweep in " +
                               std::to_string(h[0]) + " rows, first row " + std::to_string(h[1] + 1));
    return FC_OK;
  }
  const int nblocks = L.nslots / TRI_BLOCK;
  const unsigned int base = (unsigned int)(L.epoch * (unsigned long long)nblocks);
  L.epoch++;
  if (ctx->tune_sweep_p2p && L.p2p_ok)
    k_tri_sweep<MODE, true><<<nblocks, TRI_BLOCK, 0, ctx->stream>>>(
        L.rows, L.blk_level, L.lev_blocks_before, L.done, L.ready, L.ticket, L.prod, L.prod_cnt, L.flag, base,
        (unsigned int)L.epoch, ctx->ioffset, ctx->ja, ctx->diag, ctx->tpos, a, d, in, out, small, padd, ctx->n,
        guarded ? ctx->sc : nullptr);
  else
    k_tri_sweep<MODE, false><<<nblocks, TRI_BLOCK, 0, ctx->stream>>>(
        L.rows, L.blk_level, L.lev_blocks_before, L.done, L.ready, L.ticket, L.prod, L.prod_cnt, L.flag, base,
        (unsigned int)L.epoch, ctx->ioffset, ctx->ja, ctx->diag, ctx->tpos, a, d, in, out, small, padd, ctx->n,
        guarded ? ctx->sc : nullptr);
  FC_LAUNCH_CHECK();
  return FC_OK;
}

}  // namespace

void fc_levels_free(fc_levels &L) {
  cudaFree(L.rows); cudaFree(L.blk_level); cudaFree(L.lev_blocks_before); cudaFree(L.done); cudaFree(L.ready);
  cudaFree(L.ticket); cudaFree(L.prod); cudaFree(L.prod_cnt); cudaFree(L.flag); cudaFree(L.meta); cudaFree(L.blk_nlev);
  cudaFree(L.meta_rm);
  L = fc_levels{};
}

namespace {

int upload_tile_dir(fc_context *ctx, const fc_tile_dir &D, fc_levels &L) {
  fc_levels_free(L);
  L.nlev = D.nlev;
  L.nblocks = D.nblocks;
  L.nslots = D.nblocks * FC_TILE;
  FC_CHECK(fc_dev_alloc(ctx, &L.meta, D.meta.size() / 4));
  FC_CHECK(fc_dev_alloc(ctx, &L.meta_rm, D.meta_rm.size() / 4));
  FC_CHECK(fc_dev_alloc(ctx, &L.blk_nlev, D.blk_nlev.size()));
  FC_CHECK(fc_dev_alloc(ctx, &L.blk_level, D.blk_level.size()));
  FC_CHECK(fc_dev_alloc(ctx, &L.lev_blocks_before, D.lev_blocks_before.size()));
  FC_CHECK(fc_dev_alloc(ctx, &L.done, (size_t)D.nlev));
  FC_CHECK(fc_dev_alloc(ctx, &L.ready, (size_t)D.nlev));
  FC_CHECK(fc_dev_alloc(ctx, &L.ticket, 1));
  FC_CHECK(fc_dev_alloc(ctx, &L.prod, D.prod.size()));
  FC_CHECK(fc_dev_alloc(ctx, &L.prod_cnt, D.prod_cnt.size()));
  FC_CHECK(fc_dev_alloc(ctx, &L.flag, (size_t)D.nblocks));
  L.p2p_ok = D.p2p_ok;
  const struct { int *dst; const std::vector<int> *src; } up[] = {
      {(int *)L.meta, &D.meta}, {(int *)L.meta_rm, &D.meta_rm}, {L.blk_nlev, &D.blk_nlev}, {L.blk_level, &D.blk_level},
      {L.lev_blocks_before, &D.lev_blocks_before}, {L.prod, &D.prod}, {L.prod_cnt, &D.prod_cnt}};
  for (const auto &u : up)
    FC_CUDA(cudaMemcpyAsync(u.dst, u.src->data(), sizeof(int) * u.src->size(), cudaMemcpyHostToDevice, ctx->stream));
  FC_CUDA(cudaMemsetAsync(L.done, 0, sizeof(unsigned int) * (size_t)D.nlev, ctx->stream));
  FC_CUDA(cudaMemsetAsync(L.ready, 0, sizeof(unsigned int) * (size_t)D.nlev, ctx->stream));
  FC_CUDA(cudaMemsetAsync(L.ticket, 0, sizeof(unsigned int), ctx->stream));
  FC_CUDA(cudaMemsetAsync(L.flag, 0, sizeof(unsigned int) * (size_t)D.nblocks, ctx->stream));
  FC_CUDA(cudaStreamSynchronize(ctx->stream));   // the host vectors go away
  L.epoch = 0;
  return FC_OK;
}

// FC_TUNE_SWEEP_TILED: the schedule is built on the host from the pattern and the cell centres (once per mesh, about
// a second per 10 M cells); a mesh that cannot be tiled keeps the level schedule and ctx->tiles_why says why
int build_tiles(fc_context *ctx) {
  ctx->tiles_tried = true;
  ctx->tiles_ok = false;
  if (!ctx->has_mesh || !ctx->xc) {
    ctx->tiles_why = "no cell centres (explicit-CSR context)";
    return FC_OK;
  }
  const size_t n = (size_t)ctx->n, nnz = (size_t)ctx->nnz;
  std::vector<int> ioffset(n + 1), ja(nnz), diag(n);
  std::vector<double> xc(n), yc(n), zc(n);
  FC_CUDA(cudaStreamSynchronize(ctx->stream));
  FC_CUDA(cudaMemcpy(ioffset.data(), ctx->ioffset, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost));
  FC_CUDA(cudaMemcpy(ja.data(), ctx->ja, sizeof(int) * nnz, cudaMemcpyDeviceToHost));
  FC_CUDA(cudaMemcpy(diag.data(), ctx->diag, sizeof(int) * n, cudaMemcpyDeviceToHost));
  FC_CUDA(cudaMemcpy(xc.data(), ctx->xc, sizeof(double) * n, cudaMemcpyDeviceToHost));
  FC_CUDA(cudaMemcpy(yc.data(), ctx->yc, sizeof(double) * n, cudaMemcpyDeviceToHost));
  FC_CUDA(cudaMemcpy(zc.data(), ctx->zc, sizeof(double) * n, cudaMemcpyDeviceToHost));
  const fc_tile_schedule S =
      fc_build_tile_schedule(ctx->n, ioffset.data(), ja.data(), diag.data(), xc.data(), yc.data(), zc.data(),
                             // FC_TILE_MIN_SHRINK: narrower bins, a measurement knob.  On polyhedral meshes 6-cell bins
                             // measured 11 % faster at 4.2 M cells (profiles/r02_poly_bins.txt) but 12 % slower at the
                             // 20 M cells of config 5 (13.2 against 11.8 ms per ICCG iteration in the bench line), so the
                             // default stays at the widest bins that fit a tile
                             getenv("FC_TILE_MIN_SHRINK") ? atoi(getenv("FC_TILE_MIN_SHRINK")) : 0);
  ctx->tiles_why = S.why;
  if (!S.ok) return FC_OK;
  // bins cut into runs of consecutive rows (fc_tile_schedule.hpp repair_tiles) can degenerate into a long chain: keep
  // the level schedule unless the estimated critical path is clearly shorter
  const long long level_cost = fc_level_cost(ctx->lower.nlev > ctx->upper.nlev ? ctx->lower.nlev : ctx->upper.nlev);
  if (10 * S.cost > 7 * level_cost) {
    ctx->tiles_why = "tiling not worth it: estimated critical path " + std::to_string(S.cost / 10) + " us against " +
                     std::to_string(level_cost / 10) + " us of the level schedule";
    return FC_OK;
  }
  FC_CHECK(fc_dev_alloc(ctx, &ctx->tja, nnz));
  FC_CUDA(cudaMemcpy(ctx->tja, S.tja.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice));
  FC_CHECK(upload_tile_dir(ctx, S.lower, ctx->tile_lower));
  FC_CHECK(upload_tile_dir(ctx, S.upper, ctx->tile_upper));
  ctx->tiles_pre8 = S.max_tri_len > 4;
  ctx->tiles_pre3 = S.max_tri_len <= 3;   // hexahedra: one dependent subtraction less per row in k_tile_walk
  ctx->tiles_info = std::to_string(S.ntiles) + " tiles of <= " + std::to_string(S.max_tile_rows) + " rows (bins of " +
                    std::to_string(S.cells_per_axis) + " cells per axis, " + std::to_string(S.repaired_rows) +
                    " rows in cut bins), " + std::to_string(S.lower.nlev) + " / " + std::to_string(S.upper.nlev) +
                    " tile levels, <= " + std::to_string(std::max(S.lower.max_local_levels, S.upper.max_local_levels)) +
                    " local levels, <= " + std::to_string(std::max(S.lower.max_producers, S.upper.max_producers)) +
                    " producer tiles, estimated critical path " + std::to_string(S.cost / 10) + " us";
  ctx->tiles_ok = true;
  return FC_OK;
}

}  // namespace

int fc_levels_build(fc_context *ctx) {
  if (!ctx->has_levels) {
    FC_CHECK(build_one(ctx, ctx->lower, 1));
    FC_CHECK(build_one(ctx, ctx->upper, 0));
    ctx->has_levels = true;
    ctx->tiles_tried = ctx->tiles_ok = false;   // a new pattern: the old tiling is void
  }
  if (ctx->tune_sweep_tiled && !ctx->tiles_tried) FC_CHECK(build_tiles(ctx));
  return FC_OK;
}

// counters are monotone over the sweeps of one solve; restart them so that they never wrap
int fc_levels_reset(fc_context *ctx) {
  for (fc_levels *L : {&ctx->lower, &ctx->upper, &ctx->tile_lower, &ctx->tile_upper}) {
    if (!L->done) continue;   // no tiling for this mesh
    FC_CUDA(cudaMemsetAsync(L->done, 0, sizeof(unsigned int) * (size_t)L->nlev, ctx->stream));
    FC_CUDA(cudaMemsetAsync(L->ready, 0, sizeof(unsigned int) * (size_t)L->nlev, ctx->stream));
    FC_CUDA(cudaMemsetAsync(L->ticket, 0, sizeof(unsigned int), ctx->stream));
    if (L->flag) FC_CUDA(cudaMemsetAsync(L->flag, 0, sizeof(unsigned int) * (size_t)L->nblocks, ctx->stream));
    L->epoch = 0;
  }
  return FC_OK;
}

// kind: 0 DIC serial, 1 DIC src-parallel, 2 DILU
int fc_precond_factor(fc_context *ctx, int kind, const double *a, double *d, double padd) {
  ctx->vf_armed = nullptr;   // a new solve: nothing is known about the scratch vectors of the last one
  if (kind == 0) return sweep<TRI_DIC>(ctx, ctx->lower, a, nullptr, nullptr, d, 0.0, padd, false);
  if (kind == 1) return sweep<TRI_DIC_PAR>(ctx, ctx->lower, a, nullptr, nullptr, d, 0.0, padd, false);
  return sweep<TRI_DILU>(ctx, ctx->lower, a, nullptr, nullptr, d, 0.0, padd, false);
}

// z = (D+U)^-1 D (D+L)^-1 r with the reference's intermediate z/(d+small); t is scratch
int fc_precond_apply(fc_context *ctx, const double *a, const double *d, const double *r, double *t, double *z,
                     double small) {
  FC_CHECK(sweep<TRI_FWD>(ctx, ctx->lower, a, d, r, t, small, 0.0, true, z));
  return sweep<TRI_BWD>(ctx, ctx->upper, a, d, t, z, small, 0.0, true);
}
