// k_tile_sweep: the tiled DIC / DILU sweep kernel of fc_trisolve.cu (FC_TUNE_SWEEP_TILED; schedule: fc_tile_schedule.hpp).
//
// Kept in its own header, written against a handful of macros, so that tests/kernel_bodies_host/fct_host.cpp can
// compile THIS source with g++ and run it with one host thread per CUDA thread and a std::barrier for __syncthreads
// (test infrastructure; the library only ever compiles the CUDA side).  The includer provides: the TRI_* mode enum,
// ld_acquire / st_release / atom_add_acq_rel, fc_spin_guard, fc_scalars, FC_TILE / FC_TILE_MAXP.
#pragma once

#ifdef __CUDACC__
#define FCT_KERNEL(OCC) __global__ void __launch_bounds__(FC_TILE, OCC)
#define FCT_SHARED __shared__
#define FCT_TID threadIdx.x
#define FCT_SYNC() __syncthreads()
#define FCT_TICKET(p) atomicAdd((p), 1u)
#define FCT_LDCG(p) __ldcg(p)
#define FCT_UNROLL _Pragma("unroll")
// value-as-flag hand-over (k_tile_sweep_vf): relaxed 8-byte accesses that always go to L2
__device__ __forceinline__ double fct_ld_poll(const double *p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fct_st_pub(double *p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ bool fct_is_unset(double v) { return __double_as_longlong(v) == -1ll; }
__device__ __forceinline__ double fct_unset() { return __longlong_as_double(-1ll); }
#define FCT_LD_POLL(p) fct_ld_poll(p)
#define FCT_ST_PUB(p, v) fct_st_pub((p), (v))
#define FCT_WALK_KERNEL(T, OCC) __global__ void __launch_bounds__(T, OCC)
// barrier among the first N threads of the CTA only (the other warps have retired): named barrier 1
#define FCT_WALK_SYNC(N) asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory")
#define FCT_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
// k_tile_walk_vf: counters in shared memory between the helper warps (fold a value of another tile, count down) and
// the walking warps (spin until the level's counter is zero)
#define FCT_SATOMIC_ADD(p, v) atomicAdd((p), (v))
#define FCT_LD_SVOL(p) (*(const volatile int *)(p))
#define FCT_FENCE_BLOCK() __threadfence_block()
#define FCT_BACKOFF(ns) __nanosleep(ns)
#endif

// Tiled mode (FC_TUNE_SWEEP_TILED; schedule: fc_tile_schedule.hpp).  A CTA owns one spatial tile of at most FC_TILE
// rows, thread = slot.  Hand-overs through global memory happen once per TILE level (79 at 216^3 instead of 646 row
// levels), with the same done / ready counters as the level mode; inside the tile the rows are walked by local level
// with __syncthreads, and a dependency that sits in the same tile is read from shared memory (tja < 0 names its
// slot).  Each row is still summed left to right over its triangle, so the result is bit-identical.
// PRE = matrix entries of a row held in registers before the walk starts (4 covers hexahedra, 8 the 14-faced
// polyhedra): an entry fetched inside the walk would put an L2 round trip on the critical path of a local level.
// P2P (FC_TUNE_SWEEP_TILED = 2): a tile waits for the flags of the tiles it reads through global memory (3 on a
// hexahedral mesh) instead of for the whole previous tile level, so tiles run ahead where the tile graph allows.
// OCC = CTAs per SM the register allocation has to allow (2: 58-64 registers, no spills; 3: 40 registers and 36-44 B of
// spills on hexahedra) -- a measurement knob, FC_TUNE_TILE_CTAS.
template <int MODE, int PRE, bool P2P, int OCC>
FCT_KERNEL(OCC)
k_tile_sweep(const int4 *__restrict__ meta, const int *__restrict__ blk_nlev,
             const int *__restrict__ blk_level, const int *__restrict__ lev_blocks_before, unsigned int *done,
             unsigned int *ready, unsigned int *ticket, const int *__restrict__ prod,
             const int *__restrict__ prod_cnt, unsigned int *flag, unsigned int ticket_base, unsigned int sweep_no,
             const int *__restrict__ tja, const int *__restrict__ diag, const int *__restrict__ tpos, const double *__restrict__ a, const double *__restrict__ d,
             const double *__restrict__ in, double *out, double small, double padd, const fc_scalars *sc) {
  FCT_SHARED double s_z[FC_TILE];
  FCT_SHARED unsigned int s_b;
  if (sc && sc->done) return;
  if (FCT_TID == 0) s_b = FCT_TICKET(ticket) - ticket_base;
  FCT_SYNC();
  const unsigned int b = s_b;
  const int lev = blk_level[b], nl = blk_nlev[b];
  const int4 mt = meta[(size_t)b * FC_TILE + FCT_TID];   // row, local level, triangle [s, e): one 16-byte load
  const int row = mt.x, my = mt.y, s = mt.z, e = mt.w;
  double v = 0.0, di = 0.0;
  double pa[PRE], pt[PRE], zq[PRE];
  int pj[PRE];
  if (row >= 0) {
FCT_UNROLL
    for (int q = 0; q < PRE; ++q) {
      const int k = s + q;
      if (k < e) {
        pa[q] = a[k];
        pj[q] = tja[k];
        if (MODE == TRI_DILU) pt[q] = a[tpos[k]];
      }
    }
    if (MODE == TRI_FWD) { v = in[row]; di = d[row]; }
    else if (MODE == TRI_BWD) { di = d[row]; v = in[row] / (di + small); }   // z = z/(d+small), iccg.f90:102
    else v = a[diag[row]];
  }
  if (P2P) {
    const int np = prod_cnt[b];
    if ((int)FCT_TID < np) {
      const unsigned int *r = flag + prod[b * FC_TILE_MAXP + FCT_TID];
      fc_spin_guard g;
      while (ld_acquire(r) < sweep_no) g.tick();
    }
    if (np > 0) FCT_SYNC();
  } else if (lev > 0) {   // every tile of the previous tile level has published its rows
    if (FCT_TID == 0) {
      const unsigned int *r = ready + (lev - 1);
      fc_spin_guard g;      // a schedule bug must trap, not hang the device
      while (ld_acquire(r) < sweep_no) g.tick();
    }
    FCT_SYNC();
  }
  if (row >= 0) {   // rows of other tiles: complete, fetch them now (one L2 round trip for the whole tile)
FCT_UNROLL
    for (int q = 0; q < PRE; ++q)
      if (s + q < e && pj[q] >= 0) zq[q] = FCT_LDCG(out + pj[q]);
  }
  for (int l = 0; l < nl; ++l) {
    if (my == l) {
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) {
        if (s + q < e) {
          const double ak = pa[q], zj = pj[q] < 0 ? s_z[-pj[q] - 1] : zq[q];
          if (MODE == TRI_FWD || MODE == TRI_BWD) v = v - ak * zj;
          else if (MODE == TRI_DIC) v = v - (ak * ak) * zj;            // iccg.f90:80
          else if (MODE == TRI_DIC_PAR) v = v - ak * zj * ak;          // src-parallel/iccg.f90:97
          else v = v - ak * zj * pt[q];                                // bicgstab.f90:76
        }
      }
      for (int k = s + PRE; k < e; ++k) {                          // long rows (polyhedral cells)
        const int j = tja[k];
        const double zj = j < 0 ? s_z[-j - 1] : FCT_LDCG(out + j);
        const double ak = a[k];
        if (MODE == TRI_FWD || MODE == TRI_BWD) v = v - ak * zj;
        else if (MODE == TRI_DIC) v = v - (ak * ak) * zj;
        else if (MODE == TRI_DIC_PAR) v = v - ak * zj * ak;
        else v = v - ak * zj * a[tpos[k]];
      }
      const double r = (MODE == TRI_FWD || MODE == TRI_BWD) ? v * di : 1.0 / (v + padd);
      s_z[FCT_TID] = r;
      out[row] = r;
    }
    FCT_SYNC();   // the last one also orders every row's store before thread 0's release below
  }
  if (FCT_TID == 0) {
    if (P2P) {
      st_release(flag + b, sweep_no);
    } else {
      const unsigned int nb = (unsigned int)(lev_blocks_before[lev + 1] - lev_blocks_before[lev]);
      const unsigned int old = atom_add_acq_rel(done + lev, 1u);
      if (old + 1u == sweep_no * nb) st_release(ready + lev, sweep_no);
    }
  }
}


// Value-as-flag mode (FC_TUNE_SWEEP_TILED = 3): no flags, no counters, no fences.  Before the sweep every entry of
// `out` holds the "unset" pattern (all 64 bits set: a NaN no arithmetic produces; one cudaMemset byte value, and
// the sweeps re-arm each other's vectors, see below).  A row that needs a value of ANOTHER tile polls the value
// itself -- an aligned 8-byte store is atomic, so the datum is its own "ready" flag -- and it polls only when its own
// local level comes up, so a tile starts as soon as the first rows of its producers exist instead of after their last
// ones: the hand-over latency is paid once per row level on the critical path, not once per tile level with a
// release / acquire pair around the whole tile.  Tickets are drawn in tile order and a tile only reads rows of tiles
// with smaller tickets (fc_tile_schedule.hpp), so a polling CTA always waits for CTAs that are already running.
// The row sums are the same left-to-right sums, so the result is bit-identical to every other schedule.
//   `in_rw`: the input vector; with `rearm_in` the backward sweep overwrites the entry it has read with "unset" (the
//            forward sweep of the next application writes that vector again: fc_precond_apply's scratch t);
//   `arm`  : forward sweep only: the vector the backward sweep of the same application will write (z), set to
//            "unset" row by row here.
template <int MODE, int PRE, int OCC>
FCT_KERNEL(OCC)
k_tile_sweep_vf(const int4 *__restrict__ meta, const int *__restrict__ blk_nlev, unsigned int *ticket,
                unsigned int ticket_base, const int *__restrict__ tja, const int *__restrict__ diag,
                const int *__restrict__ tpos, const double *__restrict__ a, const double *__restrict__ d,
                double *in_rw, double *out, double *arm, double small, double padd, const fc_scalars *sc,
                bool rearm_in) {
  FCT_SHARED double s_z[FC_TILE];
  FCT_SHARED unsigned int s_b;
  if (sc && sc->done) return;
  if (FCT_TID == 0) s_b = FCT_TICKET(ticket) - ticket_base;
  FCT_SYNC();
  const unsigned int b = s_b;
  const int nl = blk_nlev[b];
  const int4 mt = meta[(size_t)b * FC_TILE + FCT_TID];   // row, local level, triangle [s, e): one 16-byte load
  const int row = mt.x, my = mt.y, s = mt.z, e = mt.w;
  double v = 0.0, di = 0.0;
  double pa[PRE], pt[PRE], zq[PRE];
  int pj[PRE];
  if (row >= 0) {
FCT_UNROLL
    for (int q = 0; q < PRE; ++q) {
      const int k = s + q;
      pj[q] = -1;
      if (k < e) {
        pa[q] = a[k];
        pj[q] = tja[k];
        if (MODE == TRI_DILU) pt[q] = a[tpos[k]];
      }
    }
    if (MODE == TRI_FWD) {
      v = in_rw[row]; di = d[row];
      if (arm) arm[row] = fct_unset();
    } else if (MODE == TRI_BWD) {
      di = d[row]; v = in_rw[row] / (di + small);   // z = z/(d+small), iccg.f90:102
      if (rearm_in) in_rw[row] = fct_unset();
    } else {
      v = a[diag[row]];
    }
    // first look at the rows of other tiles (one L2 round trip for the whole tile, off the critical path); what is
    // still unset is polled when the row's local level comes up
FCT_UNROLL
    for (int q = 0; q < PRE; ++q)
      if (s + q < e && pj[q] >= 0) zq[q] = FCT_LD_POLL(out + pj[q]);
  }
  for (int l = 0; l < nl; ++l) {
    if (my == l) {
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) {
        if (s + q < e) {
          double zj;
          if (pj[q] < 0) {
            zj = s_z[-pj[q] - 1];
          } else {
            zj = zq[q];
            fc_spin_guard g;
            while (fct_is_unset(zj)) { g.tick(); zj = FCT_LD_POLL(out + pj[q]); }
          }
          const double ak = pa[q];
          if (MODE == TRI_FWD || MODE == TRI_BWD) v = v - ak * zj;
          else if (MODE == TRI_DIC) v = v - (ak * ak) * zj;            // iccg.f90:80
          else if (MODE == TRI_DIC_PAR) v = v - ak * zj * ak;          // src-parallel/iccg.f90:97
          else v = v - ak * zj * pt[q];                                // bicgstab.f90:76
        }
      }
      for (int k = s + PRE; k < e; ++k) {                          // long rows (polyhedral cells)
        const int j = tja[k];
        double zj;
        if (j < 0) {
          zj = s_z[-j - 1];
        } else {
          zj = FCT_LD_POLL(out + j);
          fc_spin_guard g;
          while (fct_is_unset(zj)) { g.tick(); zj = FCT_LD_POLL(out + j); }
        }
        const double ak = a[k];
        if (MODE == TRI_FWD || MODE == TRI_BWD) v = v - ak * zj;
        else if (MODE == TRI_DIC) v = v - (ak * ak) * zj;
        else if (MODE == TRI_DIC_PAR) v = v - ak * zj * ak;
        else v = v - ak * zj * a[tpos[k]];
      }
      const double r = (MODE == TRI_FWD || MODE == TRI_BWD) ? v * di : 1.0 / (v + padd);
      s_z[FCT_TID] = r;
      FCT_ST_PUB(out + row, r);
    }
    FCT_SYNC();
  }
}


// ---------------------------------------------------------------------------------------------------------------
// k_tile_walk (FC_TUNE_SWEEP_TILED = 4): stage the tile with many threads, walk it with few.
//
// Measured on B200 (%globaltimer stamps of every 16th tile, profiles/r02_sweep_trace.txt): with one thread per slot
// (k_tile_sweep) a tile of the 216^3 mesh lives 11.7 us, 5.6 us of them in the walk of its 22 local levels -- a local
// level of an 8x8x8 tile holds <= 48 rows, yet all 16 warps execute the (predicated) level body, so the walk is bound by
// instruction issue; all CTA slots are busy all the time, i.e. the sweep (0.83 ms) is bound by tile THROUGHPUT, not by
// the tile-to-tile hand-over.  A first narrow version (64 threads do everything) was slower still: 8 slots of staging
// per thread and ~100 dependent instructions per level with 2.5 warps per scheduler is pure instruction latency.  Hence:
//   * ST threads stage the tile (NS / ST slots each, all loads of a pass in flight together) in ASCENDING ROW ORDER
//     (fc_tile_dir::meta_rm: matrix entries, d and input of consecutive rows are neighbours in memory) into shared
//     memory, scattered to the row's slot; slots are ordered by local level (fc_tile_schedule.hpp);
//   * after the producers' flags (or the tile-level counter) are seen, every value of another tile is loaded in one
//     batch and FOLDED into the staged coefficient: the entry becomes (product, dep = "one"), where slot NS of the value
//     array holds 1.0, so that the walk needs no case distinction: v - c * z[dep] with c * 1.0 = c exactly.  Entries
//     past the end of a short row are (0, "one"): v - 0 * 1 = v exactly.  (DIC_PAR / DILU carry a second factor c2:
//     v - (c * z[dep]) * c2, again 1.0 for folded entries.)  The sums are the same left-to-right sums, bit-identical;
//   * then all warps but the first WT / 32 retire and the walk runs branch-free on WT threads: level l = slots
//     [start[l], start[l+1]), one pass of 64 threads on hexahedra, ~25 instructions per row, one CTA barrier per level.
// Rows with more than PRE entries in their triangle (none on hexahedra with PRE = 4 or BCC polyhedra with PRE = 8) take
// a slow path through global memory after the staged entries.
// Shared layout (NS = FC_TILE slots): c[PRE][NS] | c2[PRE][NS] (DIC_PAR, DILU) | z[NS + 1] | di[NS] (FWD/BWD) |
// dep[PRE][NS] (u16) | row[NS] | s[NS] | e[NS] | lev[NS] (short) | start[NS + 2] (short).  Entry q of all slots is
// contiguous: the threads of a level read consecutive 8-byte words (a [slot][q] layout put them 8 PRE bytes apart:
// 16-way bank conflicts with PRE = 8, measured 670 ns per level on the polyhedral mesh against 163 ns with PRE = 3)
template <int MODE, int PRE>
struct fct_walk_layout {
  static constexpr int NS = FC_TILE;
  static constexpr bool TWO = MODE == TRI_DILU || MODE == TRI_DIC_PAR;
  static constexpr bool SOLVE = MODE == TRI_FWD || MODE == TRI_BWD;
  static constexpr size_t off_c = 0;
  static constexpr size_t off_c2 = off_c + sizeof(double) * PRE * NS;
  static constexpr size_t off_z = off_c2 + (TWO ? sizeof(double) * PRE * NS : 0);
  static constexpr size_t off_di = off_z + sizeof(double) * (NS + 2);
  static constexpr size_t off_dep = off_di + (SOLVE ? sizeof(double) * NS : 0);
  static constexpr size_t off_row = off_dep + sizeof(unsigned short) * PRE * NS;
  static constexpr size_t off_s = off_row + sizeof(int) * NS;
  static constexpr size_t off_e = off_s + sizeof(int) * NS;
  static constexpr size_t off_lev = off_e + sizeof(int) * NS;
  static constexpr size_t off_start = off_lev + sizeof(short) * NS;
  static constexpr size_t bytes = off_start + sizeof(short) * (NS + 2);
};

#ifndef FCT_TRACE   // measurement aid (-DFC_SWEEP_TRACE builds define it): time stamp number i of tile b
#define FCT_TRACE(i)
#endif
#ifndef FCT_TRACE_AT   // the same, written by thread t
#define FCT_TRACE_AT(t, i)
#endif
struct fct_handover {   // FLAGS: the producers' flags (fc_tile_dir::prod); otherwise the tile-level counters
  const int *blk_level, *lev_blocks_before, *prod, *prod_cnt;
  unsigned int *done, *ready, *flag;
  unsigned int sweep_no;
};

template <int MODE, int PRE, int ST, int WT, int OCC, bool FLAGS>
FCT_WALK_KERNEL(ST, OCC)
k_tile_walk(const int4 *__restrict__ meta_rm, const int *__restrict__ blk_nlev, unsigned int *ticket,
            unsigned int ticket_base, const int *__restrict__ tja, const int *__restrict__ diag,
            const int *__restrict__ tpos, const double *__restrict__ a, const double *__restrict__ d,
            const double *__restrict__ in, double *out, double small, double padd, const fc_scalars *sc,
            fct_handover H) {
  using L = fct_walk_layout<MODE, PRE>;
  constexpr int NS = FC_TILE;
  constexpr int ONE = NS;   // z[ONE] = 1.0
  FCT_DYN_SMEM(fct_raw);
  FCT_SHARED unsigned int s_b;
  FCT_SHARED int s_long;
  double *const s_c = reinterpret_cast<double *>(fct_raw + L::off_c);
  double *const s_c2 = reinterpret_cast<double *>(fct_raw + L::off_c2);
  double *const s_z = reinterpret_cast<double *>(fct_raw + L::off_z);
  double *const s_di = reinterpret_cast<double *>(fct_raw + L::off_di);
  unsigned short *const s_dep = reinterpret_cast<unsigned short *>(fct_raw + L::off_dep);
  int *const s_row = reinterpret_cast<int *>(fct_raw + L::off_row);
  int *const s_s = reinterpret_cast<int *>(fct_raw + L::off_s);
  int *const s_e = reinterpret_cast<int *>(fct_raw + L::off_e);
  short *const s_lev = reinterpret_cast<short *>(fct_raw + L::off_lev);
  short *const s_start = reinterpret_cast<short *>(fct_raw + L::off_start);
  if (sc && sc->done) return;
  const int tid = (int)FCT_TID;
  if (tid == 0) { s_b = FCT_TICKET(ticket) - ticket_base; s_long = 0; s_z[ONE] = 1.0; }
  FCT_SYNC();
  const unsigned int b = s_b;
  FCT_TRACE(0);
  const int nl = blk_nlev[b];
  // ---- pass 1: descriptors, in ascending row order; padding entries sit behind the rows in both orders
  //      (entry i >= rows <-> slot i >= rows) and sort behind the last level ----
  constexpr int SPT = NS / ST;
  int rw[SPT], ss[SPT], ee[SPT], sl[SPT];
FCT_UNROLL
  for (int u = 0; u < SPT; ++u) {
    const int4 mt = meta_rm[(size_t)b * NS + tid + u * ST];   // row, slot | level << 16, triangle [s, e)
    const int slot = mt.x >= 0 ? (mt.y & 0xffff) : tid + u * ST;
    rw[u] = mt.x; ss[u] = mt.z; ee[u] = mt.w; sl[u] = slot;
    s_row[slot] = mt.x;
    s_lev[slot] = (short)(mt.x >= 0 ? (mt.y >> 16) : nl);
    s_s[slot] = mt.z;
    s_e[slot] = mt.w;
    if (mt.x >= 0 && mt.w - mt.z > PRE) s_long = 1;
  }
  FCT_TRACE(1);
  // ---- pass 2: start value, d, the first PRE coefficients and columns: every load before the first store (no
  //      branches: padding, or a position past the end of the row, loads position 0 / row 0) ----
  double av[SPT][PRE], tv[SPT][PRE], vv[SPT], dv[SPT];
  int jv[SPT][PRE];
FCT_UNROLL
  for (int u = 0; u < SPT; ++u) {
    const int row = rw[u] >= 0 ? rw[u] : 0;
    if (MODE == TRI_FWD || MODE == TRI_BWD) { vv[u] = in[row]; dv[u] = d[row]; }
    else { vv[u] = a[diag[row]]; dv[u] = 0.0; }
FCT_UNROLL
    for (int q = 0; q < PRE; ++q) {
      const bool live = rw[u] >= 0 && ss[u] + q < ee[u];
      const int k = live ? ss[u] + q : 0;
      av[u][q] = a[k];
      jv[u][q] = live ? tja[k] : -(ONE + 1);   // a dead entry depends on "one" with coefficient 0
      if (MODE == TRI_DILU) tv[u][q] = a[tpos[k]];
      if (!live) av[u][q] = 0.0;
    }
  }
  // everything that does not depend on other tiles goes to shared memory now, before the wait: start value, d,
  // in-tile and dead entries
FCT_UNROLL
  for (int u = 0; u < SPT; ++u) {
    const int slot = sl[u];
    if (rw[u] < 0) continue;
    s_z[slot] = MODE == TRI_BWD ? vv[u] / (dv[u] + small) : vv[u];   // z = z/(d+small), iccg.f90:102
    if (L::SOLVE) s_di[slot] = dv[u];
FCT_UNROLL
    for (int q = 0; q < PRE; ++q) {
      if (jv[u][q] < 0) {
        double c = av[u][q], c2 = 1.0;
        const int dep = -jv[u][q] - 1;   // slot of the same tile, or ONE for a dead entry (c = 0)
        if (MODE == TRI_DIC) c = c * c;
        else if (MODE == TRI_DIC_PAR) c2 = dep == ONE ? 1.0 : c;
        else if (MODE == TRI_DILU) c2 = dep == ONE ? 1.0 : tv[u][q];
        s_c[q * NS + slot] = c;
        if (L::TWO) s_c2[q * NS + slot] = c2;
        s_dep[q * NS + slot] = (unsigned short)dep;
      }
    }
  }
  FCT_SYNC();
  // first slot of every local level (levels 0 .. nl-1 are all non-empty; `nl` = the padding)
  for (int slot = tid; slot < NS; slot += ST) {
    const int lv = s_lev[slot];
    if (slot == 0 || s_lev[slot - 1] != lv) s_start[lv] = (short)slot;
    if (slot == NS - 1 && lv != nl) s_start[nl] = (short)NS;   // a full tile has no padding slot
  }
  FCT_TRACE(2);
  // ---- the rows of other tiles exist: producers' flags, or the counter of the previous tile level ----
  if (FLAGS) {
    const int np = H.prod_cnt[b];
    if (tid < np) {
      const unsigned int *r = H.flag + H.prod[b * FC_TILE_MAXP + tid];
      fc_spin_guard g;
      while (ld_acquire(r) < H.sweep_no) g.tick();
    }
  } else {
    const int lev = H.blk_level[b];
    if (lev > 0 && tid == 0) {
      const unsigned int *r = H.ready + (lev - 1);
      fc_spin_guard g;
      while (ld_acquire(r) < H.sweep_no) g.tick();
    }
  }
  FCT_SYNC();
  FCT_TRACE(3);
  // ---- pass 3: values of other tiles, one batch, folded into the coefficient: the walk subtracts c * 1.0 (* 1.0) ----
  double zv[SPT][PRE];
FCT_UNROLL
  for (int u = 0; u < SPT; ++u)
FCT_UNROLL
    for (int q = 0; q < PRE; ++q) zv[u][q] = jv[u][q] >= 0 ? FCT_LDCG(out + jv[u][q]) : 1.0;
FCT_UNROLL
  for (int u = 0; u < SPT; ++u) {
    const int slot = sl[u];
FCT_UNROLL
    for (int q = 0; q < PRE; ++q) {
      if (jv[u][q] >= 0) {
        double c = av[u][q];
        const double zj = zv[u][q];
        if (MODE == TRI_FWD || MODE == TRI_BWD) c = c * zj;
        else if (MODE == TRI_DIC) c = (c * c) * zj;            // iccg.f90:80
        else if (MODE == TRI_DIC_PAR) c = c * zj * c;          // src-parallel/iccg.f90:97
        else c = c * zj * tv[u][q];                            // bicgstab.f90:76
        s_c[q * NS + slot] = c;
        if (L::TWO) s_c2[q * NS + slot] = 1.0;
        s_dep[q * NS + slot] = (unsigned short)ONE;
      }
    }
  }
  FCT_SYNC();
  FCT_TRACE(4);
  FCT_TRACE(5);
  if (tid >= WT) return;   // the staging warps retire; barriers below count the remaining threads only
  // ---- walk the local levels: branch-free, shared memory only ----
  // (fetching the next level's coefficients before the barrier was measured and lost: 5.4 us against 3.5 us per tile)
  const bool any_long = s_long != 0;
  for (int l = 0; l < nl; ++l) {
    const int s0 = s_start[l], s1 = s_start[l + 1];
    for (int slot = s0 + tid; slot < s1; slot += WT) {
      // every shared-memory load of the row first (coefficients and dependency slots, then the values they name),
      // then the products, then the left-to-right sum: only the PRE subtractions form a dependent chain.  (Written
      // as one loop the PRE = 8 instantiation kept load -> load -> multiply -> subtract per entry in sequence:
      // 670 ns per local level on the polyhedral mesh, profiles/r02_sweep_trace_summary.txt.)
      double cq[PRE], zq[PRE], c2q[PRE];
      int dq[PRE];
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) {
        cq[q] = s_c[q * NS + slot];
        dq[q] = s_dep[q * NS + slot];
        if (L::TWO) c2q[q] = s_c2[q * NS + slot];
      }
      double v = s_z[slot];
      const int row = s_row[slot];
      __asm__ __volatile__("" ::: "memory");   // compiler-only fence: keep the first batch of loads together ...
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) zq[q] = s_z[dq[q]];
      __asm__ __volatile__("" ::: "memory");   // ... and the dependent batch ahead of the arithmetic
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) {
        cq[q] = cq[q] * zq[q];
        if (L::TWO) cq[q] = cq[q] * c2q[q];
      }
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) v = v - cq[q];
      if (any_long) {                                              // rows longer than PRE (none on hex / BCC meshes)
        const int e = s_e[slot];
        for (int k = s_s[slot] + PRE; k < e; ++k) {
          const int j = tja[k];
          const double zj = j < 0 ? s_z[-j - 1] : FCT_LDCG(out + j);
          const double ak = a[k];
          if (MODE == TRI_FWD || MODE == TRI_BWD) v = v - ak * zj;
          else if (MODE == TRI_DIC) v = v - (ak * ak) * zj;
          else if (MODE == TRI_DIC_PAR) v = v - ak * zj * ak;
          else v = v - ak * zj * a[tpos[k]];
        }
      }
      const double r = L::SOLVE ? v * s_di[slot] : 1.0 / (v + padd);
      s_z[slot] = r;
      out[row] = r;
    }
    FCT_WALK_SYNC(WT);   // the last one also orders every row's store before thread 0's release below
  }
  FCT_TRACE(6);
  if (tid == 0) {
    if (FLAGS) {
      st_release(H.flag + b, H.sweep_no);
    } else {
      const int lev = H.blk_level[b];
      const unsigned int nb = (unsigned int)(H.lev_blocks_before[lev + 1] - H.lev_blocks_before[lev]);
      const unsigned int old = atom_add_acq_rel(H.done + lev, 1u);
      if (old + 1u == H.sweep_no * nb) st_release(H.ready + lev, H.sweep_no);
    }
  }
  FCT_TRACE(7);
}


// ---------------------------------------------------------------------------------------------------------------
// k_tile_walk_vf (FC_TUNE_SWEEP_TILED = 5): the staged walk of k_tile_walk with the value-as-flag hand-over of
// k_tile_sweep_vf, the polls taken off the walk's critical path.
//
// With flags a tile starts when its producers have FINISHED: the tile-to-tile critical path of the 216^3 sweep is
// 79 tile levels x (flag visible 1.8 + fold 1.3 + walk 3.6 + release 0.8 us) = 0.59 ms, three times the time its bytes
// need.  But the first rows of a tile only need the rows on the near face of its producers, which those produce a
// third of the way through their own walk.  Here `out` holds the "unset" pattern before the sweep (as in
// k_tile_sweep_vf; the sweeps re-arm each other's vectors) and the CTA has two kinds of threads after the staging:
//   * ST staging threads turn into HELPERS: each keeps the (at most SPT x PRE) entries of its slots that name rows of
//     other tiles in registers, polls their values round-robin -- every round one batch of L2 reads for everything
//     still missing -- and, when a value has arrived, folds it into the staged coefficient exactly as pass 3 of
//     k_tile_walk does and counts the slot's local level down in shared memory;
//   * WT WALKERS (their own two warps: they stage nothing) walk the local levels as before and, before level l, wait
//     until the level's count of unfolded entries is zero.  They publish every row with a relaxed store the moment it
//     is computed; no flag, no fence.
// In the steady state a tile trails its producers by the levels of the face it needs plus one hand-over latency
// instead of by their whole walk, and the helpers run ahead of the walkers, so no poll sits on a level's critical
// path.  Tickets are drawn in tile order and a tile only reads rows of tiles with smaller tickets, so every wait is
// for a CTA that is already running.  Same operands, same left-to-right sums: bit-identical to every other schedule.
template <int MODE, int PRE, int ST, int WT, int OCC>
FCT_WALK_KERNEL(ST + WT, OCC)
k_tile_walk_vf(const int4 *__restrict__ meta_rm, const int *__restrict__ blk_nlev, unsigned int *ticket,
               unsigned int ticket_base, const int *__restrict__ tja, const int *__restrict__ diag,
               const int *__restrict__ tpos, const double *__restrict__ a, const double *__restrict__ d,
               double *in_rw, double *out, double *arm, double small, double padd, const fc_scalars *sc,
               bool rearm_in, unsigned int backoff_ns) {
  using L = fct_walk_layout<MODE, PRE>;
  constexpr int NS = FC_TILE;
  constexpr int ONE = NS;   // z[ONE] = 1.0
  constexpr int TT = ST + WT;
  FCT_DYN_SMEM(fct_raw);
  FCT_SHARED unsigned int s_b;
  FCT_SHARED int s_long;
  double *const s_c = reinterpret_cast<double *>(fct_raw + L::off_c);
  double *const s_c2 = reinterpret_cast<double *>(fct_raw + L::off_c2);
  double *const s_z = reinterpret_cast<double *>(fct_raw + L::off_z);
  double *const s_di = reinterpret_cast<double *>(fct_raw + L::off_di);
  unsigned short *const s_dep = reinterpret_cast<unsigned short *>(fct_raw + L::off_dep);
  int *const s_row = reinterpret_cast<int *>(fct_raw + L::off_row);
  int *const s_s = reinterpret_cast<int *>(fct_raw + L::off_s);
  int *const s_e = reinterpret_cast<int *>(fct_raw + L::off_e);
  short *const s_lev = reinterpret_cast<short *>(fct_raw + L::off_lev);
  short *const s_start = reinterpret_cast<short *>(fct_raw + L::off_start);
  int *const s_pend = reinterpret_cast<int *>(fct_raw + ((L::bytes + 15) & ~(size_t)15));   // [NS + 2] per local level
  if (sc && sc->done) return;
  const int tid = (int)FCT_TID;
  if (tid == 0) { s_b = FCT_TICKET(ticket) - ticket_base; s_long = 0; s_z[ONE] = 1.0; }
  for (int i = tid; i < NS + 2; i += TT) s_pend[i] = 0;
  FCT_SYNC();
  const unsigned int b = s_b;
  FCT_TRACE(0);
  const int nl = blk_nlev[b];
  const bool walker = tid < WT;
  const int st = tid - WT;   // staging / helper thread number
  constexpr int SPT = NS / ST;
  int rw[SPT], ss[SPT], ee[SPT], sl[SPT], lv[SPT];
  double av[SPT][PRE], tv[SPT][PRE], vv[SPT], dv[SPT];
  int jv[SPT][PRE];
  if (!walker) {
    // ---- pass 1: descriptors, in ascending row order (see k_tile_walk) ----
FCT_UNROLL
    for (int u = 0; u < SPT; ++u) {
      const int4 mt = meta_rm[(size_t)b * NS + st + u * ST];   // row, slot | level << 16, triangle [s, e)
      const int slot = mt.x >= 0 ? (mt.y & 0xffff) : st + u * ST;
      rw[u] = mt.x; ss[u] = mt.z; ee[u] = mt.w; sl[u] = slot;
      lv[u] = mt.x >= 0 ? (mt.y >> 16) : nl;
      s_row[slot] = mt.x;
      s_lev[slot] = (short)lv[u];
      s_s[slot] = mt.z;
      s_e[slot] = mt.w;
      if (mt.x >= 0 && mt.w - mt.z > PRE) s_long = 1;
    }
    // ---- pass 2: start value, d, the first PRE coefficients and columns ----
FCT_UNROLL
    for (int u = 0; u < SPT; ++u) {
      const int row = rw[u] >= 0 ? rw[u] : 0;
      if (MODE == TRI_FWD || MODE == TRI_BWD) { vv[u] = in_rw[row]; dv[u] = d[row]; }
      else { vv[u] = a[diag[row]]; dv[u] = 0.0; }
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) {
        const bool live = rw[u] >= 0 && ss[u] + q < ee[u];
        const int k = live ? ss[u] + q : 0;
        av[u][q] = a[k];
        jv[u][q] = live ? tja[k] : -(ONE + 1);   // a dead entry depends on "one" with coefficient 0
        if (MODE == TRI_DILU) tv[u][q] = a[tpos[k]];
        if (!live) av[u][q] = 0.0;
      }
    }
FCT_UNROLL
    for (int u = 0; u < SPT; ++u) {
      const int slot = sl[u];
      if (rw[u] < 0) continue;
      // the vectors the next sweeps hand over through: unset again (see k_tile_sweep_vf)
      if (MODE == TRI_FWD && arm) arm[rw[u]] = fct_unset();
      if (MODE == TRI_BWD && rearm_in) in_rw[rw[u]] = fct_unset();
      s_z[slot] = MODE == TRI_BWD ? vv[u] / (dv[u] + small) : vv[u];   // z = z/(d+small), iccg.f90:102
      if (L::SOLVE) s_di[slot] = dv[u];
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) {
        if (jv[u][q] < 0) {
          double c = av[u][q], c2 = 1.0;
          const int dep = -jv[u][q] - 1;   // slot of the same tile, or ONE for a dead entry (c = 0)
          if (MODE == TRI_DIC) c = c * c;
          else if (MODE == TRI_DIC_PAR) c2 = dep == ONE ? 1.0 : c;
          else if (MODE == TRI_DILU) c2 = dep == ONE ? 1.0 : tv[u][q];
          s_c[q * NS + slot] = c;
          if (L::TWO) s_c2[q * NS + slot] = c2;
          s_dep[q * NS + slot] = (unsigned short)dep;
        } else {
          FCT_SATOMIC_ADD(s_pend + lv[u], 1);   // a row of another tile: the level waits for it
        }
      }
    }
  }
  FCT_SYNC();
  FCT_TRACE(1);
  for (int slot = tid; slot < NS; slot += TT) {   // first slot of every local level
    const int lvl = s_lev[slot];
    if (slot == 0 || s_lev[slot - 1] != lvl) s_start[lvl] = (short)slot;
    if (slot == NS - 1 && lvl != nl) s_start[nl] = (short)NS;   // a full tile has no padding slot
  }
  FCT_SYNC();
  if (!walker) {
    // ---- helpers: poll what other tiles produce, fold it in, count the level down ----
    unsigned int pend = 0u;
    double zv[SPT][PRE];
FCT_UNROLL
    for (int u = 0; u < SPT; ++u)
FCT_UNROLL
      for (int q = 0; q < PRE; ++q)
        if (rw[u] >= 0 && jv[u][q] >= 0) {
          pend |= 1u << (u * PRE + q);
          zv[u][q] = FCT_LD_POLL(out + jv[u][q]);
        }
    fc_spin_guard g;
    while (pend) {
      bool progress = false;
FCT_UNROLL
      for (int u = 0; u < SPT; ++u) {
FCT_UNROLL
        for (int q = 0; q < PRE; ++q) {
          if (pend & (1u << (u * PRE + q))) {
            const double zj = zv[u][q];
            if (!fct_is_unset(zj)) {
              const int slot = sl[u];
              double c = av[u][q];
              if (MODE == TRI_FWD || MODE == TRI_BWD) c = c * zj;
              else if (MODE == TRI_DIC) c = (c * c) * zj;            // iccg.f90:80
              else if (MODE == TRI_DIC_PAR) c = c * zj * c;          // src-parallel/iccg.f90:97
              else c = c * zj * tv[u][q];                            // bicgstab.f90:76
              s_c[q * NS + slot] = c;
              if (L::TWO) s_c2[q * NS + slot] = 1.0;
              s_dep[q * NS + slot] = (unsigned short)ONE;
              FCT_FENCE_BLOCK();   // the folded entry before the count
              FCT_SATOMIC_ADD(s_pend + lv[u], -1);
              pend &= ~(1u << (u * PRE + q));
              progress = true;
            } else {
              zv[u][q] = FCT_LD_POLL(out + jv[u][q]);
            }
          }
        }
      }
      // a helper that found nothing sleeps: 18 spinning helper warps per SM would take the issue slots of the 6 walking ones
      if (!progress) { g.tick(); FCT_BACKOFF(backoff_ns); }
    }
    FCT_TRACE_AT(WT, 6);
    return;
  }
  FCT_TRACE(2);
  // ---- walkers: the local levels, branch-free, shared memory only; a level starts when nothing of it is unfolded ----
  const bool any_long = s_long != 0;
  for (int l = 0; l < nl; ++l) {
    {
      fc_spin_guard g;
      while (FCT_LD_SVOL(s_pend + l) != 0) g.tick();
    }
    FCT_FENCE_BLOCK();
    if (l == 0) FCT_TRACE(3);
    if (l == nl / 2) FCT_TRACE(4);
    const int s0 = s_start[l], s1 = s_start[l + 1];
    for (int slot = s0 + tid; slot < s1; slot += WT) {
      double cq[PRE], zq[PRE], c2q[PRE];
      int dq[PRE];
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) {
        cq[q] = s_c[q * NS + slot];
        dq[q] = s_dep[q * NS + slot];
        if (L::TWO) c2q[q] = s_c2[q * NS + slot];
      }
      double v = s_z[slot];
      const int row = s_row[slot];
      __asm__ __volatile__("" ::: "memory");
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) zq[q] = s_z[dq[q]];
      __asm__ __volatile__("" ::: "memory");
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) {
        cq[q] = cq[q] * zq[q];
        if (L::TWO) cq[q] = cq[q] * c2q[q];
      }
FCT_UNROLL
      for (int q = 0; q < PRE; ++q) v = v - cq[q];
      if (any_long) {                                              // rows longer than PRE (none on hex / BCC meshes)
        const int e = s_e[slot];
        for (int k = s_s[slot] + PRE; k < e; ++k) {
          const int j = tja[k];
          double zj;
          if (j < 0) {
            zj = s_z[-j - 1];
          } else {
            zj = FCT_LD_POLL(out + j);
            fc_spin_guard g;
            while (fct_is_unset(zj)) { g.tick(); zj = FCT_LD_POLL(out + j); }
          }
          const double ak = a[k];
          if (MODE == TRI_FWD || MODE == TRI_BWD) v = v - ak * zj;
          else if (MODE == TRI_DIC) v = v - (ak * ak) * zj;
          else if (MODE == TRI_DIC_PAR) v = v - ak * zj * ak;
          else v = v - ak * zj * a[tpos[k]];
        }
      }
      const double r = L::SOLVE ? v * s_di[slot] : 1.0 / (v + padd);
      s_z[slot] = r;
      FCT_ST_PUB(out + row, r);
    }
    FCT_WALK_SYNC(WT);
  }
  FCT_TRACE(5);
  FCT_TRACE(7);
}
