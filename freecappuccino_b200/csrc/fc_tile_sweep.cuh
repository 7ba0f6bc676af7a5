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
