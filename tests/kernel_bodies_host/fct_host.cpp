// TEST INFRASTRUCTURE ONLY -- not part of the product, never loaded by freecappuccino_b200.
//
// Compiles the tile-schedule builder of the triangular sweeps (freecappuccino_b200/csrc/fc_tile_schedule.hpp, the
// same header fc_trisolve.cu includes) with g++ and walks the schedule exactly as k_tile_sweep does -- tiles in
// ticket order, local levels in order, rows of a local level in any order, in-tile values through the tile's
// shared array, the others through the output vector -- so that tests/test_tile_schedule.py can check on a machine
// without a GPU that (a) no row is visited before the rows it depends on, (b) a value read through global memory
// was produced by a tile of a strictly lower tile level (the only ordering the kernel's level counters give) and
// (c) the result is bit-identical to the natural-order sweep of iccg.f90:94-111 / bicgstab.f90:68-79.
//
// fct_emu_sweep goes one step further: it compiles the kernel source itself (fc_tile_sweep.cuh) for the host -- one
// std::thread per CUDA thread of a CTA, std::barrier for __syncthreads, GCC atomics for the acquire / release words --
// and runs the CTAs one after the other in ticket order, twice in a row on the same counters (sweep numbers 1, 2).
// That checks the kernel's own index arithmetic, register prefetch and barrier structure, not a restatement of it.
#include <barrier>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <thread>

#include "../../freecappuccino_b200/csrc/fc_tile_schedule.hpp"

enum { TRI_FWD = 0, TRI_BWD = 1, TRI_DIC = 2, TRI_DIC_PAR = 3, TRI_DILU = 4 };

// ---- host stand-ins for what fc_tile_sweep.cuh expects from its includer ----
struct fc_scalars { int done; };
struct int4 { int x, y, z, w; };
namespace {
thread_local unsigned fct_tid = 0;
std::barrier<> *fct_bar = nullptr;
}
#define FCT_KERNEL(OCC) static void
#define FCT_SHARED static
#define FCT_TID fct_tid
#define FCT_SYNC() fct_bar->arrive_and_wait()
#define FCT_TICKET(p) __atomic_fetch_add((p), 1u, __ATOMIC_RELAXED)
#define FCT_LDCG(p) (*(const volatile double *)(p))
#define FCT_UNROLL
// value-as-flag mode: relaxed 8-byte atomics stand in for ld.relaxed.gpu / st.relaxed.gpu
#include <cstring>
static inline double fct_ld_poll(const double *p) {
  const long long x = __atomic_load_n((const long long *)p, __ATOMIC_RELAXED);
  double v;
  std::memcpy(&v, &x, 8);
  return v;
}
static inline void fct_st_pub(double *p, double v) {
  long long x;
  std::memcpy(&x, &v, 8);
  __atomic_store_n((long long *)p, x, __ATOMIC_RELAXED);
}
static inline bool fct_is_unset(double v) { long long x; std::memcpy(&x, &v, 8); return x == -1ll; }
static inline double fct_unset() { const long long x = -1ll; double v; std::memcpy(&v, &x, 8); return v; }
#define FCT_LD_POLL(p) fct_ld_poll(p)
#define FCT_ST_PUB(p, v) fct_st_pub((p), (v))
#define FCT_WALK_KERNEL(T, OCC) static void
namespace { std::barrier<> *fct_walk_bar = nullptr; }
#define FCT_WALK_SYNC(N) fct_walk_bar->arrive_and_wait()
#define FCT_DYN_SMEM(name) alignas(16) static unsigned char name[160 * 1024]
#define FCT_SATOMIC_ADD(p, v) __atomic_fetch_add((p), (v), __ATOMIC_ACQ_REL)
#define FCT_LD_SVOL(p) __atomic_load_n((p), __ATOMIC_ACQUIRE)
#define FCT_FENCE_BLOCK() __atomic_thread_fence(__ATOMIC_SEQ_CST)
#define FCT_BACKOFF(ns) std::this_thread::yield()
static inline unsigned int ld_acquire(const unsigned int *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void st_release(unsigned int *p, unsigned int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline unsigned int atom_add_acq_rel(unsigned int *p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
struct fc_spin_guard {   // CTAs run in ticket order here, so a wait that does not end at once is a schedule bug
  unsigned long long n = 0;
  void tick() { if (++n > (1ull << 24)) std::abort(); std::this_thread::yield(); }
};
#include "../../freecappuccino_b200/csrc/fc_tile_sweep.cuh"

static inline double step(int mode, double v, double ak, double zj, double atk) {
  if (mode == TRI_FWD || mode == TRI_BWD) return v - ak * zj;
  if (mode == TRI_DIC) return v - (ak * ak) * zj;
  if (mode == TRI_DIC_PAR) return v - ak * zj * ak;
  return v - ak * zj * atk;
}
static inline double start(int mode, int row, const int *diag, const double *a, const double *d, const double *in,
                           double small, double &di) {
  di = 0.0;
  if (mode == TRI_FWD) { di = d[row]; return in[row]; }
  if (mode == TRI_BWD) { di = d[row]; return in[row] / (di + small); }
  return a[diag[row]];
}
static inline double finish(int mode, double v, double di, double padd) {
  return (mode == TRI_FWD || mode == TRI_BWD) ? v * di : 1.0 / (v + padd);
}

extern "C" {

void *fct_build(int n, const int *ioffset, const int *ja, const int *diag, const double *xc, const double *yc,
                const double *zc) {
  return new fc_tile_schedule(fc_build_tile_schedule(n, ioffset, ja, diag, xc, yc, zc));
}
// the same with narrower starting bins (the library starts polyhedral meshes at min_shrink = 2: 6 cells per axis)
void *fct_build_shrunk(int n, const int *ioffset, const int *ja, const int *diag, const double *xc, const double *yc,
                       const double *zc, int min_shrink) {
  return new fc_tile_schedule(fc_build_tile_schedule(n, ioffset, ja, diag, xc, yc, zc, min_shrink));
}
void fct_free(void *h) { delete (fc_tile_schedule *)h; }
int fct_ok(void *h) { return ((fc_tile_schedule *)h)->ok ? 1 : 0; }
const char *fct_why(void *h) { return ((fc_tile_schedule *)h)->why.c_str(); }
void fct_info(void *h, int *out) {
  const fc_tile_schedule &S = *(fc_tile_schedule *)h;
  out[0] = S.ntiles; out[1] = S.cells_per_axis; out[2] = S.max_tile_rows;
  out[3] = S.lower.nlev; out[4] = S.lower.max_local_levels;
  out[5] = S.upper.nlev; out[6] = S.upper.max_local_levels;
  out[7] = S.lower.p2p_ok && S.upper.p2p_ok ? 1 : 0;
  out[8] = std::max(S.lower.max_producers, S.upper.max_producers);
  out[9] = S.repaired_rows;
  out[10] = (int)std::min<long long>(S.cost, 2000000000LL);
}

// the sweep in the reference's own order: rows ascending (lower triangle) or descending (upper)
void fct_reference_sweep(int mode, int n, const int *ioffset, const int *ja, const int *diag, const int *tpos,
                         const double *a, const double *d, const double *in, double *out, double small, double padd) {
  const bool bwd = mode == TRI_BWD;
  for (int q = 0; q < n; ++q) {
    const int row = bwd ? n - 1 - q : q;
    const int s = bwd ? diag[row] + 1 : ioffset[row], e = bwd ? ioffset[row + 1] : diag[row];
    double di, v = start(mode, row, diag, a, d, in, small, di);
    for (int k = s; k < e; ++k) v = step(mode, v, a[k], out[ja[k]], mode == TRI_DILU ? a[tpos[k]] : 0.0);
    out[row] = finish(mode, v, di, padd);
  }
}

// the sweep as k_tile_sweep walks it; `out` must arrive filled with NaN.  Returns the number of ordering violations.
int fct_sweep(void *h, int mode, int n, const int *ioffset, const int *diag, const int *tpos, const double *a,
              const double *d, const double *in, double *out, double small, double padd) {
  const fc_tile_schedule &S = *(fc_tile_schedule *)h;
  const fc_tile_dir &D = mode == TRI_BWD ? S.upper : S.lower;
  const bool bwd = mode == TRI_BWD;
  std::vector<int> produced_level(n, -1), produced_block(n, -1);
  std::vector<double> s_z(FC_TILE);
  int bad = 0, visited = 0;
  for (int b = 0; b < D.nblocks; ++b) {
    const int lev = D.blk_level[b];
    if (b < D.lev_blocks_before[lev] || b >= D.lev_blocks_before[lev + 1]) ++bad;   // ticket order is level-major
    std::fill(s_z.begin(), s_z.end(), std::numeric_limits<double>::quiet_NaN());
    for (int l = 0; l < D.blk_nlev[b]; ++l)
      for (int t = FC_TILE - 1; t >= 0; --t) {   // any order inside a local level: take the reverse one
        const size_t slot = (size_t)b * FC_TILE + t;
        const int row = D.rows[slot];
        if (row < 0 || D.llev[slot] != l) continue;
        const int s = bwd ? diag[row] + 1 : ioffset[row], e = bwd ? ioffset[row + 1] : diag[row];
        double di, v = start(mode, row, diag, a, d, in, small, di);
        for (int k = s; k < e; ++k) {
          const int j = S.tja[k];
          double zj;
          if (j < 0) zj = s_z[-j - 1];
          else {
            zj = out[j];
            if (produced_level[j] < 0 || produced_level[j] >= lev) ++bad;
            if (D.p2p_ok) {   // point-to-point hand-over: the producing tile must be one the tile waits for
              bool named = false;
              for (int p = 0; p < D.prod_cnt[b]; ++p) named = named || D.prod[(size_t)b * FC_TILE_MAXP + p] == produced_block[j];
              if (!named || produced_block[j] >= b) ++bad;
            }
          }
          if (std::isnan(zj)) ++bad;
          v = step(mode, v, a[k], zj, mode == TRI_DILU ? a[tpos[k]] : 0.0);
        }
        const double r = finish(mode, v, di, padd);
        s_z[t] = r;
        out[row] = r;
        produced_level[row] = lev;
        produced_block[row] = b;
        ++visited;
      }
  }
  if (visited != n) ++bad;
  return bad;
}

// can this machine run FC_TILE threads at once?  (the emulation would dead-lock in its barrier otherwise)
int fct_can_emulate(void) {
  std::vector<std::thread> th;
  bool ok = true;
  try {
    for (int t = 0; t < FC_TILE; ++t) th.emplace_back([]() {});
  } catch (...) {
    ok = false;
  }
  for (auto &x : th) x.join();
  return ok ? 1 : 0;
}

// the kernel source itself, CTA by CTA; `nsweeps` launches in a row on the same counters.  Returns 0, or -1 for an
// unknown mode.  `out` is refilled with NaN before every launch.
int fct_emu_sweep(void *h, int mode, int pre8, int p2p, int nsweeps, int n, const int *ioffset, const int *diag,
                  const int *tpos, const double *a, const double *d, const double *in, double *out, double small,
                  double padd) {
  const fc_tile_schedule &S = *(fc_tile_schedule *)h;
  const fc_tile_dir &D = mode == TRI_BWD ? S.upper : S.lower;
  std::vector<unsigned int> done(D.nlev, 0), ready(D.nlev, 0), flag(D.nblocks, 0);
  unsigned int ticket = 0;
  using kernel_t = void (*)(const int4 *, const int *, const int *, const int *, unsigned int *, unsigned int *,
                            unsigned int *, const int *, const int *, unsigned int *, unsigned int, unsigned int,
                            const int *, const int *, const int *, const double *, const double *, const double *,
                            double *, double, double, const fc_scalars *);
  kernel_t k = nullptr;
#define FCT_PICK(M)                                                                                          \
  case M:                                                                                                    \
    k = pre8 ? (p2p ? k_tile_sweep<M, 8, true, 2> : k_tile_sweep<M, 8, false, 2>)                            \
             : (p2p ? k_tile_sweep<M, 4, true, 2> : k_tile_sweep<M, 4, false, 2>);                           \
    break;
  switch (mode) {
    FCT_PICK(TRI_FWD) FCT_PICK(TRI_BWD) FCT_PICK(TRI_DIC) FCT_PICK(TRI_DIC_PAR) FCT_PICK(TRI_DILU)
    default: return -1;
  }
#undef FCT_PICK
  std::barrier<> bar(FC_TILE);
  fct_bar = &bar;
  for (int sweep = 1; sweep <= nsweeps; ++sweep) {
    for (int i = 0; i < n; ++i) out[i] = std::numeric_limits<double>::quiet_NaN();
    const unsigned int base = (unsigned int)(sweep - 1) * (unsigned int)D.nblocks;
    std::vector<std::thread> th;
    th.reserve(FC_TILE);
    for (int t = 0; t < FC_TILE; ++t)
      th.emplace_back([&, t]() {
        fct_tid = (unsigned)t;
        for (int b = 0; b < D.nblocks; ++b) {   // one CTA after the other: the static "shared" arrays are reused
          k((const int4 *)D.meta.data(), D.blk_nlev.data(), D.blk_level.data(), D.lev_blocks_before.data(),
            done.data(), ready.data(), &ticket, D.prod.data(), D.prod_cnt.data(), flag.data(), base,
            (unsigned int)sweep, S.tja.data(), diag, tpos, a, d, in, out, small, padd, nullptr);
          bar.arrive_and_wait();
        }
      });
    for (auto &x : th) x.join();
  }
  fct_bar = nullptr;
  return 0;
}

// k_tile_sweep_vf (value-as-flag hand-over), CTA by CTA in ticket order.  `out` arrives in any state: it is filled
// with the "unset" pattern here, as the library's memset does; `in` is copied because the backward sweep re-arms it.
// `arm` (n doubles or null) must come back all-unset after a forward sweep, `in_copy` all-unset after a backward one:
// returns the number of entries for which that does not hold (0 = ok), -1 for an unknown mode.
int fct_emu_sweep_vf(void *h, int mode, int pre8, int n, const int *ioffset, const int *diag, const int *tpos,
                     const double *a, const double *d, const double *in, double *out, double small, double padd) {
  const fc_tile_schedule &S = *(fc_tile_schedule *)h;
  const fc_tile_dir &D = mode == TRI_BWD ? S.upper : S.lower;
  unsigned int ticket = 0;
  using kernel_t = void (*)(const int4 *, const int *, unsigned int *, unsigned int, const int *, const int *,
                            const int *, const double *, const double *, double *, double *, double *, double, double,
                            const fc_scalars *, bool);
  kernel_t k = nullptr;
#define FCT_PICK(M) case M: k = pre8 ? k_tile_sweep_vf<M, 8, 2> : k_tile_sweep_vf<M, 4, 2>; break;
  switch (mode) {
    FCT_PICK(TRI_FWD) FCT_PICK(TRI_BWD) FCT_PICK(TRI_DIC) FCT_PICK(TRI_DIC_PAR) FCT_PICK(TRI_DILU)
    default: return -1;
  }
#undef FCT_PICK
  std::vector<double> in_copy(in, in + n), arm(n, 0.0);
  for (int i = 0; i < n; ++i) out[i] = fct_unset();
  std::barrier<> bar(FC_TILE);
  fct_bar = &bar;
  std::vector<std::thread> th;
  th.reserve(FC_TILE);
  for (int t = 0; t < FC_TILE; ++t)
    th.emplace_back([&, t]() {
      fct_tid = (unsigned)t;
      for (int b = 0; b < D.nblocks; ++b) {
        k((const int4 *)D.meta.data(), D.blk_nlev.data(), &ticket, 0u, S.tja.data(), diag, tpos, a, d, in_copy.data(),
          out, mode == TRI_FWD ? arm.data() : nullptr, small, padd, nullptr, true);
        bar.arrive_and_wait();
      }
    });
  for (auto &x : th) x.join();
  fct_bar = nullptr;
  int bad = 0;
  if (mode == TRI_FWD) for (int i = 0; i < n; ++i) bad += fct_is_unset(arm[i]) ? 0 : 1;
  if (mode == TRI_BWD) for (int i = 0; i < n; ++i) bad += fct_is_unset(in_copy[i]) ? 0 : 1;
  for (int i = 0; i < n; ++i) bad += fct_is_unset(out[i]) ? 1 : 0;
  return bad;
}

// k_tile_walk (256 threads stage a tile in shared memory, 64 walk it; flags or tile-level counters), CTA by CTA in
// ticket order, two launches in a row on the same flags / counters.  `flags` = 1: producer flags, 0: level counters.
int fct_emu_walk(void *h, int mode, int pre8, int flags, int n, const int *ioffset, const int *diag, const int *tpos,
                 const double *a, const double *d, const double *in, double *out, double small, double padd) {
  const fc_tile_schedule &S = *(fc_tile_schedule *)h;
  const fc_tile_dir &D = mode == TRI_BWD ? S.upper : S.lower;
  unsigned int ticket = 0;
  constexpr int ST = 256, WT = 64;
  using kernel_t = void (*)(const int4 *, const int *, unsigned int *, unsigned int, const int *, const int *,
                            const int *, const double *, const double *, const double *, double *, double, double,
                            const fc_scalars *, fct_handover);
  kernel_t k = nullptr;
#define FCT_PICK(M)                                                                                        \
  case M:                                                                                                  \
    k = pre8 == 2 ? (flags ? k_tile_walk<M, 3, ST, WT, 2, true> : k_tile_walk<M, 3, ST, WT, 2, false>)     \
      : flags     ? (pre8 ? k_tile_walk<M, 8, ST, WT, 2, true> : k_tile_walk<M, 4, ST, WT, 2, true>)       \
                  : (pre8 ? k_tile_walk<M, 8, ST, WT, 2, false> : k_tile_walk<M, 4, ST, WT, 2, false>);    \
    break;
  switch (mode) {
    FCT_PICK(TRI_FWD) FCT_PICK(TRI_BWD) FCT_PICK(TRI_DIC) FCT_PICK(TRI_DIC_PAR) FCT_PICK(TRI_DILU)
    default: return -1;
  }
#undef FCT_PICK
  std::vector<unsigned int> done(D.nlev, 0), ready(D.nlev, 0), flag(D.nblocks, 0);
  std::barrier<> bar(ST), wbar(WT);
  fct_bar = &bar;
  fct_walk_bar = &wbar;
  for (int sweep = 1; sweep <= 2; ++sweep) {
    // NaN, so that a value read before it was produced cannot go unnoticed
    for (int i = 0; i < n; ++i) out[i] = std::numeric_limits<double>::quiet_NaN();
    const fct_handover H{D.blk_level.data(), D.lev_blocks_before.data(), D.prod.data(), D.prod_cnt.data(), done.data(),
                         ready.data(), flag.data(), (unsigned int)sweep};
    const unsigned int base = (unsigned int)(sweep - 1) * (unsigned int)D.nblocks;
    std::vector<std::thread> th;
    th.reserve(ST);
    for (int t = 0; t < ST; ++t)
      th.emplace_back([&, t]() {
        fct_tid = (unsigned)t;
        for (int b = 0; b < D.nblocks; ++b) {
          k((const int4 *)D.meta_rm.data(), D.blk_nlev.data(), &ticket, base, S.tja.data(), diag, tpos, a, d, in, out,
            small, padd, nullptr, H);
          bar.arrive_and_wait();   // next CTA: the static "shared" arrays are reused
        }
      });
    for (auto &x : th) x.join();
  }
  fct_bar = nullptr;
  fct_walk_bar = nullptr;
  return 0;
}

// k_tile_walk_vf (256 staging / helper threads + 64 walkers, value-as-flag hand-over), CTA by CTA in ticket order.
// Same checks as fct_emu_sweep_vf: the result, and that the forward sweep re-arms `arm`, the backward sweep its input.
int fct_emu_walk_vf(void *h, int mode, int pre8, int n, const int *ioffset, const int *diag, const int *tpos,
                    const double *a, const double *d, const double *in, double *out, double small, double padd) {
  const fc_tile_schedule &S = *(fc_tile_schedule *)h;
  const fc_tile_dir &D = mode == TRI_BWD ? S.upper : S.lower;
  unsigned int ticket = 0;
  constexpr int ST = 256, WT = 64;
  using kernel_t = void (*)(const int4 *, const int *, unsigned int *, unsigned int, const int *, const int *,
                            const int *, const double *, const double *, double *, double *, double *, double, double,
                            const fc_scalars *, bool, unsigned int);
  kernel_t k = nullptr;
#define FCT_PICK(M)                                                                                   \
  case M:                                                                                             \
    k = pre8 == 2 ? k_tile_walk_vf<M, 3, ST, WT, 2> : pre8 ? k_tile_walk_vf<M, 8, ST, WT, 2>          \
                                                           : k_tile_walk_vf<M, 4, ST, WT, 2>;         \
    break;
  switch (mode) {
    FCT_PICK(TRI_FWD) FCT_PICK(TRI_BWD) FCT_PICK(TRI_DIC) FCT_PICK(TRI_DIC_PAR) FCT_PICK(TRI_DILU)
    default: return -1;
  }
#undef FCT_PICK
  std::vector<double> in_copy(in, in + n), arm(n, 0.0);
  for (int i = 0; i < n; ++i) out[i] = fct_unset();
  std::barrier<> bar(ST + WT), wbar(WT);
  fct_bar = &bar;
  fct_walk_bar = &wbar;
  std::vector<std::thread> th;
  th.reserve(ST + WT);
  for (int t = 0; t < ST + WT; ++t)
    th.emplace_back([&, t]() {
      fct_tid = (unsigned)t;
      for (int b = 0; b < D.nblocks; ++b) {
        k((const int4 *)D.meta_rm.data(), D.blk_nlev.data(), &ticket, 0u, S.tja.data(), diag, tpos, a, d, in_copy.data(),
          out, mode == TRI_FWD ? arm.data() : nullptr, small, padd, nullptr, true, 0u);
        bar.arrive_and_wait();   // next CTA: the static "shared" arrays are reused
      }
    });
  for (auto &x : th) x.join();
  fct_bar = nullptr;
  fct_walk_bar = nullptr;
  int bad = 0;
  if (mode == TRI_FWD) for (int i = 0; i < n; ++i) bad += fct_is_unset(arm[i]) ? 0 : 1;
  if (mode == TRI_BWD) for (int i = 0; i < n; ++i) bad += fct_is_unset(in_copy[i]) ? 0 : 1;
  for (int i = 0; i < n; ++i) bad += fct_is_unset(out[i]) ? 1 : 0;
  return bad;
}

}  // extern "C"
