// Two-level ("tiled") schedule of the DIC / DILU triangular sweeps -- plain C++, no CUDA, so that the same code is
// compiled into the library (fc_trisolve.cu) and into the g++ test harness (tests/kernel_bodies_host).
//
// The level schedule of fc_trisolve.cu needs one grid-wide hand-over per dependency level: 646 of them on the
// 216^3 hexahedral mesh (3*216-2 hyperplanes), about 4 us each, although a level holds only ~16 k rows.  Here the
// rows are grouped into spatial tiles of at most FC_TILE cells (bins of ~8 cells per axis over the cell centres); a
// CTA owns a tile and walks the tile's *local* dependency levels with __syncthreads and shared memory, and only the
// tile-to-tile dependencies (3*27-2 = 79 levels at 216^3) go through global memory.  The preconditioner is unchanged:
// rows keep their natural numbering (iccg.f90:77-111, bicgstab.f90:68-79, :117-136), every row is still summed left
// to right, only the order in which independent rows are visited differs.
//
// The tile-to-tile dependency graph must be acyclic and no tile may exceed FC_TILE rows.  Bins over a
// lexicographically numbered structured mesh satisfy both; anywhere else (block-structured or renumbered meshes,
// strong grading) the offending bins are found (strongly connected components) and cut into runs of consecutive row
// numbers, which are acyclic by construction (repair_tiles).  The result is always valid (build_dir re-checks); it is
// not always *good* -- a mesh whose numbering ignores space altogether ends up as one long chain -- so the schedule
// carries a critical-path estimate (`cost`) and the caller keeps the level schedule unless the tiling is clearly
// cheaper.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

constexpr int FC_TILE = 512;       // rows per tile = threads per CTA of k_tile_sweep
constexpr int FC_TILE_MAXP = 16;   // producer tiles a tile can name for the point-to-point hand-over

struct fc_tile_dir {                     // one sweep direction (strict lower or strict upper triangle)
  int nlev = 0;                          // tile levels
  int nblocks = 0;                       // tiles, in ticket order (tile-level-major)
  int max_local_levels = 0;
  std::vector<int> rows;                 // [nblocks * FC_TILE] row id (0-based) or -1 (padding, at the end of a tile);
                                         // slots of a tile are ordered by local level, then row id
  std::vector<int> llev;                 // [nblocks * FC_TILE] local level of that row, -1 for padding
  std::vector<int> meta;                 // [nblocks * FC_TILE * 4] what the kernel reads per slot in one 16-byte load:
                                         // row, local level, first and one-past-last position of its triangle in a / tja
  std::vector<int> meta_rm;              // [nblocks * FC_TILE * 4] the same rows in ASCENDING ROW ORDER inside the tile (the
                                         // order in which their matrix entries lie in memory): row, slot | level << 16,
                                         // triangle [s, e); -1 padding.  k_tile_walk stages the tile with these (coalesced)
  std::vector<int> blk_nlev;             // [nblocks] local levels of the tile
  std::vector<int> blk_level;            // [nblocks] tile level
  std::vector<int> lev_blocks_before;    // [nlev + 1] tiles in tile levels < L
  // point-to-point hand-over: the tiles (block numbers, all smaller than the tile's own) whose rows a tile reads
  // through global memory; p2p_ok = every tile has at most FC_TILE_MAXP of them
  bool p2p_ok = false;
  int max_producers = 0;
  std::vector<int> prod;                 // [nblocks * FC_TILE_MAXP]
  std::vector<int> prod_cnt;             // [nblocks]
};

struct fc_tile_schedule {
  bool ok = false;
  std::string why;                       // why not, when !ok
  int ntiles = 0;
  int cells_per_axis = 0;                // the bin width that worked
  int max_tile_rows = 0;
  int repaired_rows = 0;                 // rows of bins that had to be cut into runs of consecutive row numbers
  long long cost = 0;                    // critical path estimate in 0.1 us: see fc_tile_cost
  int max_tri_len = 0;                   // longest strict-triangle row (how many entries the kernel keeps in registers)
  std::vector<int> tja;                  // [nnz] column j, or -(q+1) when row j is slot q of the same tile IN THE DIRECTION
                                         // THAT READS THE ENTRY (strict lower triangle: lower.rows, upper: upper.rows)
  fc_tile_dir lower, upper;
};

namespace fc_tile_detail {

// tile id of every row from bins over the cell centres; returns the number of (non-empty, renumbered) tiles
// `shrink` = 0, 1, 2 ...: bins of FC_TILE^(1/dims) cells per axis (8 in 3-D, 22 in 2-D), each step 1/8 narrower
inline int assign_tiles(int n, const int *ioffset, const int *ja, const int *diag, const double *xc, const double *yc,
                        const double *zc, int shrink, std::vector<int> &tile, int &target) {
  const double *c[3] = {xc, yc, zc};
  double lo[3], len[3];
  int dims = 0;
  double vol = 1.0;
  double longest = 0.0;
  for (int ax = 0; ax < 3; ++ax) {
    double mn = c[ax][0], mx = c[ax][0];
    for (int i = 1; i < n; ++i) { mn = std::min(mn, c[ax][i]); mx = std::max(mx, c[ax][i]); }
    lo[ax] = mn; len[ax] = mx - mn;
    longest = std::max(longest, len[ax]);
  }
  for (int ax = 0; ax < 3; ++ax) {
    if (len[ax] <= 1e-9 * longest) len[ax] = 0.0;   // one layer of cells (2-D cases): rounding noise is not an extent
    if (len[ax] > 0.0) { vol *= len[ax]; ++dims; }
  }
  int nb[3] = {1, 1, 1};
  target = FC_TILE;
  if (dims > 0) {
    target = (int)std::floor(std::pow((double)FC_TILE, 1.0 / dims) + 1e-9);
    target = std::max(2, target - (shrink * std::max(1, target / 8)));
    // mean spacing; the box of the centres is one spacing short of the box of the cells
    const double h0 = std::pow(vol / n, 1.0 / dims);
    double vol1 = 1.0;
    for (int ax = 0; ax < 3; ++ax)
      if (len[ax] > 0.0) { lo[ax] -= 0.5 * h0; len[ax] += h0; vol1 *= len[ax]; }
    const double h = std::pow(vol1 / n, 1.0 / dims);
    for (int ax = 0; ax < 3; ++ax)
      if (len[ax] > 0.0) nb[ax] = std::max(1, (int)std::ceil(len[ax] / (target * h) - 1e-9));
  }
  // bin coordinates, then made monotone along the dependencies: a row never sits in a lower bin (in any direction)
  // than a row it depends on.  With all three coordinates non-decreasing along every edge, tiles cannot depend on
  // each other in a circle.  A lexicographically numbered mesh is monotone already; the pass matters for cells whose
  // centre lies on a bin boundary (jittered or unstructured meshes), which would otherwise land on either side at random.
  std::vector<int> bin(3 * (size_t)n);
  for (int i = 0; i < n; ++i)
    for (int ax = 0; ax < 3; ++ax) {
      int b = 0;
      if (len[ax] > 0.0) b = std::min(nb[ax] - 1, std::max(0, (int)((c[ax][i] - lo[ax]) / len[ax] * nb[ax])));
      bin[3 * (size_t)i + ax] = b;
    }
  for (int i = 0; i < n; ++i)
    for (int k = ioffset[i]; k < diag[i]; ++k) {
      const int j = ja[k];
      if (j >= i) continue;   // the strict lower triangle only
      for (int ax = 0; ax < 3; ++ax) bin[3 * (size_t)i + ax] = std::max(bin[3 * (size_t)i + ax], bin[3 * (size_t)j + ax]);
    }
  std::vector<long long> raw(n);
  for (int i = 0; i < n; ++i)
    raw[i] = bin[3 * (size_t)i] + (long long)nb[0] * (bin[3 * (size_t)i + 1] + (long long)nb[1] * bin[3 * (size_t)i + 2]);
  // renumber the non-empty bins densely, in ascending bin order
  std::vector<long long> used(raw);
  std::sort(used.begin(), used.end());
  used.erase(std::unique(used.begin(), used.end()), used.end());
  tile.resize(n);
  for (int i = 0; i < n; ++i) tile[i] = (int)(std::lower_bound(used.begin(), used.end(), raw[i]) - used.begin());
  return (int)used.size();
}

// dependency range of row i in the triangle of `lower`
inline void tri_range(const int *ioffset, const int *diag, int i, bool lower, int &s, int &e) {
  if (lower) { s = ioffset[i]; e = diag[i]; }
  else { s = diag[i] + 1; e = ioffset[i + 1]; }
}

// `pos` (out): slot of every row inside its tile for this direction -- rows ordered by local level, then row id, so
// that the rows of one local level are consecutive slots (k_tile_walk hands a level to consecutive threads)
inline bool build_dir(int n, const int *ioffset, const int *ja, const int *diag, const std::vector<int> &tile,
                      std::vector<int> &pos, int ntiles, bool lower, fc_tile_dir &D, std::string &why) {
  // tile-to-tile edges producer -> consumer
  std::vector<uint64_t> edges;
  for (int i = 0; i < n; ++i) {
    int s, e;
    tri_range(ioffset, diag, i, lower, s, e);
    for (int k = s; k < e; ++k) {
      const int j = ja[k];
      if (j < n && tile[j] != tile[i]) edges.push_back(((uint64_t)(uint32_t)tile[j] << 32) | (uint32_t)tile[i]);
    }
    if ((i & 0xfffff) == 0xfffff) {   // keep the list short on large meshes
      std::sort(edges.begin(), edges.end());
      edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    }
  }
  std::sort(edges.begin(), edges.end());
  edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
  std::vector<int> indeg(ntiles, 0), adj_off(ntiles + 1, 0), adj(edges.size());
  for (uint64_t ed : edges) { adj_off[(ed >> 32) + 1]++; indeg[(uint32_t)ed]++; }
  for (int t = 0; t < ntiles; ++t) adj_off[t + 1] += adj_off[t];
  {
    std::vector<int> fill(adj_off.begin(), adj_off.end() - 1);
    for (uint64_t ed : edges) adj[fill[ed >> 32]++] = (int)(uint32_t)ed;
  }
  // Kahn: tile level = longest path from a source
  std::vector<int> tlev(ntiles, 0), queue;
  queue.reserve(ntiles);
  for (int t = 0; t < ntiles; ++t)
    if (indeg[t] == 0) queue.push_back(t);
  for (size_t h = 0; h < queue.size(); ++h) {
    const int t = queue[h];
    for (int q = adj_off[t]; q < adj_off[t + 1]; ++q) {
      const int c = adj[q];
      tlev[c] = std::max(tlev[c], tlev[t] + 1);
      if (--indeg[c] == 0) queue.push_back(c);
    }
  }
  if ((int)queue.size() != ntiles) {
    why = "the tile-to-tile dependency graph has a cycle (cell numbering not monotone across the bins)";
    return false;
  }
  // local levels: only dependencies inside the tile count, the others are complete before the tile starts
  std::vector<int> ll(n, 0);
  if (lower) {
    for (int i = 0; i < n; ++i) {
      int s, e, l = 0;
      tri_range(ioffset, diag, i, true, s, e);
      for (int k = s; k < e; ++k) {
        const int j = ja[k];
        if (j < n && tile[j] == tile[i]) l = std::max(l, ll[j] + 1);
      }
      ll[i] = l;
    }
  } else {
    for (int i = n - 1; i >= 0; --i) {
      int s, e, l = 0;
      tri_range(ioffset, diag, i, false, s, e);
      for (int k = s; k < e; ++k) {
        const int j = ja[k];
        if (j < n && tile[j] == tile[i]) l = std::max(l, ll[j] + 1);
      }
      ll[i] = l;
    }
  }
  // slots: counting sort of every tile's rows by local level (ascending row id inside a level)
  {
    std::vector<int> tnl(ntiles, 0);
    for (int i = 0; i < n; ++i) tnl[tile[i]] = std::max(tnl[tile[i]], ll[i] + 1);
    std::vector<size_t> hoff(ntiles + 1, 0);
    for (int t = 0; t < ntiles; ++t) hoff[t + 1] = hoff[t] + (size_t)tnl[t];
    std::vector<int> cnt(hoff[ntiles], 0);
    for (int i = 0; i < n; ++i) cnt[hoff[tile[i]] + ll[i]]++;
    for (int t = 0; t < ntiles; ++t) {   // exclusive prefix inside the tile
      int run = 0;
      for (size_t q = hoff[t]; q < hoff[t + 1]; ++q) { const int c = cnt[q]; cnt[q] = run; run += c; }
    }
    pos.assign(n, 0);
    for (int i = 0; i < n; ++i) pos[i] = cnt[hoff[tile[i]] + ll[i]]++;
  }
  // tiles in ticket order: by tile level, then tile id
  D.nlev = 0;
  for (int t = 0; t < ntiles; ++t) D.nlev = std::max(D.nlev, tlev[t] + 1);
  D.nblocks = ntiles;
  D.lev_blocks_before.assign(D.nlev + 1, 0);
  for (int t = 0; t < ntiles; ++t) D.lev_blocks_before[tlev[t] + 1]++;
  for (int l = 0; l < D.nlev; ++l) D.lev_blocks_before[l + 1] += D.lev_blocks_before[l];
  std::vector<int> block_of_tile(ntiles), fill(D.lev_blocks_before.begin(), D.lev_blocks_before.end() - 1);
  for (int t = 0; t < ntiles; ++t) block_of_tile[t] = fill[tlev[t]]++;
  D.rows.assign((size_t)ntiles * FC_TILE, -1);
  D.llev.assign((size_t)ntiles * FC_TILE, -1);
  D.meta.assign((size_t)ntiles * FC_TILE * 4, -1);
  D.meta_rm.assign((size_t)ntiles * FC_TILE * 4, -1);
  std::vector<int> rm_fill(ntiles, 0);
  D.blk_nlev.assign(ntiles, 0);
  D.blk_level.assign(ntiles, 0);
  D.max_local_levels = 0;
  for (int t = 0; t < ntiles; ++t) D.blk_level[block_of_tile[t]] = tlev[t];
  for (int i = 0; i < n; ++i) {
    const int b = block_of_tile[tile[i]];
    const size_t slot = (size_t)b * FC_TILE + pos[i];
    D.rows[slot] = i;
    D.llev[slot] = ll[i];
    int ts, te;
    tri_range(ioffset, diag, i, lower, ts, te);
    D.meta[4 * slot] = i; D.meta[4 * slot + 1] = ll[i]; D.meta[4 * slot + 2] = ts; D.meta[4 * slot + 3] = te;
    const size_t rm = (size_t)b * FC_TILE + rm_fill[b]++;   // rows are visited in ascending order
    D.meta_rm[4 * rm] = i; D.meta_rm[4 * rm + 1] = pos[i] | (ll[i] << 16); D.meta_rm[4 * rm + 2] = ts; D.meta_rm[4 * rm + 3] = te;
    D.blk_nlev[b] = std::max(D.blk_nlev[b], ll[i] + 1);
    D.max_local_levels = std::max(D.max_local_levels, ll[i] + 1);
  }
  D.prod.assign((size_t)ntiles * FC_TILE_MAXP, -1);
  D.prod_cnt.assign(ntiles, 0);
  D.p2p_ok = true;
  D.max_producers = 0;
  std::vector<int> nprod(ntiles, 0);
  for (uint64_t ed : edges) {   // unique (producer, consumer) pairs
    const int pb = block_of_tile[ed >> 32], cb = block_of_tile[(uint32_t)ed];
    if (nprod[cb] < FC_TILE_MAXP) D.prod[(size_t)cb * FC_TILE_MAXP + nprod[cb]] = pb;
    nprod[cb]++;
  }
  for (int b = 0; b < ntiles; ++b) {
    D.max_producers = std::max(D.max_producers, nprod[b]);
    D.prod_cnt[b] = std::min(nprod[b], FC_TILE_MAXP);
    if (nprod[b] > FC_TILE_MAXP) D.p2p_ok = false;
  }
  return true;
}

// Make any bin assignment usable: bins that depend on each other in a circle (cell numbering not monotone across
// them -- multi-block or renumbered meshes) are merged into one group, and every group that is merged or larger than
// FC_TILE rows is cut again into runs of consecutive row numbers.  Runs of one group depend on each other only from
// low to high row number (the matrix is triangular in the natural numbering), and a tile outside a group is either
// upstream or downstream of the whole group (otherwise it would belong to it), so the result has no cycle.
// Returns the new number of tiles; `repaired` counts the rows that ended up in such runs.
inline int repair_tiles(int n, const int *ioffset, const int *ja, const int *diag, std::vector<int> &tile, int ntiles,
                        int &repaired) {
  // tile graph of the lower triangle (the upper one is its reverse: same strongly connected components)
  std::vector<uint64_t> edges;
  for (int i = 0; i < n; ++i) {
    for (int k = ioffset[i]; k < diag[i]; ++k) {
      const int j = ja[k];
      if (j < n && tile[j] != tile[i]) edges.push_back(((uint64_t)(uint32_t)tile[j] << 32) | (uint32_t)tile[i]);
    }
    if ((i & 0xfffff) == 0xfffff) {
      std::sort(edges.begin(), edges.end());
      edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    }
  }
  std::sort(edges.begin(), edges.end());
  edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
  std::vector<int> off(ntiles + 1, 0), adj(edges.size());
  for (uint64_t ed : edges) off[(ed >> 32) + 1]++;
  for (int t = 0; t < ntiles; ++t) off[t + 1] += off[t];
  {
    std::vector<int> fill(off.begin(), off.end() - 1);
    for (uint64_t ed : edges) adj[fill[ed >> 32]++] = (int)(uint32_t)ed;
  }
  // Tarjan's strongly connected components, iterative
  std::vector<int> index(ntiles, -1), low(ntiles, 0), comp(ntiles, -1), stack, next(ntiles, 0), call;
  std::vector<char> on_stack(ntiles, 0);
  int counter = 0, ncomp = 0;
  for (int root = 0; root < ntiles; ++root) {
    if (index[root] >= 0) continue;
    call.push_back(root);
    while (!call.empty()) {
      const int v = call.back();
      if (index[v] < 0) { index[v] = low[v] = counter++; stack.push_back(v); on_stack[v] = 1; next[v] = off[v]; }
      bool descended = false;
      while (next[v] < off[v + 1]) {
        const int w = adj[next[v]++];
        if (index[w] < 0) { call.push_back(w); descended = true; break; }
        if (on_stack[w]) low[v] = std::min(low[v], index[w]);
      }
      if (descended) continue;
      if (low[v] == index[v]) {
        for (;;) {
          const int w = stack.back();
          stack.pop_back();
          on_stack[w] = 0;
          comp[w] = ncomp;
          if (w == v) break;
        }
        ++ncomp;
      }
      call.pop_back();
      if (!call.empty()) low[call.back()] = std::min(low[call.back()], low[v]);
    }
  }
  std::vector<int> comp_tiles(ncomp, 0), comp_rows(ncomp, 0);
  {
    std::vector<char> seen(ntiles, 0);
    for (int i = 0; i < n; ++i) {
      const int t = tile[i];
      comp_rows[comp[t]]++;
      if (!seen[t]) { seen[t] = 1; comp_tiles[comp[t]]++; }
    }
  }
  // new numbering: untouched groups keep one tile; the others get ceil(rows / FC_TILE) runs of (nearly) equal length
  std::vector<int> first(ncomp, 0), runs(ncomp, 1);
  int total = 0;
  repaired = 0;
  for (int c = 0; c < ncomp; ++c) {
    const bool cut = comp_tiles[c] > 1 || comp_rows[c] > FC_TILE;
    runs[c] = cut ? (comp_rows[c] + FC_TILE - 1) / FC_TILE : 1;
    if (cut) repaired += comp_rows[c];
    first[c] = total;
    total += runs[c];
  }
  std::vector<int> seen_rows(ncomp, 0);
  for (int i = 0; i < n; ++i) {   // ascending row number: the q-th row of a group goes to run q * runs / rows
    const int c = comp[tile[i]];
    const int q = seen_rows[c]++;
    tile[i] = first[c] + (int)((long long)q * runs[c] / comp_rows[c]);
  }
  return total;
}

// critical path of one sweep in units of 0.1 us: per tile level one hand-over through global memory (~3.5 us) plus the
// local walk of its slowest tile (~0.15 us per local level)
inline long long dir_cost(const fc_tile_dir &D) {
  long long c = 0;
  for (int l = 0; l < D.nlev; ++l) {
    int worst = 0;
    for (int b = D.lev_blocks_before[l]; b < D.lev_blocks_before[l + 1]; ++b) worst = std::max(worst, D.blk_nlev[b]);
    c += 35 + (3 * (long long)worst + 1) / 2;
  }
  return c;
}

}  // namespace fc_tile_detail

// 0-based CSR (columns ascending inside a row, diag = position of the diagonal) and the cell centres of its rows
// `min_shrink` > 0 starts from narrower bins (7, 6, 5 ... cells per axis instead of 8): smaller tiles with fewer local
// levels each, more tile levels -- a measurement knob (FC_TILE_MIN_SHRINK in the library)
inline fc_tile_schedule fc_build_tile_schedule(int n, const int *ioffset, const int *ja, const int *diag,
                                               const double *xc, const double *yc, const double *zc,
                                               int min_shrink = 0) {
  using namespace fc_tile_detail;
  fc_tile_schedule S;
  if (n < 1) { S.why = "empty matrix"; return S; }
  std::vector<int> tile, count;
  int ntiles = 0;
  for (int shrink = std::max(0, std::min(min_shrink, 4)); shrink <= 4; ++shrink) {
    int target = 0;
    ntiles = assign_tiles(n, ioffset, ja, diag, xc, yc, zc, shrink, tile, target);
    count.assign(ntiles, 0);
    for (int i = 0; i < n; ++i) count[tile[i]]++;
    S.max_tile_rows = *std::max_element(count.begin(), count.end());
    S.cells_per_axis = target;
    if (S.max_tile_rows <= FC_TILE) break;
  }
  // circular bins (numbering not monotone across them) and oversized bins are cut into runs of consecutive rows
  ntiles = repair_tiles(n, ioffset, ja, diag, tile, ntiles, S.repaired_rows);
  count.assign(ntiles, 0);
  for (int i = 0; i < n; ++i) count[tile[i]]++;
  S.max_tile_rows = *std::max_element(count.begin(), count.end());
  if (S.max_tile_rows > FC_TILE || *std::min_element(count.begin(), count.end()) < 1) {
    S.why = "internal: tile repair left an empty or oversized tile";
    return S;
  }
  S.ntiles = ntiles;
  std::vector<int> pos_lower, pos_upper;
  if (!build_dir(n, ioffset, ja, diag, tile, pos_lower, ntiles, true, S.lower, S.why)) return S;
  if (!build_dir(n, ioffset, ja, diag, tile, pos_upper, ntiles, false, S.upper, S.why)) return S;
  // in-tile dependencies name the slot of the direction that reads them: an entry of the strict lower triangle is
  // only ever read by the forward / factor sweeps, one of the strict upper triangle by the backward sweep
  const int nnz = ioffset[n];
  S.tja.resize(nnz);
  for (int i = 0; i < n; ++i) {
    S.max_tri_len = std::max(S.max_tri_len, std::max(diag[i] - ioffset[i], ioffset[i + 1] - diag[i] - 1));
    for (int k = ioffset[i]; k < ioffset[i + 1]; ++k) {
      const int j = ja[k];
      const std::vector<int> &pos = k < diag[i] ? pos_lower : pos_upper;
      S.tja[k] = (j != i && j < n && tile[j] == tile[i]) ? -(pos[j] + 1) : j;
    }
  }
  S.cost = std::max(fc_tile_detail::dir_cost(S.lower), fc_tile_detail::dir_cost(S.upper));
  S.ok = true;
  return S;
}

// The same estimate for the one-level schedule of fc_trisolve.cu (one hand-over of about 4.2 us per row level, measured
// on B200): the caller keeps the level schedule unless the tiling promises to be clearly faster.
inline long long fc_level_cost(int row_levels) { return 42LL * row_levels; }
