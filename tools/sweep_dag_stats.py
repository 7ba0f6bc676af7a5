#!/usr/bin/env python
"""Design groundwork for point-to-point (flag per block) triangular sweeps: for the level schedule the library builds
(rows sorted by level, 128-row blocks that never straddle a level; fc_trisolve.cu) count, for the strict lower triangle,
how many PRODUCER blocks each block gathers from and how far back they lie.  CPU only (scipy).
    python tools/sweep_dag_stats.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freecappuccino_b200 import cases  # noqa: E402

B = 128


def stats(name, mesh):
    n, F = mesh.numCells, mesh.numInnerFaces
    o, nb = mesh.owner[:F] - 1, mesh.neighbour - 1
    lo, hi = np.minimum(o, nb), np.maximum(o, nb)          # row hi depends on row lo (lower triangle)
    order = np.argsort(hi, kind="stable")
    lo, hi = lo[order], hi[order]
    level = np.zeros(n, np.int64)
    ptr = np.searchsorted(hi, np.arange(n + 1))
    for i in range(n):                                       # natural ordering: dependencies have smaller indices
        s, e = ptr[i], ptr[i + 1]
        if e > s:
            level[i] = level[lo[s:e]].max() + 1
    nlev = int(level.max()) + 1
    rows = np.argsort(level, kind="stable")
    counts = np.bincount(level, minlength=nlev)
    blocks_per_level = (counts + B - 1) // B
    first_block = np.concatenate([[0], np.cumsum(blocks_per_level)])
    pos_in_level = np.arange(n) - np.repeat(np.cumsum(counts) - counts, counts)
    block_of = np.empty(n, np.int64)
    block_of[rows] = first_block[level[rows]] + pos_in_level // B
    nblocks = int(first_block[-1])
    prod = [set() for _ in range(nblocks)]
    for c, p in zip(block_of[hi], block_of[lo]):
        prod[c].add(p)
    # deadlock freedom of the ticket scheme: a block only waits for blocks with smaller numbers (drawn earlier)
    assert all(p < b for b, s in enumerate(prod) for p in s)
    k = np.array([len(s) for s in prod])
    back = np.array([max((b - min(s)) if s else 0 for b, s in [(b, prod[b])]) for b in range(nblocks)])
    blk_level = np.repeat(np.arange(nlev), blocks_per_level)
    lev_dist = np.array([max((blk_level[b] - blk_level[p]) for p in prod[b]) if prod[b] else 0 for b in range(nblocks)])
    print(f"{name}: {n} rows, {nlev} levels, {nblocks} blocks; producers per block mean {k.mean():.2f} max {k.max()} "
          f"(p99 {np.percentile(k, 99):.0f}); furthest producer {back.max()} blocks / {lev_dist.max()} levels back; "
          f"blocks with <= 8 producers {100.0 * (k <= 8).mean():.1f} %, <= 16: {100.0 * (k <= 16).mean():.1f} %")


if __name__ == "__main__":
    stats("hex 48^3", cases.hex_case(48, 48, 48))
    stats("hex 96x96x1 slab", cases.hex_case(96, 96, 1))
    stats("poly 2*16^3", cases.poly_case(16))
    g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pitzDaily.npz")
    stats("pitzDaily", cases.golden_mesh(g))
