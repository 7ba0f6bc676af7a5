#!/usr/bin/env python
"""A/B timing of the kernel variants on one GPU (synthetic hex Poisson, walls x/y, symmetry z):
SpMV as the CSR-stream kernel vs the TMA pipeline, DPCG as one launch per vector operation vs the
persistent cooperative kernel (with its per-phase clocks).
    python tools/variant_bench.py [n ...]          # default 216 (10.1 M cells) and 108 (1.26 M = the 8-GPU share)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freecappuccino_b200 import cases, lib, mesh as M  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [216, 108]
    ctas = [int(c) for c in os.environ.get("CTAS", "0").split(",")]
    geos = [int(c) for c in os.environ.get("GEOS", "0,1,2,3").split(",")]
    for n in sizes:
        m = M.hex_mesh(n, n, n, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry"))
        ctx = lib.Context(0)
        ctx.set_mesh(m)
        ctx.create_csr(download=False)
        su = cases.poisson_rhs(m)
        ctx.upload("APU", -np.ones(m.numCells))
        ctx.upload("SU", su)
        ctx.fill("PP", 0.0)
        ctx.laplacian("APU", "PP")
        ctx.upload("USER0", np.random.default_rng(1).standard_normal(m.numTotal))
        spmv_bytes = 12 * m.nnz + 20 * m.numCells
        iter_bytes = 12 * m.nnz + 116 * m.numCells
        ys = {}
        for variant, geo, name in [(0, 0, "stream")] + [(1, g, f"tma{g}") for g in geos]:
            ctx.set_tuning(lib.TUNE_SPMV_KERNEL, variant)
            ctx.set_tuning(lib.TUNE_PIPE_GEOMETRY, geo)
            ms = min(ctx.time_spmv("USER0", "SCRATCH_T", 50) for _ in range(3))
            ys[name] = ctx.download("SCRATCH_T")[:m.numCells]
            print(json.dumps(dict(n=n, kernel=f"spmv_{name}", ms=ms, gbs=spmv_bytes / ms / 1e6,
                                  bit_identical_to_stream=bool(np.array_equal(ys["stream"], ys[name])))), flush=True)
        exact = np.sin(2 * np.pi * m.xc) * np.sin(2 * np.pi * m.yc)
        for persist, spmv, c, geo in [(0, 0, 0, 0)] + [(1, 1, c, g) for c in ctas for g in geos]:
            ctx.set_tuning(lib.TUNE_DPCG_PERSISTENT, persist)
            ctx.set_tuning(lib.TUNE_PIPE_GEOMETRY, geo)
            ctx.set_tuning(lib.TUNE_SPMV_KERNEL, spmv)
            ctx.set_tuning(lib.TUNE_CTAS_PER_SM, c)
            best = None
            for _ in range(3):
                ctx.fill("PP", 0.0)
                rep = ctx.solve("dpcg", "PP", lib.solver_opts(1e-8, 20000))
                t = ctx.timings()
                r = dict(n=n, dpcg="persistent" if persist else "multi-kernel", spmv=f"tma{geo}" if spmv else "stream",
                         iters=rep.iters, solve_ms=t.solve_ms, ms_per_iter=t.solve_ms / max(rep.iters, 1),
                         iter_gbs=iter_bytes * rep.iters / t.solve_ms / 1e6, resl=rep.resl)
                if persist:
                    it = max(t.persist_iters, 1)
                    r.update(grid=t.persist_grid, kernel_ms=t.persist_ms, pupdate_us=1e3 * t.persist_pupdate_ms / it,
                             spmv_us=1e3 * t.persist_spmv_ms / it, update_us=1e3 * t.persist_update_ms / it,
                             sync_us=1e3 * (t.persist_ms - t.persist_pupdate_ms - t.persist_spmv_ms
                                            - t.persist_update_ms) / it,
                             spmv_phase_gbs=spmv_bytes / (t.persist_spmv_ms / it) / 1e6)
                if best is None or r["solve_ms"] < best["solve_ms"]:
                    best = r
            best["linf"] = float(np.max(np.abs(ctx.download("PP")[:m.numCells] - exact)))
            print(json.dumps(best), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
