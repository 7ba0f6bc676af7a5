#!/usr/bin/env python
"""Per-solver timing on the synthetic hex Poisson problem (BASELINE config 3 and the 216^3 size):
iterations to rsm < 1e-8 and ms per iteration for dpcg / iccg / bicgstab on one GPU.
    python tools/solver_bench.py [n ...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freecappuccino_b200 import cases, lib, mesh as M  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [100]
    for n in sizes:
        m = M.hex_mesh(n, n, n, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry"))
        ctx = lib.Context(0)
        ctx.set_mesh(m)
        t0 = time.perf_counter()
        ctx.create_csr(download=False)
        ctx.synchronize()
        t_csr = time.perf_counter() - t0
        su = cases.poisson_rhs(m)
        ctx.upload("APU", -np.ones(m.numCells))
        for solver in os.environ.get("SOLVERS", "dpcg,iccg").split(","):
            ctx.upload("SU", su)
            ctx.fill("PP", 0.0)
            ctx.laplacian("APU", "PP")
            best = None
            for rep_i in range(2):     # first run builds the level schedule
                ctx.fill("PP", 0.0)
                t0 = time.perf_counter()
                rep = ctx.solve(solver, "PP", lib.solver_opts(1e-8, 20000))
                wall = time.perf_counter() - t0
                t = ctx.timings()
                best = dict(n=n, cells=m.numCells, solver=solver, iters=rep.iters, solve_ms=t.solve_ms,
                            ms_per_iter=t.solve_ms / max(rep.iters, 1), wall_s=wall, res0=rep.res0, resl=rep.resl,
                            csr_build_s=t_csr)
            exact = np.sin(2 * np.pi * m.xc) * np.sin(2 * np.pi * m.yc)
            best["linf"] = float(np.max(np.abs(ctx.download("PP")[:m.numCells] - exact)))
            print(json.dumps(best), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
