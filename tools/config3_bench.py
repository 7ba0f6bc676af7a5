#!/usr/bin/env python
"""BASELINE config 3: synthetic 100^3 structured-hex Poisson problem (the path of src/poisson.f90:55-86: laplacian(mu,p),
su = 8 pi^2 sin(2 pi x) sin(2 pi y) vol, mu = -1, p = 0 start, walls in x and y, symmetry in z), DPCG and ICCG to
rsm < 1e-8 on one GPU.  One JSON line: iterations, ms per iteration, algorithmic GB/s and the fraction of the measured
HBM peak for both solvers, the oracle's iteration counts on the same system (test infrastructure, checker only) and the
L-infinity error against the analytic solution sin(2 pi x) sin(2 pi y).

    python tools/config3_bench.py [n] [--no-oracle]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freecappuccino_b200 import cases, lib  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 100
    with_oracle = "--no-oracle" not in sys.argv
    mesh = cases.hex_case(n, n, n, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry"))
    nc, nnz = mesh.numCells, mesh.nnz
    su = cases.poisson_rhs(mesh)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    ctx = lib.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    exact = np.sin(2 * np.pi * mesh.xc[:nc]) * np.sin(2 * np.pi * mesh.yc[:nc])
    out = {"workload": f"config 3: {n}^3 hex Poisson (poisson.f90 path), rsm < 1e-8, nsw 10000", "cells": nc, "nnz": nnz,
           "hbm_peak_gbs": peak, "solvers": {}}
    ref = {}
    if with_oracle:
        from oracle import oracle
        csr = oracle.create_csr(mesh)
        su_ref = su.copy()
        a = oracle.laplacian(mesh, csr, -np.ones(nc), np.zeros(mesh.numTotal), su_ref)
        for solver in ("dpcg", "iccg"):
            fi = np.zeros(mesh.numTotal)
            t0 = time.perf_counter()
            res0, resl, iters, _ = oracle.solve(solver, csr, a, su_ref, fi, sor=1e-8, nsw=10000)
            ref[solver] = dict(iters=int(iters), res0=res0, fi=fi[:nc].copy(), seconds=time.perf_counter() - t0)
    for solver, per_iter in (("dpcg", 12 * nnz + 116 * nc), ("iccg", 24 * nnz + 164 * nc)):
        best = None
        for _ in range(3):
            ctx.upload("APU", -np.ones(nc))
            ctx.upload("SU", su)
            ctx.fill("PP", 0.0)
            ctx.laplacian("APU", "PP")
            rep = ctx.solve(solver, "PP", lib.solver_opts(1e-8, 10000))
            ms = ctx.timings().solve_ms
            best = ms if best is None else min(best, ms)
        got = ctx.download("PP")[:nc]
        ms_it = best / max(rep.iters, 1)
        o = {"iters": int(rep.iters), "res0": rep.res0, "resl": rep.resl, "solve_ms": best, "ms_per_iteration": ms_it,
             "algorithmic_gbs": per_iter / ms_it / 1e6, "frac_of_hbm_peak": per_iter / ms_it / 1e6 / peak,
             "linf_error_vs_analytic": float(np.max(np.abs(got - exact)))}
        if solver in ref:
            o.update({"iters_oracle": ref[solver]["iters"], "iters_within_1": abs(int(rep.iters) - ref[solver]["iters"]) <= 1,
                      "res0_oracle": ref[solver]["res0"], "rel_l2_vs_oracle": cases.rel_l2(got, ref[solver]["fi"]),
                      "oracle_seconds_1_core": ref[solver]["seconds"]})
        if solver == "iccg":
            o["sweep_schedule"] = ctx.sweep_schedule_info()[:200]
        out["solvers"][solver] = o
    ctx.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
