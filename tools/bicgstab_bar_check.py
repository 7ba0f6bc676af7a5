#!/usr/bin/env python
"""Offline companion of tests/mgpu_check.py (no GPU): recompute, with the lock-step oracle alone, the BiCGStab bar of the
8-rank pressure-correction check -- 4 x the largest change of a field under six one-ulp perturbations of the input
velocities -- and hold the GPU-vs-oracle deviations recorded by an 8-GPU run (a log of tests/mgpu_check.py) against it.
    python tools/bicgstab_bar_check.py profiles/r02_mgpu_n8_final.log"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freecappuccino_b200 import cases, mesh as M  # noqa: E402
from oracle import oracle as O, oracle_par as OP  # noqa: E402

ns = {"__file__": os.path.join(ROOT, "tests", "mgpu_check.py"), "__name__": "mgpu_check_helpers"}
exec(open(ns["__file__"]).read().split("def main")[0], ns)
scatter_case = ns["scatter_case"]

log = open(sys.argv[1]).read()
world = 8
gpu = {}
for mesh_name, rank, field, err in re.findall(r"FAIL (\w+)/bicgstab/npcor1: rank (\d+) field (\w+) rel L2 ([0-9.e+-]+)", log):
    gpu[(mesh_name, field)] = max(gpu.get((mesh_name, field), 0.0), float(err))
ok = True
for name, g in (("hex_mixed", cases.hex_case(12, 9, 4 * world, kinds=("inlet", "outlet", "wall", "symmetry", "wall", "wall"))),
                ("skew", cases.skew_case(9, 8, 3 * world + 2))):
    f = cases.flow_fields(g)
    fmi, flomas = cases.inlet_fluxes(g, f)
    gp = O.grad_gauss(g, f["p"], 1)
    parts = M.partition(g, M.slab_ranks(g.numCells, world), world)
    oo = O.calcp_opts(solver="bicgstab", flomas=flomas, npcor=1, lsq_flag=False, nigrad=1, sor=1e-7, nsw=2000, flux_variant=0)
    oo.sol.parallel = 1
    runs = []
    for seed in range(0, 7):
        rng = np.random.default_rng(seed)
        pc = OP.ParCase(parts)
        for m, fl in zip(parts, pc.fields):
            sc = scatter_case(g, m, f, fmi, gp)
            for k in ("u", "v", "w", "p", "den", "apu", "apv", "apw", "dPdxi"):
                getattr(fl, k)[:] = sc[k]
            if seed:
                for k in ("u", "v", "w"):
                    arr = getattr(fl, k)
                    arr[:] = arr * (1.0 + 2.2e-16 * rng.integers(-1, 2, arr.size))
            fl.fmi[:sc["fmi"].size] = sc["fmi"]
        rep = pc.calcp(oo)
        runs.append((rep.rep[0].iters, {k: [np.array(getattr(pc.fields[r], k)).copy() for r in range(world)]
                                        for k in ("u", "v", "w", "p", "pp", "flmass")}))
    for k in ("u", "v", "w", "p", "pp", "flmass"):
        sens = max(cases.rel_l2(a, b) for it, out in runs[1:] for a, b in zip(out[k], runs[0][1][k]))
        bar = max(1e-10, 4.0 * sens)
        dev = gpu.get((name, k), 0.0)
        ok = ok and dev <= bar
        print(f"{name:10s} {k:7s} oracle iterations {runs[0][0]} (perturbed: {sorted(set(r[0] for r in runs[1:]))})  "
              f"largest change under one-ulp noise {sens:.2e}  bar {bar:.2e}  GPU vs oracle (8 GPUs, worst rank) {dev:.2e}  "
              f"{'within' if dev <= bar else 'ABOVE'}")
print("ALL WITHIN THE BAR" if ok else "ABOVE THE BAR")
