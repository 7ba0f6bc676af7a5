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
scatter_case, noise_bars = ns["scatter_case"], ns["noise_bars"]

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
    pc = OP.ParCase(parts)
    for m, fl in zip(parts, pc.fields):
        sc = scatter_case(g, m, f, fmi, gp)
        for k in ("u", "v", "w", "p", "den", "apu", "apv", "apw", "dPdxi"):
            getattr(fl, k)[:] = sc[k]
        fl.fmi[:sc["fmi"].size] = sc["fmi"]
    rep = pc.calcp(oo)
    bars = noise_bars(parts, g, f, fmi, gp, oo, pc)   # the function tests/mgpu_check.py uses
    for k in ("u", "v", "w", "p", "pp", "flmass"):
        dev = gpu.get((name, k), 0.0)
        ok = ok and dev <= bars[k]
        print(f"{name:10s} {k:7s} oracle iterations {rep.rep[0].iters}  bar = 4 x largest change under one-ulp noise = {bars[k]:.2e}  "
              f"GPU vs oracle (8 GPUs, worst rank) {dev:.2e}  {'within' if dev <= bars[k] else 'ABOVE'}")
print("ALL WITHIN THE BAR" if ok else "ABOVE THE BAR")
