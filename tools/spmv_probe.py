#!/usr/bin/env python
"""Short run for ncu: a few stand-alone TMA-pipeline SpMV launches and one short persistent DPCG solve on the
config-4 p' system (python tools/spmv_probe.py n [iters] with FCAPP_TUNE for the knobs)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freecappuccino_b200 import lib  # noqa: E402
from tools.sweep_bench import setup  # noqa: E402

n = int(sys.argv[1])
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
m, ctx = setup(n)
ctx.upload("USER0", np.random.default_rng(1).standard_normal(m.numTotal))
ctx.set_tuning(lib.TUNE_SPMV_KERNEL, 1)
print("spmv ms", ctx.time_spmv("USER0", "SCRATCH_T", 6))
for _ in range(2):
    ctx.fill("PP", 0.0)
    rep = ctx.solve("dpcg", "PP", lib.solver_opts(1e-30, iters))
t = ctx.timings()
print("dpcg iters", rep.iters, "us/iter", 1e3 * t.solve_ms / max(rep.iters, 1), "index bytes", t.persist_index_bytes)
ctx.close()
