#!/usr/bin/env python
"""Device-resident SIMPLE iterations (calcuvw -> calcp) on a 3-D lid-driven cavity of n^3 hexahedra with the
settings of the shipped cavity case (examples/cavity/input: muscl-f, gauss gradients, urf 0.7/0.3, sor 1e-2,
nsw 20 for u,v,w and 100 for p, dpcg, bdf with timestep 1e20, viscos 0.01): device time of every phase of one
SIMPLE iteration, measured with CUDA events on the library's stream (fc_get_timings).
    python tools/simple_iter_bench.py [n] [warmup] [steps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freecappuccino_b200 import cases, lib  # noqa: E402


def measure(ctx, m, warm=2, steps=3):
    """`warm` + `steps` SIMPLE iterations from rest on the context's mesh `m` (a box with six walls, the y+ wall
    moving); returns the mean device / wall times of the timed ones.  Overwrites U, V, W, P, VIS, FLMASS."""
    nt, nc, F = m.numTotal, m.numCells, m.numInnerFaces
    B = m.numFaces - F
    u = np.zeros(nt)
    fs, sl = m.boundary_faces("wall"), m.boundary_slots("wall")
    lid = sl[m.ary[fs] > 0.5 * np.abs(m.ary[fs]).max()]
    u[lid] = 1.0
    ctx.upload("U", u)
    for name in ("V", "W", "P", "FLMASS", "UO", "VO", "WO", "DPDXI"):
        ctx.fill(name, 0.0)
    ctx.fill("DEN", 1.0)
    ctx.upload("VIS", np.full(nt, 0.01))
    uo = lib.calcuvw_opts(scheme="muscl-f", urf=(0.7,) * 3, sor=(1e-2,) * 3, nsw=(20,) * 3, bdf=True, btime=0.0,
                          timestep=1e20, viscos=0.01)
    po = lib.calcp_opts(solver="dpcg", sor=1e-2, nsw=100, urf_p=0.3, pRefCell=1, const_mflux=True)
    rows = []
    l0 = ctx.timings().launches
    for it in range(warm + steps):
        if it == warm:
            l0 = ctx.timings().launches
        t0 = time.perf_counter()
        ru = ctx.calcuvw(uo)
        tu = ctx.timings()
        t1 = time.perf_counter()
        rp = ctx.calcp(po)
        tp = ctx.timings()
        t2 = time.perf_counter()
        if it >= warm:
            rows.append(dict(uvw_assemble_ms=tu.uvw_assemble_ms, uvw_solve_ms=tu.uvw_solve_ms,
                             uvw_iters=[ru.rep[k].iters for k in range(3)], uvw_wall_ms=1e3 * (t1 - t0),
                             p_assemble_ms=tp.assemble_ms, p_solve_ms=tp.solve_ms, p_correct_ms=tp.correct_ms,
                             p_iters=rp.rep[0].iters, p_wall_ms=1e3 * (t2 - t1), res_u=ru.rep[0].res0,
                             res_p=rp.rep[0].res0))
    mean = lambda k: float(np.mean([r[k] for r in rows]))
    # algorithmic bytes of the explicit part of calcuvw (fc_momentum.cu header): 5 Gauss passes + face + row kernels
    grad = 64 * F + 64 * nc + 36 * B
    uvw_bytes = 5 * grad + (120 * F + 160 * nc) + (184 * F + 88 * nc)
    return dict(case="lid-driven cavity from rest, settings of examples/cavity/input (muscl-f, urf 0.7/0.3, sor 1e-2, "
                     "nsw 20/100, dpcg for p')", cells=nc, inner_faces=F, steps=steps, warmup=warm,
                simple_iter_ms=mean("uvw_wall_ms") + mean("p_wall_ms"),
                calcuvw=dict(assemble_ms=mean("uvw_assemble_ms"), solve_ms=mean("uvw_solve_ms"),
                             wall_ms=mean("uvw_wall_ms"), bicgstab_iters=rows[-1]["uvw_iters"],
                             assemble_algorithmic_gb=uvw_bytes / 1e9,
                             assemble_gbs=uvw_bytes / 1e6 / mean("uvw_assemble_ms")),
                calcp=dict(assemble_ms=mean("p_assemble_ms"), solve_ms=mean("p_solve_ms"),
                           correct_ms=mean("p_correct_ms"), wall_ms=mean("p_wall_ms"), dpcg_iters=rows[-1]["p_iters"]),
                residuals_last=dict(u=rows[-1]["res_u"], p=rows[-1]["res_p"]),
                umax=float(np.abs(ctx.download("U")[:nc]).max()),
                gpu_launches_per_iteration=int((ctx.timings().launches - l0) / max(steps, 1)))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    m = cases.hex_case(n, n, n)          # six walls; the y+ wall is the lid
    ctx = lib.Context(0)
    ctx.set_mesh(m)
    ctx.create_csr(download=False)
    print(json.dumps(measure(ctx, m, warm, steps)), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
