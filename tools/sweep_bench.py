#!/usr/bin/env python
"""A/B timings on one GPU (JSON lines):
  * ICCG / BiCGStab per-iteration cost on the config-4 p' system for every triangular-sweep schedule
    (FC_TUNE_SWEEP_TILED 0..3) and tile occupancy (FC_TUNE_TILE_CTAS 2, 3), with a bit-identity check of the iterate
    against the level schedule;
  * the persistent DPCG kernel's phase times with and without the L2 evict_last marking of the Krylov vectors
    (FC_TUNE_L2_KEEP) -- at 108^3 one GPU holds what a rank of the 8-GPU 216^3 run holds.

    python tools/sweep_bench.py sweeps 216     |     python tools/sweep_bench.py l2 108 136
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freecappuccino_b200 import cases, lib  # noqa: E402


def setup(n):
    poly = os.environ.get("MESH") == "poly"   # BCC-Voronoi cells (config 5's cell type), 2 n^3 of them
    m = cases.poly_case(n) if poly else cases.hex_case(n, n, n)
    f = cases.config4_fields(m)
    ctx = lib.Context(0)
    ctx.set_mesh(m)
    ctx.create_csr(download=False)
    for k, name in (("u", "U"), ("v", "V"), ("w", "W"), ("p", "P"), ("den", "DEN"), ("apu", "APU"), ("apv", "APV"),
                    ("apw", "APW")):
        ctx.upload(name, f[k])
    ctx.grad_gauss("P", "DPDXI", 1)
    ctx.calcp_assemble(lib.calcp_opts(solver="dpcg", const_mflux=True, sor=1e-8, nsw=100000, flux_variant=1 if poly else 0))
    return m, ctx


def sweeps(n):
    m, ctx = setup(n)
    nnz, nc = m.nnz, m.numCells
    ref = {}
    for tiled, occ in [(int(t.split(":")[0]), int(t.split(":")[1]) if ":" in t else 2) for t in os.environ.get("SWEEP_MODES", "0,2,4,5").split(",")]:
        ctx.set_tuning(lib.TUNE_SWEEP_TILED, tiled)
        ctx.set_tuning(lib.TUNE_TILE_CTAS, occ)
        for solver, nbytes, its in (("iccg", 24 * nnz + 164 * nc, 20), ("bicgstab", 2 * (24 * nnz + 164 * nc), 10)):
            try:
                best = None
                for _ in range(3):
                    ctx.fill("PP", 0.0)
                    rep = ctx.solve(solver, "PP", lib.solver_opts(1e-30, its))
                    ms = ctx.timings().solve_ms
                    best = ms if best is None else min(best, ms)
                x = ctx.download("PP")[:nc]
                if tiled == 0:
                    ref[solver] = x
                ms_it = best / max(rep.iters, 1)
                print(json.dumps(dict(op=f"{solver} iteration", n=n, sweep_tiled=tiled, tile_ctas=occ, ms_per_iteration=ms_it,
                                      gbs=nbytes / ms_it / 1e6, iters=rep.iters, resl=rep.resl,
                                      bit_identical_to_level_schedule=bool(np.array_equal(x, ref[solver])) if solver in ref else None,
                                      schedule=ctx.sweep_schedule_info()[:160])), flush=True)
            except lib.FcError as e:
                print(json.dumps(dict(op=f"{solver} iteration", sweep_tiled=tiled, tile_ctas=occ, error=str(e))), flush=True)
    ctx.close()


def l2(n, key=None):
    m, ctx = setup(n)
    nnz, nc = m.nnz, m.numCells
    key = lib.TUNE_L2_KEEP if key is None else key
    for keep in (0, 1, 0, 1):
        ctx.set_tuning(key, keep)
        for _ in range(2):
            ctx.fill("PP", 0.0)
            rep = ctx.solve("dpcg", "PP", lib.solver_opts(1e-30, 400))
        t = ctx.timings()
        it = max(t.persist_iters, 1)
        print(json.dumps(dict(op="dpcg persistent", n=n, cells=nc, knob="l2_keep" if key == lib.TUNE_L2_KEEP else "dpcg_fused", value=keep, iters=rep.iters, us_per_iteration=1e3 * t.solve_ms / it,
                              gbs=(12 * nnz + 116 * nc) / (t.solve_ms / it) / 1e6,
                              p_update_us=1e3 * t.persist_pupdate_ms / it, spmv_us=1e3 * t.persist_spmv_ms / it,
                              update_us=1e3 * t.persist_update_ms / it,
                              sync_us=1e3 * (t.persist_ms - t.persist_pupdate_ms - t.persist_spmv_ms - t.persist_update_ms) / it,
                              resl=rep.resl)), flush=True)
    ctx.close()


def knobs(n, specs, iters=400):
    """Persistent DPCG phase times for a list of knob settings: each spec is "key=value[+key=value...]"
    (TUNE_KEYS names), e.g.  python tools/sweep_bench.py knobs 108 mat_keep=0 mat_keep=50 mat_keep=50+l2_keep=1"""
    m, ctx = setup(n)
    nnz, nc = m.nnz, m.numCells
    ref = None
    for spec in specs:
        kv = dict(item.split("=") for item in spec.split("+"))
        for k, v in kv.items():
            ctx.set_tuning(lib.TUNE_KEYS[k], int(v))
        best = None
        for _ in range(3):
            ctx.fill("PP", 0.0)
            rep = ctx.solve("dpcg", "PP", lib.solver_opts(1e-30, iters))
            t = ctx.timings()
            if best is None or t.solve_ms < best[0]:
                best = (t.solve_ms, t.persist_ms, t.persist_pupdate_ms, t.persist_spmv_ms, t.persist_update_ms, t.persist_iters,
                        t.persist_grid, t.persist_index_bytes)
        x = ctx.download("PP")[:nc]
        if ref is None:
            ref = x
        it = max(best[5], 1)
        print(json.dumps(dict(op="dpcg persistent", n=n, cells=nc, knobs=spec, iters=rep.iters, us_per_iteration=1e3 * best[0] / it,
                              gbs=(12 * nnz + 116 * nc) / (best[0] / it) / 1e6,
                              p_update_us=1e3 * best[2] / it, spmv_us=1e3 * best[3] / it, update_us=1e3 * best[4] / it,
                              sync_us=1e3 * (best[1] - best[2] - best[3] - best[4]) / it, grid=best[6], index_bytes=best[7],
                              bit_identical_to_first=bool(np.array_equal(x, ref)), resl=rep.resl)), flush=True)
    ctx.close()


def faces(n):
    """Assembly times for every register budget of the face kernels (FC_TUNE_FACE_OCC)."""
    import torch
    m, ctx = setup(n)
    nc, F, B = m.numCells, m.numInnerFaces, m.numFaces - m.numInnerFaces
    ctx.upload("VIS", np.full(m.numTotal, 0.01))
    ctx.fill("FLMASS", 0.0)
    stream = torch.cuda.ExternalStream(ctx.lib.fc_stream(ctx.h), device=torch.device("cuda", 0))
    po = lib.calcp_opts(solver="dpcg", const_mflux=True, sor=1e-8, nsw=100000)
    uo = lib.calcuvw_opts(scheme="muscl-f", bdf=True, timestep=1e-2)
    grad_b = 64 * F + 64 * nc + 36 * B

    def timed(fn, r=10):
        fn(); ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(r):
            fn()
        e1.record(stream)
        ctx.synchronize(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / r
    ref = None
    for occ in (2, 3, 4, 2, 3):
        ctx.set_tuning(lib.TUNE_FACE_OCC, occ)
        tp = timed(lambda: ctx.calcp_assemble(po))
        out = [ctx.download(k) for k in ("A", "SU", "FLMASS")]
        tu = timed(lambda: ctx.calcuvw_assemble(uo))
        out += [ctx.download(k) for k in ("A", "SU", "SV", "SW")]
        if ref is None:
            ref = out
        same = all(np.array_equal(a, b) for a, b in zip(ref, out))
        bp = 3 * grad_b + 96 * F + 208 * nc + 16 * F
        bu = 5 * grad_b + (120 * F + 160 * nc) + (184 * F + 88 * nc)
        print(json.dumps(dict(op="assembly", n=n, face_occ=occ, calcp_assemble_ms=tp, calcp_gbs=bp / tp / 1e6,
                              calcuvw_assemble_ms=tu, calcuvw_gbs=bu / tu / 1e6, bit_identical=bool(same))), flush=True)
    ctx.close()


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "knobs":
        knobs(int(sys.argv[2]), sys.argv[3:])
        sys.exit(0)
    for n in [int(x) for x in sys.argv[2:]]:
        if what == "sweeps":
            sweeps(n)
        elif what == "faces":
            faces(n)
        elif what == "fused":
            l2(n, lib.TUNE_DPCG_FUSED)
        else:
            l2(n)
