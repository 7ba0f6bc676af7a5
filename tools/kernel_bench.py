#!/usr/bin/env python
"""Per-operation roofline table of the path on one GPU: every C-ABI operation that is (mostly) one kernel is timed
with CUDA events on the library's stream (torch ExternalStream) over `reps` back-to-back calls, and its achieved
GB/s is reported against the algorithmic byte count of DESIGN.md 4 and the measured HBM peak.

    python tools/kernel_bench.py [n] [reps]         # default 216 (10.1 M cells), 20 repetitions

Covers: grad_gauss pass, bpres, the calcp assembly (gradients + face kernel + row gather), the momentum predictor's
explicit part, the least-squares gradients and limiters, SpMV, and the per-iteration cost of dpcg / iccg / bicgstab
(solve_ms / iterations), and the same with the experimental sweep schedules (FC_TUNE_SWEEP_TILED, FC_TUNE_SWEEP_P2P).
Meant as the first gpurun call of a round: one JSON line per operation.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freecappuccino_b200 import cases, lib  # noqa: E402


def peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d.get("hbm_gbs")), "measured"
    except Exception:
        return 6650.0, "fallback"


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    m = cases.hex_case(n, n, n)
    f = cases.config4_fields(m)
    nc, F, B, nnz, nt = m.numCells, m.numInnerFaces, m.numFaces - m.numInnerFaces, m.nnz, m.numTotal
    ctx = lib.Context(0)
    ctx.set_mesh(m)
    ctx.create_csr(download=False)
    for k, name in (("u", "U"), ("v", "V"), ("w", "W"), ("p", "P"), ("den", "DEN"), ("apu", "APU"), ("apv", "APV"),
                    ("apw", "APW")):
        ctx.upload(name, f[k])
    ctx.upload("VIS", np.full(nt, 0.01))
    ctx.upload("USER0", f["p"])
    stream = torch.cuda.ExternalStream(ctx.lib.fc_stream(ctx.h), device=torch.device("cuda", 0))
    pk, src = peak()

    def timed(fn, r=reps):
        fn()                                  # warm-up
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(r):
            fn()
        e1.record(stream)
        ctx.synchronize()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / r

    def report(op, ms, nbytes, **kw):
        gbs = nbytes / ms / 1e6
        print(json.dumps(dict(op=op, n=n, cells=nc, ms=ms, algorithmic_mb=nbytes / 1e6, gbs=gbs, frac_of_peak=gbs / pk,
                              peak_gbs=pk, peak_source=src, **kw)), flush=True)

    grad_b = 64 * F + 64 * nc + 36 * B
    report("grad_gauss (1 pass)", timed(lambda: ctx.grad_gauss("USER0", "DPDXI", 1)), grad_b)
    report("bpres stage 2", timed(lambda: ctx.bpres("USER0", 2)), 76 * B)
    report("spmv", ctx.time_spmv("USER0", "SCRATCH_T", 50), 12 * nnz + 20 * nc)
    po = lib.calcp_opts(solver="dpcg", const_mflux=True, sor=1e-8, nsw=100000)
    ctx.set_tuning(lib.TUNE_FUSED_GRAD, 0)
    report("calcp assembly (3 grads + faces + rows)", timed(lambda: ctx.calcp_assemble(po), 5), 3 * grad_b + 96 * F + 208 * nc + 16 * F)
    uo = lib.calcuvw_opts(scheme="muscl-f", bdf=True, timestep=1e-2)
    ctx.fill("FLMASS", 0.0)
    report("calcuvw explicit part (5 grads + bpres + faces + rows)", timed(lambda: ctx.calcuvw_assemble(uo), 5),
           5 * grad_b + (120 * F + 160 * nc) + (184 * F + 88 * nc))
    ctx.set_tuning(lib.TUNE_FUSED_GRAD, 1)
    report("calcp assembly, fused u/v/w gradient", timed(lambda: ctx.calcp_assemble(po), 5),
           3 * grad_b + 96 * F + 208 * nc + 16 * F)
    report("calcuvw explicit part, fused u/v/w gradient", timed(lambda: ctx.calcuvw_assemble(uo), 5),
           5 * grad_b + (120 * F + 160 * nc) + (184 * F + 88 * nc))
    for method, nbytes in (("lstsq", 44 * (2 * F + B) + 96 * nc), ("lstsq_dm", 44 * (2 * F + B) + 96 * nc),
                           ("lstsq_qr", 12 * (2 * F + B) + 168 * nc)):
        try:
            ctx.set_gradient(method, "no-limit")
            report(f"grad {method}", timed(lambda: ctx.grad("USER0", "DPDXI", 1)), nbytes)
        except lib.FcError as e:
            print(json.dumps(dict(op=f"grad {method}", error=str(e))), flush=True)
    for limiter in ("Barth-Jespersen", "Venkatakrishnan", "mVenkatakrishnan"):
        ctx.set_gradient("gauss", limiter)
        t_both = timed(lambda: ctx.grad("USER0", "DPDXI", 1))
        ctx.set_gradient("gauss", "no-limit")
        t_plain = timed(lambda: ctx.grad("USER0", "DPDXI", 1))
        report(f"limiter {limiter}", max(t_both - t_plain, 1e-6), 36 * nnz + 56 * nc)
    ctx.set_gradient("gauss", "no-limit")
    # Krylov solvers: per-iteration cost on the assembled p' system (level-scheduled sweeps first)
    ctx.calcp_assemble(po)
    ctx.set_tuning(lib.TUNE_SWEEP_TILED, 0)
    for solver, nbytes in (("dpcg", 12 * nnz + 116 * nc), ("iccg", 24 * nnz + 164 * nc), ("bicgstab", 2 * (24 * nnz + 164 * nc))):
        its = 200 if solver == "dpcg" else 20
        for _ in range(2):                     # first run builds the level schedules
            ctx.fill("PP", 0.0)
            rep = ctx.solve(solver, "PP", lib.solver_opts(1e-30, its))
        t = ctx.timings()
        report(f"{solver} iteration", t.solve_ms / max(rep.iters, 1), nbytes, iters=rep.iters, solve_ms=t.solve_ms)
    # experimental sweep schedules (off by default), last so that a failure cannot take the lines above with it
    ctx.set_tuning(lib.TUNE_SWEEP_TILED, 0)
    only = os.environ.get("KB_SWEEPS")       # e.g. KB_SWEEPS=3 : only the value-as-flag mode
    variants = ((lib.TUNE_SWEEP_TILED, 1, "tiled sweeps"), (lib.TUNE_SWEEP_TILED, 2, "tiled sweeps + p2p flags"),
                (lib.TUNE_SWEEP_TILED, 3, "tiled sweeps, value-as-flag"), (lib.TUNE_SWEEP_P2P, 1, "p2p sweeps"))
    if only:
        variants = tuple(v for v in variants if v[0] == lib.TUNE_SWEEP_TILED and str(v[1]) in only.split(","))
    for key, value, label in variants:
        try:
            ctx.set_tuning(key, value)
            for solver, nbytes in (("iccg", 24 * nnz + 164 * nc), ("bicgstab", 2 * (24 * nnz + 164 * nc))):
                for _ in range(2):
                    ctx.fill("PP", 0.0)
                    rep = ctx.solve(solver, "PP", lib.solver_opts(1e-30, 20))
                t = ctx.timings()
                report(f"{solver} iteration, {label}", t.solve_ms / max(rep.iters, 1), nbytes, iters=rep.iters,
                       solve_ms=t.solve_ms, sweep_tiles=t.sweep_tiles, schedule=ctx.sweep_schedule_info())
        except lib.FcError as e:
            print(json.dumps(dict(op=label, error=str(e))), flush=True)
        ctx.set_tuning(key, 0)
    ctx.close()


if __name__ == "__main__":
    main()
