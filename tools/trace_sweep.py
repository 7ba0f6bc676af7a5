#!/usr/bin/env python
"""Measurement aid: per-tile time stamps of the point-to-point tiled forward sweep (library built with -DFC_SWEEP_TRACE)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FC_SWEEP_TRACE_FILE"] = os.path.join(ROOT, "gpurun_out", "sweep_trace.txt")
from freecappuccino_b200 import cases, lib
lib.LIB_PATH = os.environ.get("FCAPP_LIB", os.path.join(ROOT, "freecappuccino_b200", "libfcapp_cuda_trace.so"))   # the -DFC_SWEEP_TRACE build
n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
poly = len(sys.argv) > 3 and sys.argv[3] == "poly"
m = cases.poly_case(n) if poly else cases.hex_case(n, n, n)
f = cases.config4_fields(m)
ctx = lib.Context(0)
ctx.set_mesh(m); ctx.create_csr(download=False)
for k, name in (("u", "U"), ("v", "V"), ("w", "W"), ("p", "P"), ("den", "DEN"), ("apu", "APU"), ("apv", "APV"), ("apw", "APW")):
    ctx.upload(name, f[k])
ctx.grad_gauss("P", "DPDXI", 1)
ctx.calcp_assemble(lib.calcp_opts(solver="dpcg", const_mflux=True, sor=1e-8, nsw=100000, flux_variant=1 if poly else 0))
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx.set_tuning(lib.TUNE_SWEEP_TILED, mode)
ctx.fill("PP", 0.0)
rep = ctx.solve("iccg", "PP", lib.solver_opts(1e-30, 10))
print("iters", rep.iters, "solve_ms", ctx.timings().solve_ms)
t = np.loadtxt(os.environ["FC_SWEEP_TRACE_FILE"])
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = ["start->ticket", "ticket->meta", "meta->loaded", "loaded->flags", "flags->ldcg", "ldcg->walk_done", "walk_done->released"]
if mode == 4:
    names = ["ticket->metas", "metas->pass2", "pass2->flags", "flags->pass3", "pass3->levstart", "levstart->walk_done", "walk_done->released"]
if mode == 5:
    # stamps: 0 ticket, 1 staged, 2 level table, 3 level 0 starts, 4 middle level starts, 5 walk done, 6 helper of thread 64 done
    t = t[t[:, 5] > 0]
    t0 = t[:, 0].min()
    print("tiles traced", len(t), "sweep span us", (t[:, 5].max() - t0) / 1e3)
    for nm, a, b in (("ticket->staged", 0, 1), ("staged->level table", 1, 2), ("table->level 0 starts (wait)", 2, 3),
                     ("level 0->middle level", 3, 4), ("middle level->walk done", 4, 5), ("whole tile", 0, 5),
                     ("staged->first helper done", 1, 6)):
        x = t[:, b] - t[:, a]
        x = x[(t[:, b] > 0) & (t[:, a] > 0)]
        print(f"{nm:32s} median {np.median(x):8.0f} ns  mean {x.mean():8.0f}  p90 {np.percentile(x, 90):8.0f}")
    ev = np.concatenate([np.stack([t[:, 0], np.ones(len(t))], 1), np.stack([t[:, 5], -np.ones(len(t))], 1)])
    ev = ev[np.argsort(ev[:, 0])]
    print("alive tiles (x16): mean", 16 * np.cumsum(ev[:, 1]).mean(), "max", 16 * np.cumsum(ev[:, 1]).max())
    ctx.close()
    sys.exit(0)
d = np.diff(t[:, :8], axis=1)
print("tiles traced", len(t), "sweep span us", (t[:, 7].max() - t0) / 1e3)
for i, nm in enumerate(names):
    print(f"{nm:22s} median {np.median(d[:, i]):8.0f} ns  mean {d[:, i].mean():8.0f}  p90 {np.percentile(d[:, i], 90):8.0f}")
if mode != 4:
    print("first 8 levels of the walk: median", np.median(t[:, 8] - t[:, 5]), "ns ; levels", np.median(t[:, 9]))
print("whole tile (start->released): median", np.median(t[:, 7] - t[:, 0]), "mean", (t[:, 7] - t[:, 0]).mean())
# concurrency: how many traced tiles are alive at a time (x16 since every 16th is traced)
ev = np.concatenate([np.stack([t[:, 0], np.ones(len(t))], 1), np.stack([t[:, 7], -np.ones(len(t))], 1)])
ev = ev[np.argsort(ev[:, 0])]
alive = np.cumsum(ev[:, 1])
print("alive tiles (x16): mean", 16 * alive.mean(), "max", 16 * alive.max())
ctx.close()
