#!/usr/bin/env python
"""BASELINE config 5: synthetic non-orthogonal polyhedral mesh (BCC Voronoi cells, 14 faces each, jittered centres),
pressure correction with grad(pp,'gauss_corrected') seeded by a Gauss pass, `npcor` correctors and ICCG, on 1..8
GPUs (recursive-coordinate-bisection partition, src-parallel semantics).  One JSON line from rank 0.

    python tools/poly_bench.py [--n 128] [--npcor 1] [--steps 3] [--warmup 2] [--solver iccg]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/poly_bench.py --n 160

2 n^3 cells (n = 216 is the 20 M-cell size of SURVEY 8d; every rank builds the global mesh first, about 40 GB of host
memory and 2.5 minutes at that size, so the default is smaller).  With several ranks and npcor > 1 the reference's
non-orthogonal corrector system is inconsistent (each rank evaluates fluxmc of a shared face from its own side), so
the default is npcor = 1.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--npcor", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--solver", default="iccg", choices=["dpcg", "iccg", "bicgstab"])
    ap.add_argument("--sor", type=float, default=1e-8)
    ap.add_argument("--nsw", type=int, default=2000)
    args = ap.parse_args()
    import torch
    from freecappuccino_b200 import cases, lib, mesh as M, parallel
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.perf_counter()
    g = cases.poly_case(args.n)
    f = cases.flow_fields(g)
    fmi, flomas = cases.inlet_fluxes(g, f)
    t_mesh = time.perf_counter() - t0
    ctx = lib.Context(local)
    if world > 1:
        parallel.init_comm(ctx)
        mesh = M.partition(g, M.rcb_ranks(g, world), world, only=rank)[0]
        fl = {k: M.scatter_total(g, mesh, f[k]) for k in ("u", "v", "w", "p", "den")}
        fl.update({k: M.scatter_cells(g, mesh, f[k]) for k in ("apu", "apv", "apw")})
        c = mesh.count("inlet")
        gf = mesh.face_global[mesh.faces_start("inlet"):mesh.faces_start("inlet") + c]
        fmi_l = np.ascontiguousarray(fmi[gf - g.faces_start("inlet")]) if c else np.zeros(0)
    else:
        mesh, fl, fmi_l = g, f, fmi
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    p2p = parallel.enable_p2p(ctx) if world > 1 else False
    for k, name in (("u", "USER0"), ("v", "USER1"), ("w", "USER2"), ("p", "USER3"), ("den", "DEN"), ("apu", "APU"),
                    ("apv", "APV"), ("apw", "APW")):
        ctx.upload(name, fl[k])
    if fmi_l.size:
        ctx.upload("FMI", fmi_l)
    ctx.upload("P", fl["p"])
    ctx.grad_gauss("P", "DPDXI", 1)
    opts = lib.calcp_opts(solver=args.solver, flomas=flomas, npcor=args.npcor, lsq_flag=True, sor=args.sor, nsw=args.nsw,
                          pRefCell=3, flux_variant=1, parallel=world > 1)

    def step():
        for s, d in (("USER0", "U"), ("USER1", "V"), ("USER2", "W"), ("USER3", "P")):
            ctx.copy(s, d)
        return ctx.calcp(opts)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        ctx.synchronize()

    for _ in range(args.warmup):
        rep = step()
    barrier()
    iters, solve_ms, asm_ms, corr_ms = 0, 0.0, 0.0, 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rep = step()
        t = ctx.timings()
        iters += sum(rep.rep[k].iters for k in range(args.npcor))
        solve_ms += t.solve_ms; asm_ms += t.assemble_ms; corr_ms += t.correct_ms
    barrier()
    wall = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([wall, solve_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall, solve_ms = float(tt[0]), float(tt[1])
    if rank == 0:
        n, nnz = g.numCells, g.nnz
        per_iter = {"dpcg": 12 * nnz + 116 * n, "iccg": 24 * nnz + 164 * n, "bicgstab": 2 * (24 * nnz + 164 * n)}[args.solver]
        ms_it = solve_ms / max(iters, 1)
        print(json.dumps({
            "workload": f"config 5: BCC-Voronoi polyhedral mesh 2*{args.n}^3, calcp with gauss_corrected + {args.solver}",
            "cells": n, "inner_faces": g.numInnerFaces, "nnz": nnz, "n_gpus": world,
            "partition": "1 rank" if world == 1 else f"rcb, {mesh.npro} processor faces on rank 0",
            "comm": "none" if world == 1 else ("p2p" if p2p else "nccl"), "npcor": args.npcor,
            "iterations_per_step": iters / args.steps, "iter_per_s": iters / wall, "ms_per_iteration": ms_it,
            "algorithmic_gbs_per_iteration": per_iter / ms_it / 1e6,
            "ms_per_step": {"assemble": asm_ms / args.steps, "solve": solve_ms / args.steps, "correct": corr_ms / args.steps,
                            "wall": 1e3 * wall / args.steps},
            "res0": rep.rep[0].res0, "resl": rep.rep[0].resl, "mesh_build_s": t_mesh}), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
