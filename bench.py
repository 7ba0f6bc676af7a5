#!/usr/bin/env python
"""Benchmark of the pressure-correction hot path (BASELINE.json: "pCorr solve: DPCG iter/s &
SpMV HBM GB/s at 10M cells; SIMPLE iter time").

A *step* is one `calcp` call -- assembly of the p' system, DPCG solve to rsm < 1e-8 and the
flux / velocity / pressure correction -- on the synthetic 216^3 hex pressure-correction case
(SURVEY.md 8d config 4, 10 077 696 cells).  `value` = DPCG iterations per second of whole
steps with all inputs resident in HBM; `e2e` = the same through fc_calcp_host with pinned HOST
buffers (H2D of u,v,w,p,apu,apv,apw and D2H of u,v,w,p,pp,flmass inside the timed region);
`roofline` is the SpMV(+p.Ap) kernel of the Krylov loop, timed with CUDA events on the
library's stream inside the timed steps.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 216]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SOR, NSW = 1e-8, 100000


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            d = json.load(fh)
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """DRAM bytes of the dominant kernel from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as fh:
            return json.load(fh)
    except Exception:
        return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.f:
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, val in zip(names, c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def build_case(n: int):
    from freecappuccino_b200 import cases, mesh as M
    m = M.hex_mesh(n, n, n)
    return m, cases.config4_fields(m)


def workload_config(n, mesh):
    """The `config` object: identical in both arms (the driver compares them textually)."""
    return {"workload": f"synthetic {n}^3 hex pressure-correction (config 4): calcp = assemble + DPCG to rsm<1e-8 "
                        f"+ correct", "cells": int(mesh.numCells), "nnz": int(mesh.nnz), "solver": "dpcg", "sor": SOR,
            "l2": ("inputs_exceed_l2" if 12 * mesh.nnz + 20 * mesh.numCells > 126e6 else "inputs_fit_l2") +
                  " (SpMV working set %.0f MB per iteration, L2 = 126 MB)" % ((12 * mesh.nnz + 20 * mesh.numCells) / 1e6)}


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def oracle_case(gmesh, gf, R):
    """The oracle's assembled config-4 system on the R-rank z-slab partition (R = 1: serial `src` semantics,
    R > 1: `src-parallel` semantics in lock step).  Returns a dict with a `solve(nsw, sor)` closure that starts from
    fi = 0 and returns (iterations, res0, resl) and `pp()` -> global cell values of the last solution."""
    from oracle import oracle, oracle_par
    from freecappuccino_b200 import mesh as M
    oo = oracle.calcp_opts(solver="dpcg", const_mflux=True, sor=SOR, nsw=NSW)
    t0 = time.perf_counter()
    if R == 1:
        csr = oracle.create_csr(gmesh)
        of = oracle.Fields(gmesh, csr.nnz)
        for k in ("u", "v", "w", "p", "den", "apu", "apv", "apw"):
            getattr(of, k)[:] = gf[k]
        of.dPdxi[:] = oracle.grad_gauss(gmesh, of.p, 1)
        t0 = time.perf_counter()
        oracle.calcp_assemble(gmesh, csr, of, oo)
        asm_s = time.perf_counter() - t0
        fis = [np.zeros(gmesh.numTotal)]

        def solve(nsw, sor=1e-30):
            fis[0][:] = 0.0
            res0, resl, used, _ = oracle.solve("dpcg", csr, of.a, of.su, fis[0], sor=sor, nsw=nsw)
            return used, res0, resl
        return dict(solve=solve, pp=lambda: fis[0][:gmesh.numCells].copy(), assemble_s=asm_s, su=[of.su], a=[of.a],
                    parts=[gmesh], keep=(csr, of))
    parts = M.partition(gmesh, M.slab_ranks(gmesh.numCells, R), R)
    pc = oracle_par.ParCase(parts)
    for r, part in enumerate(parts):
        fr = pc.fields[r]
        for k in ("u", "v", "w", "p", "den"):
            getattr(fr, k)[:] = M.scatter_total(gmesh, part, gf[k])
        for k in ("apu", "apv", "apw"):
            getattr(fr, k)[:] = M.scatter_cells(gmesh, part, gf[k])
    for r, g in enumerate(pc.grad_gauss([fr.p for fr in pc.fields], 1)):
        pc.fields[r].dPdxi[:] = g
    po = oracle.calcp_opts(solver="dpcg", const_mflux=True, sor=SOR, nsw=NSW)
    po.sol.parallel = 1
    t0 = time.perf_counter()
    pc.calcp_assemble(po)
    asm_s = time.perf_counter() - t0
    fis = [np.zeros(part.numTotal) for part in parts]

    def solve(nsw, sor=1e-30):
        for x in fis:
            x[:] = 0.0
        rep = pc.solve("dpcg", fis, sor=sor, nsw=nsw)
        return rep.iters, rep.res0, rep.resl
    return dict(solve=solve, pp=lambda: M.gather_cells(gmesh, parts, [x[:pt.numCells] for x, pt in zip(fis, parts)]),
                assemble_s=asm_s, su=[fr.su for fr in pc.fields], a=[fr.a for fr in pc.fields], parts=parts, keep=pc)


def algorithmic_bytes(mesh):
    n, nnz = mesh.numCells, mesh.nnz
    return {"spmv": 12 * nnz + 20 * n, "dpcg_iter": 12 * nnz + 116 * n}


def cpu_port_sample(mesh, a, su, iters: int):
    """Oracle DPCG (serial src semantics, 1 core) on the same system: `iters` iterations."""
    from oracle import oracle
    csr = oracle.create_csr(mesh)
    fi = np.zeros(mesh.numTotal)
    t0 = time.perf_counter()
    res0, resl, used, _ = oracle.solve("dpcg", csr, a, su, fi, sor=1e-30, nsw=iters)
    dt = time.perf_counter() - t0
    return used / dt, dt, used


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm for the path, on the box's host cores.  The reference is
    Fortran-only and cannot be compiled in this image (no Fortran compiler, no MPI), so both of its builds are
    timed through the oracle port (kind "port"):
      * `src-parallel` semantics with R = min(--ref-ranks [32], host cores) ranks, one host thread per rank (the oracle's
        lock-step multi-rank solver with OpenMP over the ranks, z-slab partition) -- the line's `value`,
        "all the host threads it can use";
      * serial `src` semantics on one core -- reported beside it as `serial`.
    After the timed steps (untimed): one whole R-rank solve to rsm < 1e-8 (`iters_to_tol`, `calcp_s`), which also
    fills BASELINE.md's "calcp call time" column for the CPU builds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_par
    mesh, f = build_case(args.n)
    per_step = args.ref_iters
    cores = host_cores()

    # ---- serial src build, one core ----
    oracle_par.set_threads2(1, 1)
    ser = oracle_case(mesh, f, 1)
    ser["solve"](2)
    t0 = time.perf_counter()
    used, _, _ = ser["solve"](per_step)
    serial_v = used / (time.perf_counter() - t0)
    serial_asm_s = ser["assemble_s"]
    del ser

    # ---- src-parallel build, R ranks on R host threads ----
    R = max(1, min(args.ref_ranks, cores))
    threads = oracle_par.set_threads(R) if R > 1 else 1
    if threads > 1:
        par = oracle_case(mesh, f, R)
        kind_txt = (f"src-parallel semantics, {R} ranks (z-slabs) on {threads} host threads of {cores} "
                    f"(OpenMP over the ranks of the lock-step oracle)")
    else:
        R = threads = 1
        par = oracle_case(mesh, f, 1)
        kind_txt = f"serial src semantics, 1 of {cores} host cores (the oracle was built without OpenMP)"
    solve = par["solve"]
    for _ in range(args.warmup):
        solve(2)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        done += solve(per_step)[0]
    dt = time.perf_counter() - t0
    v = done / dt
    # ---- untimed: the whole solve once, to the tolerance of the GPU arm ----
    to_tol = None
    if not args.no_tol_solve:
        t0 = time.perf_counter()
        it, res0, resl = solve(NSW, SOR)
        sol_s = time.perf_counter() - t0
        to_tol = {"ranks": R, "threads": threads, "iters_to_tol": int(it), "res0": res0, "resl": resl,
                  "solve_s": sol_s, "assemble_s": par["assemble_s"],
                  "calcp_s_parallel_build": par["assemble_s"] + sol_s,
                  "calcp_s_serial_build_estimate": serial_asm_s + it / serial_v,
                  "note": "calcp = assemble + DPCG to rsm<1e-8 (the correction sweeps are < 1 % and not timed); the "
                          "serial figure is its measured assembly + iters_to_tol / measured serial iter/s"}
    sample = (f"{per_step} DPCG iterations per step on the {args.n}^3 p' system; {kind_txt}; serial src build on 1 core: "
              f"{serial_v:.2f} iter/s (oracle assembly {serial_asm_s:.1f} s, untimed)")
    print(json.dumps({
        "impl": "reference", "metric": "pcorr_dpcg_iterations_per_second", "value": v, "unit": "iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.n, mesh),
        "cpu_baseline": {"value": v, "unit": "iter/s", "cores": threads, "kind": "port", "sample": sample},
        "serial": {"value": serial_v, "unit": "iter/s", "cores": 1},
        "e2e": {"value": v, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_assemble_s": serial_asm_s,
        "to_tolerance": to_tol,
    }))


def run_tool(cmd, timeout):
    """Run a tools/ script in its own process and return its last JSON line (or an error object)."""
    try:
        r = subprocess.run([sys.executable] + [os.path.join(ROOT, cmd[0])] + cmd[1:], capture_output=True, text=True,
                           timeout=timeout)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"error": f"exit {r.returncode}: {r.stderr[-300:]}"}
    except Exception as e:
        return {"error": repr(e)}


def parity_check(args, ctx, opts, step, gmesh, gf, mesh, rank, world):
    """Outside every timed region: one more device-resident step, then rank 0 runs the oracle (test infrastructure,
    here only as the checker) on the SAME z-slab partition -- assembly and a DPCG solve to the same tolerance, with all
    host threads (ranks x row loops; sums stay sequential, so the oracle's results do not depend on the thread
    count) -- and compares iteration counts (north star: within +-1), the right-hand side (bit for bit), rank 0's
    matrix coefficients (bit for bit) and the solved pressure correction (relative L2 over all cells)."""
    import torch
    from freecappuccino_b200 import mesh as M
    rep = step()
    n = mesh.numCells
    pp = np.ascontiguousarray(ctx.download("PP")[:n])
    su = np.ascontiguousarray(ctx.download("SU")[:n])
    if world > 1:
        import torch.distributed as dist
        pps, sus = [None] * world, [None] * world
        dist.all_gather_object(pps, pp)
        dist.all_gather_object(sus, su)
    else:
        pps, sus = [pp], [su]
    res = None
    if rank == 0:
        from oracle import oracle_par
        t0 = time.perf_counter()
        cores = host_cores()
        outer, inner = oracle_par.set_threads2(world, max(1, cores // world))
        oc = oracle_case(gmesh, gf, world)
        it, res0, resl = oc["solve"](NSW, SOR)
        pp_o = oc["pp"]()
        pp_g = M.gather_cells(gmesh, oc["parts"], pps) if world > 1 else pps[0]
        a0 = ctx.download("A")
        res = {"iters_gpu": int(rep.rep[0].iters), "iters_oracle_same_partition": int(it),
               "iters_within_1": bool(abs(int(rep.rep[0].iters) - int(it)) <= 1),
               "res0_gpu": rep.rep[0].res0, "res0_oracle": res0,
               "rel_l2_pp": float(np.linalg.norm(pp_g - pp_o) / np.linalg.norm(pp_o)),
               "su_bit_identical_all_ranks": bool(all(np.array_equal(g, o[:g.size]) for g, o in zip(sus, oc["su"]))),
               "a_bit_identical_rank0": bool(np.array_equal(a0, oc["a"][0][:a0.size])),
               "oracle": f"{'serial src' if world == 1 else 'src-parallel'} semantics, {world} rank(s) x {inner} "
                         f"thread(s) of {cores} host cores, {time.perf_counter() - t0:.1f} s",
               "tolerance": {"iterations": 1, "rel_l2_pp": "reported; both solves stop at rsm < 1e-8"}}
        del oc
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    return res


def run_ours(args):
    import torch
    from freecappuccino_b200 import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from freecappuccino_b200 import mesh as M, parallel
    gmesh, gf = build_case(args.n)
    ctx = lib.Context(local)
    if world > 1:
        # cell partition into contiguous z-slabs (src-parallel layout: halo slots, processor faces, apr strip)
        parallel.init_comm(ctx)
        mesh = M.partition(gmesh, M.slab_ranks(gmesh.numCells, world), world, only=rank)[0]
        f = {k: M.scatter_total(gmesh, mesh, gf[k]) for k in ("u", "v", "w", "p", "den")}
        f.update({k: M.scatter_cells(gmesh, mesh, gf[k]) for k in ("apu", "apv", "apw")})
    else:
        mesh, f = gmesh, gf
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    p2p = parallel.enable_p2p(ctx) if world > 1 else False
    for k, name in (("u", "USER0"), ("v", "USER1"), ("w", "USER2"), ("p", "USER3"), ("den", "DEN"), ("apu", "APU"),
                    ("apv", "APV"), ("apw", "APW")):
        ctx.upload(name, f[k])
    ctx.upload("P", f["p"])
    ctx.grad_gauss("P", "DPDXI", 1)          # incoming pressure gradient of the SIMPLE iteration
    opts = lib.calcp_opts(solver="dpcg", const_mflux=True, sor=SOR, nsw=NSW, pRefCell=1, urf_p=0.3,
                          parallel=world > 1)
    ctx.set_spmv_sampling(512)
    if args.no_persist:
        ctx.set_tuning(lib.TUNE_DPCG_PERSISTENT, 0)
    if args.ctas_per_sm > 0:      # A/B knobs of the persistent kernel (library defaults otherwise)
        ctx.set_tuning(lib.TUNE_CTAS_PER_SM, args.ctas_per_sm)
    if args.pipe >= 0:
        ctx.set_tuning(lib.TUNE_PIPE_GEOMETRY, args.pipe)
    stream = torch.cuda.ExternalStream(ctx.lib.fc_stream(ctx.h), device=torch.device("cuda", local))

    def restore():  # device-to-device: the step always starts from the same fields
        for s, d in (("USER0", "U"), ("USER1", "V"), ("USER2", "W"), ("USER3", "P")):
            ctx.copy(s, d)
        # calcp leaves grad(pp) in DPDXI (calcp:132-143); the incoming pressure gradient of the SIMPLE iteration --
        # what calcuvw would hand over -- is recomputed so that EVERY step assembles the documented config-4 system
        ctx.grad_gauss("P", "DPDXI", 1)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        ctx.synchronize()
        torch.cuda.synchronize()

    def step():
        restore()
        return ctx.calcp(opts)

    for _ in range(args.warmup):
        rep = step()
    barrier()
    clocks = ClockSampler(local)
    l0 = ctx.timings().launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 0
    iters_list = []
    spmv_ms, spmv_n, asm_ms, corr_ms, solve_ms = 0.0, 0, 0.0, 0.0, 0.0
    persist = dict(ms=0.0, pupdate_ms=0.0, spmv_ms=0.0, update_ms=0.0, mail_ms=0.0, iters=0, grid=0)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rep = step()
        iters += rep.rep[0].iters
        iters_list.append(int(rep.rep[0].iters))
        t = ctx.timings()
        spmv_ms += t.spmv_ms * t.spmv_samples
        spmv_n += t.spmv_samples
        asm_ms += t.assemble_ms; corr_ms += t.correct_ms; solve_ms += t.solve_ms
        if t.persist_iters:
            persist["ms"] += t.persist_ms; persist["pupdate_ms"] += t.persist_pupdate_ms
            persist["spmv_ms"] += t.persist_spmv_ms; persist["update_ms"] += t.persist_update_ms
            persist["mail_ms"] += t.persist_mail_ms; persist["iters"] += t.persist_iters
            persist["grid"] = t.persist_grid
            persist["index_bytes"] = t.persist_index_bytes
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.timings().launches - l0
    clk = clocks.stop()
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms = float(tt.item())
    value = iters / (dev_ms * 1e-3)

    # ---- e2e: host buffers through fc_calcp_host, copies inside the timed region ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hb = {k: pin(f[k]) for k in ("u", "v", "w", "p", "apu", "apv", "apw")}
    hpp = torch.zeros(mesh.numTotal, dtype=torch.float64).pin_memory()
    hfl = torch.zeros(mesh.numInnerFaces, dtype=torch.float64).pin_memory()
    src = {k: torch.from_numpy(np.ascontiguousarray(f[k])) for k in ("u", "v", "w", "p")}
    e2e_steps = max(1, args.steps)
    h2d = 8 * (4 * mesh.numTotal + 3 * (mesh.numCells + mesh.npro))
    d2h = 8 * (5 * mesh.numTotal + mesh.numInnerFaces)

    def e2e_step():
        # untimed: fresh host inputs (fc_calcp_host returns the corrected u, v, w, p in place, like `call calcp`) and
        # the incoming pressure gradient of the SIMPLE iteration, which stays device-resident
        for k in ("u", "v", "w", "p"):
            hb[k].copy_(src[k])
        ctx.grad_gauss("USER3", "DPDXI", 1)
        barrier()
        # timed: H2D of u,v,w,p,ap* from pinned memory + calcp + D2H of u,v,w,p,pp,flmass (the call returns after
        # its last copy has completed)
        t0 = time.perf_counter()
        r = ctx.calcp_host(opts, hb["u"].numpy(), hb["v"].numpy(), hb["w"].numpy(), hb["p"].numpy(), hpp.numpy(),
                           hb["apu"].numpy(), hb["apv"].numpy(), hb["apw"].numpy(), hfl.numpy())
        return r, time.perf_counter() - t0
    e2e_step()
    e2e_iters, e2e_s = 0, 0.0
    for _ in range(e2e_steps):
        r, dt = e2e_step()
        if world > 1:   # a step is over when its slowest rank has its results on the host
            import torch.distributed as dist
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e_iters += r.rep[0].iters
        e2e_s += dt
    e2e_value = e2e_iters / e2e_s

    # ---- roofline of the dominant kernel ----
    # Default path: the whole Krylov loop of a step is ONE launch of the persistent kernel k_dpcg_persist, so the
    # launch duration is the solve bracket (CUDA events on the library's stream) and the launch processes
    # `iterations` units of SURVEY 8(d)'s per-iteration figure 12 nnz + 116 n.  Its SpMV phase is timed inside the
    # kernel on the GPU's global timer (grid barrier release -> arrival of the last CTA at the next barrier).
    # With FC_TUNE persistent=0 (or NCCL mode) the SpMV(+p.Ap) launches are bracketed with events instead.
    ab = algorithmic_bytes(mesh)   # this rank's share
    peak, peak_src = measured_peak()
    iters_per_step = iters / args.steps
    alone_ms = ctx.time_spmv("PP", "SCRATCH_T", 50)
    traffic = measured_traffic()
    if persist["iters"]:
        launch_ms = solve_ms / args.steps
        achieved = ab["dpcg_iter"] * iters_per_step / (launch_ms * 1e-3) / 1e9
        it = persist["iters"]
        spmv_us = 1e3 * persist["spmv_ms"] / it
        ib = persist.get("index_bytes", 4) or 4
        # what the kernel as written moves per iteration: (8 + index bytes) per non-zero, 100 per row
        moved = (8 + ib) * mesh.nnz + 100 * mesh.numCells
        roof = {"bound": "hbm", "kernel": "k_dpcg_persist<256,2304,2> (whole DPCG solve, one cooperative launch per step; column "
                                          + ("indices as one-byte codes)" if ib == 1 else "indices as int32)"),
                "index_bytes_per_nonzero": ib, "bytes_moved_per_iteration": moved,
                "note": ("achieved / frac use SURVEY 8(d)'s unit 12 nnz + 116 n per iteration; with one-byte column codes "
                         "the kernel moves (8 + 1) nnz + 100 n (ncu: profiles/ncu_traffic.json), so frac can exceed 1 -- "
                         "moved_gbs / moved_frac are the same launch time on the bytes actually moved") if ib == 1 else
                        "achieved / frac use SURVEY 8(d)'s unit 12 nnz + 116 n per iteration; the kernel moves 12 nnz + 100 n",
                "moved_gbs": moved * iters_per_step / (launch_ms * 1e-3) / 1e9,
                "moved_frac": moved * iters_per_step / (launch_ms * 1e-3) / 1e9 / peak,
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8000": achieved / 8000.0,
                "traffic": (traffic["dram_bytes_per_iteration"] * iters_per_step)
                if traffic and world == 1 and args.n == traffic.get("n", 216) else None,
                "traffic_source": traffic["source"] if traffic and world == 1 and args.n == traffic.get("n", 216) else None,
                "algorithmic_bytes_per_launch": ab["dpcg_iter"] * iters_per_step,
                "algorithmic_bytes_per_iteration": ab["dpcg_iter"], "iterations_per_launch": iters_per_step,
                "mean_launch_ms": launch_ms, "launches_timed": args.steps, "grid": persist["grid"],
                "phases_us_per_iteration": {
                    "p_update": 1e3 * persist["pupdate_ms"] / it, "spmv_dot": spmv_us,
                    "x_r_update": 1e3 * persist["update_ms"] / it,
                    "barriers_reductions_halo": 1e3 * (persist["ms"] - persist["pupdate_ms"] - persist["spmv_ms"]
                                                       - persist["update_ms"]) / it,
                    "of_which_waiting_for_other_ranks": 1e3 * persist["mail_ms"] / it},
                "spmv_phase": {"algorithmic_bytes": ab["spmv"], "us": spmv_us,
                               "achieved": ab["spmv"] / (spmv_us * 1e-6) / 1e9,
                               "frac": ab["spmv"] / (spmv_us * 1e-6) / 1e9 / peak,
                               "frac_of_nominal_8000": ab["spmv"] / (spmv_us * 1e-6) / 1e9 / 8000.0,
                               "clock": "in-kernel %globaltimer, barrier release -> last CTA arrival"},
                "standalone_spmv_ms": alone_ms, "standalone_spmv_gbs": ab["spmv"] / (alone_ms * 1e-3) / 1e9}
    else:
        spmv_mean_ms = spmv_ms / max(spmv_n, 1)
        achieved = ab["spmv"] / (spmv_mean_ms * 1e-3) / 1e9 if spmv_n else None
        roof = {"bound": "hbm", "kernel": "k_spmv<256,2304,DOT> (SpMV + p.Ap)", "achieved": achieved, "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "frac_of_nominal_8000": (achieved / 8000.0) if achieved else None, "traffic": None,
                "algorithmic_bytes_per_launch": ab["spmv"], "mean_launch_ms": spmv_mean_ms,
                "launches_sampled": spmv_n, "standalone_spmv_ms": alone_ms,
                "standalone_spmv_gbs": ab["spmv"] / (alone_ms * 1e-3) / 1e9}
    iter_ms = solve_ms / max(iters, 1)
    out = {
        "metric": "pcorr_dpcg_iterations_per_second", "value": value, "unit": "iter/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.n, gmesh),
        "layout": {"partition": "1 rank" if world == 1 else f"{world} z-slabs, {mesh.npro} processor faces on rank 0",
                   "comm": "none" if world == 1 else ("p2p: NVLink stores + flags inside the persistent kernel (CUDA IPC)"
                                                      if p2p else "nccl")},
        "dpcg_iterations_per_step": iters / args.steps,
        "dpcg_iterations_each_step": iters_list,
        "simple_iter_ms": {"assemble": asm_ms / args.steps, "solve": solve_ms / args.steps,
                           "correct": corr_ms / args.steps},
        "dpcg_iter_ms": iter_ms,
        "dpcg_iter_gbs": ab["dpcg_iter"] / (iter_ms * 1e-3) / 1e9 if iters else None,
        "wall_s": wall,
        "e2e": {"value": e2e_value, "unit": "iter/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
                "timed": "per step: fc_calcp_host call (H2D + calcp + D2H, returns after the last copy), max over "
                         "ranks; the host-side refresh of the input buffers between steps is not timed"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof,
    }
    # ---- the other two Krylov solvers of the path on the same system and partition (src/iccg.f90, src/bicgstab.f90;
    #      the shipped parallel calcp solves p' with bicgstab, the serial one with iccg): per-iteration cost of a short
    #      run, outside the headline's timed region ----
    out["other_solvers"] = {}
    ctx.calcp_assemble(opts)
    for solver, its, per_iter in (("iccg", 30, 24 * gmesh.nnz + 164 * gmesh.numCells),
                                  ("bicgstab", 15, 2 * (24 * gmesh.nnz + 164 * gmesh.numCells))):
        try:
            best = None
            for _ in range(3):   # the first run builds the sweep schedules
                ctx.fill("PP", 0.0)
                r = ctx.solve(solver, "PP", lib.solver_opts(1e-30, its, parallel=world > 1))
                ms = ctx.timings().solve_ms
                if world > 1:
                    import torch.distributed as dist
                    tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    ms = float(tt.item())
                best = ms if best is None else min(best, ms)
            ms_it = best / max(r.iters, 1)
            out["other_solvers"][solver] = {
                "iterations_timed": int(r.iters), "ms_per_iteration": ms_it, "iter_per_s": 1e3 / ms_it,
                "algorithmic_gbs_all_gpus": per_iter / ms_it / 1e6, "frac_of_hbm_peak_per_gpu": per_iter / ms_it / 1e6 / world / peak,
                "sweeps": ctx.sweep_schedule_info()[:220]}
        except Exception as e:
            out["other_solvers"][solver] = {"error": repr(e)}
    if not args.no_parity:
        out["parity"] = parity_check(args, ctx, opts, step, gmesh, gf, mesh, rank, world)
    if rank == 0 and world == 1 and not args.no_cpu:
        a = ctx.download("A")
        su = ctx.download("SU")
        v, dt, used = cpu_port_sample(mesh, a, su, args.cpu_iters)
        out["cpu_baseline"] = {"value": v, "unit": "iter/s", "cores": 1, "kind": "port",
                               "sample": f"{used} DPCG iterations of the same {args.n}^3 system (oracle, serial src "
                                         f"semantics, {dt:.1f} s)"}
    ctx.close()
    if world == 1 and not args.no_simple:
        # one whole device-resident SIMPLE iteration (calcuvw -> calcp) on the same mesh, the "SIMPLE iter time" of
        # BASELINE.json's metric: momentum predictor + this path, no field crossing PCIe.  Run in
        # its own process (tools/simple_iter_bench.py) so that the headline line cannot depend on the widened step.
        try:
            import subprocess
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "simple_iter_bench.py"), str(args.n), "2", "3"],
                               capture_output=True, text=True, timeout=240)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            out["simple_iteration"] = json.loads(lines[-1]) if r.returncode == 0 and lines else \
                {"error": f"exit {r.returncode}: {r.stderr[-300:]}"}
        except Exception as e:
            out["simple_iteration"] = {"error": repr(e)}
    if world == 1 and not args.no_configs:
        # the other two GPU configurations of BASELINE.json, each in its own process: config 3 (100^3 Poisson, DPCG and
        # ICCG, oracle iteration counts beside them) and config 5 (20 M-cell polyhedral mesh, gauss_corrected +
        # npcor = 2 + ICCG; needs ~25 GB of host memory for the mesh, skipped with a note when the box has less)
        out["config3"] = run_tool(["tools/config3_bench.py", "100"], 300)
        try:
            import psutil
            avail = psutil.virtual_memory().available / 1e9
        except Exception:
            avail = 0.0
        if avail >= 40.0:
            out["config5"] = run_tool(["tools/poly_bench.py", "--n", str(args.poly_n), "--npcor", "2", "--steps", "1",
                                       "--warmup", "1"], 900)
        else:
            out["config5"] = {"skipped": f"{avail:.0f} GB of host memory available, the 2*{args.poly_n}^3 mesh needs ~25 GB"}
    if world > 1 and not args.no_configs and args.poly_n % world == 0:
        # config 5 on the same ranks (layer slabs generated rank by rank; npcor = 1: with several ranks the reference's
        # non-orthogonal corrector system is inconsistent, see tools/poly_bench.py)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import poly_bench
            res = poly_bench.run(poly_bench.parse(["--n", str(args.poly_n), "--npcor", "1", "--steps", "1", "--warmup", "1"]))
            if rank == 0:
                out["config5"] = res
        except Exception as e:   # the headline line must not depend on the widened run
            if rank == 0:
                out["config5"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("FC_BENCH_N", "216")),
                    help="cells per edge of the synthetic hex box (also FC_BENCH_N: torchrun's own parser trips over "
                         "an option that starts with --n)")
    ap.add_argument("--cpu-iters", type=int, default=60, help="DPCG iterations of the cpu_baseline sample")
    ap.add_argument("--ref-iters", type=int, default=20, help="DPCG iterations per step of the reference arm")
    ap.add_argument("--ref-ranks", type=int, default=32, help="reference arm: ranks (= host threads) of the "
                    "src-parallel build, capped by the host's core count")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", default=os.environ.get("FC_BENCH_NO_PARITY") == "1", help="skip the full-size oracle comparison (1-2 minutes of "
                    "host time on rank 0, outside the timed regions)")
    ap.add_argument("--no-tol-solve", action="store_true", help="reference arm: skip the untimed solve to rsm<1e-8")
    ap.add_argument("--no-configs", action="store_true", default=os.environ.get("FC_BENCH_NO_CONFIGS") == "1", help="skip the config 3 / config 5 sub-benchmarks (N = 1)")
    ap.add_argument("--poly-n", type=int, default=int(os.environ.get("FC_BENCH_POLY_N", "216")), help="config 5: 2 n^3 polyhedral cells (216 = 20.2 M)")
    ap.add_argument("--no-simple", action="store_true", help="skip the SIMPLE-iteration (calcuvw + calcp) timing")
    ap.add_argument("--ctas-per-sm", type=int, default=0, help="persistent DPCG kernel: CTAs per SM (0 = library default)")
    ap.add_argument("--pipe", type=int, default=-1, help="TMA pipeline geometry 0..3 (-1 = library default)")
    ap.add_argument("--no-persist", action="store_true", help="one launch per vector operation instead of the "
                    "persistent DPCG kernel (A/B)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
