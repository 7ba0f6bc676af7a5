"""Opt-in checks of the kernel variants behind fc_set_tuning (the default ones included: every sweep schedule is compared
with the level schedule).  They take a minute and repeat what the default suite checks for the default kernels, so the
default `pytest -m gpu` skips them; enable with FCAPP_EXPERIMENTAL=1 (profiles/r02_pytest_experimental_final.log: 54 passed
on B200).

FC_TUNE_SWEEP_P2P: point-to-point block flags instead of one counter per level in the DIC / DILU triangular sweeps
(fc_trisolve.cu).  FC_TUNE_SWEEP_TILED: two-level schedule, spatial tiles walked inside one CTA (fc_tile_schedule.hpp;
its schedule and the kernel sources are checked on the CPU by tests/test_tile_schedule.py), hand-over by level counters,
producer flags (modes 1, 2, 4) or the values themselves (modes 3, 5).  The row sums are unchanged in both, so every
iterate must be bit-identical to the default mode."""
import os
import time

import numpy as np
import pytest

from freecappuccino_b200 import cases

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FCAPP_EXPERIMENTAL") != "1", reason="set FCAPP_EXPERIMENTAL=1")]

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MESHES = {
    "hex": lambda: cases.hex_case(40, 36, 30, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry")),
    "poly": lambda: cases.poly_case(10),
    "pitzDaily": lambda: cases.golden_mesh(os.path.join(GOLD, "pitzDaily.npz")),
    "slab": lambda: cases.hex_case(60, 60, 1, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry")),
}


@pytest.fixture(scope="module")
def fc():
    from freecappuccino_b200 import lib
    return lib


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("solver", ["iccg", "bicgstab"])
@pytest.mark.parametrize("mode", ["p2p", "tiled", "tiled-p2p", "tiled-vf", "tiled-walk", "tiled-walk-vf"])
def test_sweep_variants_are_bit_identical_to_level_sweeps(fc, name, solver, mode):
    mesh = MESHES[name]()
    key, value = {"p2p": (fc.TUNE_SWEEP_P2P, 1), "tiled": (fc.TUNE_SWEEP_TILED, 1), "tiled-p2p": (fc.TUNE_SWEEP_TILED, 2),
                  "tiled-vf": (fc.TUNE_SWEEP_TILED, 3), "tiled-walk": (fc.TUNE_SWEEP_TILED, 4), "tiled-walk-vf": (fc.TUNE_SWEEP_TILED, 5)}[mode]
    su = np.random.default_rng(5).standard_normal(mesh.numCells) * mesh.vol[:mesh.numCells]
    res = []
    for p2p in (0, 1):
        ctx = fc.Context(0)
        ctx.set_mesh(mesh)
        ctx.create_csr(download=False)
        ctx.set_tuning(fc.TUNE_SWEEP_TILED, 0)     # the baseline of the comparison is the level schedule
        ctx.set_tuning(key, value * p2p)
        if mode.startswith("tiled") and p2p:
            ctx.set_tuning(fc.TUNE_SWEEP_CHECK, 1)    # every tiled sweep is compared with the level sweep on the device
        ctx.upload("APU", -np.ones(mesh.numCells))
        ctx.upload("SU", su)
        ctx.fill("PP", 0.0)
        ctx.laplacian("APU", "PP")
        best = None
        for rep_i in range(3):
            if rep_i == 1:
                ctx.set_tuning(fc.TUNE_SWEEP_CHECK, 0)   # the first solve checks, the others are timed
            ctx.fill("PP", 0.0)
            t0 = time.perf_counter()
            rep = ctx.solve(solver, "PP", fc.solver_opts(1e-9, 300))
            wall = time.perf_counter() - t0
            ms = ctx.timings().solve_ms
            best = ms if best is None else min(best, ms)
        res.append((rep.iters, rep.res0, rep.resl, ctx.download("PP"), ctx.download("RES"), best))
        if mode.startswith("tiled") and p2p:
            print("\n[schedule]", name, ctx.sweep_schedule_info())
            if name in ("hex", "slab", "poly"):
                assert ctx.timings().sweep_tiles > 0, "the tiled schedule was not used: " + ctx.sweep_schedule_info()
        ctx.close()
    (i0, a0, b0, x0, r0, t0), (i1, a1, b1, x1, r1, t1) = res
    print(f"\\n[{mode} sweeps] {name} {solver}: {i0} iterations, level mode {t0:.3f} ms, {mode} mode {t1:.3f} ms")
    assert i0 == i1 and a0 == a1 and b0 == b1
    assert np.array_equal(x0, x1) and np.array_equal(r0, r1)


@pytest.mark.parametrize("name", ["hex", "poly", "pitzDaily"])
@pytest.mark.parametrize("nigrad", [1, 2])
def test_fused_velocity_gradients_are_bit_identical(fc, name, nigrad):
    """FC_TUNE_FUSED_GRAD: grad(U), grad(V), grad(W) of calcp / calcuvw in one kernel per pass; gradients, matrix and
    right-hand side of both assemblies must not change by a bit."""
    mesh = MESHES[name]()
    f = cases.flow_fields(mesh)
    res = []
    for fused in (0, 1):
        ctx = fc.Context(0)
        ctx.set_mesh(mesh)
        ctx.create_csr(download=False)
        ctx.set_tuning(fc.TUNE_FUSED_GRAD, fused)
        for k, fld in (("u", "U"), ("v", "V"), ("w", "W"), ("p", "P"), ("den", "DEN"), ("apu", "APU"), ("apv", "APV"),
                       ("apw", "APW")):
            ctx.upload(fld, f[k])
        ctx.upload("VIS", np.full(mesh.numTotal, 0.01))
        ctx.fill("FLMASS", 0.0)
        ctx.grad_gauss("P", "DPDXI", 1)
        ctx.calcp_assemble(fc.calcp_opts(solver="iccg", const_mflux=True, nigrad=nigrad))
        out = [ctx.download(k) for k in ("DUDXI", "DVDXI", "DWDXI", "A", "SU", "FLMASS")]
        ctx.calcuvw_assemble(fc.calcuvw_opts(scheme="muscl-f", nigrad=nigrad))
        out += [ctx.download(k) for k in ("DUDXI", "A", "SU", "SV", "SW")]
        res.append(out)
        ctx.close()
    for a, b in zip(*res):
        assert np.array_equal(a, b)
