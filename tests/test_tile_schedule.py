"""The tiled schedule of the DIC / DILU triangular sweeps (freecappuccino_b200/csrc/fc_tile_schedule.hpp, used by
k_tile_sweep of fc_trisolve.cu under FC_TUNE_SWEEP_TILED) checked on the CPU: tests/kernel_bodies_host/fct_host.cpp
compiles the same header with g++ and walks the schedule the way the kernel does.  Every sweep mode must be
bit-identical to the natural-order sweep of the reference (iccg.f90:77-111, bicgstab.f90:68-79, :117-136) and must
never read a value before its producer -- through shared memory inside a tile, through a strictly lower tile level
otherwise."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from freecappuccino_b200 import cases, mesh as M
from oracle import oracle

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernel_bodies_host")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FWD, BWD, DIC, DIC_PAR, DILU = range(5)


@pytest.fixture(scope="module")
def host():
    src = os.path.join(HERE, "fct_host.cpp")
    so = os.path.join(HERE, "_fct_host.so")
    hdrs = [os.path.join(HERE, "..", "..", "freecappuccino_b200", "csrc", f) for f in ("fc_tile_schedule.hpp",
                                                                                        "fc_tile_sweep.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++20", "-pthread", "-o", so, src])
    lib = C.CDLL(so)
    lib.fct_build.restype = C.c_void_p
    lib.fct_build_shrunk.restype = C.c_void_p
    lib.fct_why.restype = C.c_char_p
    return lib


def ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class System:
    """0-based CSR of the mesh's pattern with a diagonally dominant non-symmetric matrix on it."""

    def __init__(self, mesh, seed=3):
        csr = oracle.create_csr(mesh)
        self.n = n = mesh.numCells
        self.ioffset = (csr.ioffset - 1).astype(np.int32)
        self.ja = (csr.ja - 1).astype(np.int32)
        self.diag = (csr.diag - 1).astype(np.int32)
        rng = np.random.default_rng(seed)
        self.a = -(0.2 + rng.random(self.ja.size))
        row = np.repeat(np.arange(n), np.diff(self.ioffset))
        rowsum = np.zeros(n)
        np.add.at(rowsum, row, np.abs(self.a))
        self.a[self.diag] = rowsum + 1.0
        # position of the transposed entry (bicgstab.f90:68-79 finds a_ki by a search in row k)
        key = row.astype(np.int64) * n + self.ja
        order = np.argsort(key)
        tkey = self.ja.astype(np.int64) * n + row
        self.tpos = order[np.searchsorted(key[order], tkey)].astype(np.int32)
        assert np.array_equal(self.ja[self.tpos], row)
        self.xc, self.yc, self.zc = (np.ascontiguousarray(v[:n], dtype=np.float64) for v in (mesh.xc, mesh.yc, mesh.zc))
        self.r = rng.standard_normal(n)


MESHES = {
    "hex-24x20x17": lambda: cases.hex_case(24, 20, 17),
    "hex-32^3": lambda: cases.hex_case(32, 32, 32),
    "hex-9^3": lambda: cases.hex_case(9, 9, 9),
    "slab-60x60x1": lambda: cases.hex_case(60, 60, 1, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry")),
    "line-100": lambda: cases.hex_case(100, 1, 1),
    "one-cell": lambda: cases.hex_case(1, 1, 1),
    "skew": lambda: cases.skew_case(12),
    "poly-6": lambda: cases.poly_case(6),
    "cavity": lambda: cases.golden_mesh(os.path.join(GOLD, "cavity.npz")),
    "pitzDaily": lambda: cases.golden_mesh(os.path.join(GOLD, "pitzDaily.npz")),
    "hex-rank-of-3": lambda: M.partition(cases.hex_case(16, 16, 24), M.slab_ranks(16 * 16 * 24, 3), 3)[1],
    "poly-rank-of-4": lambda: (lambda g: M.partition(g, M.rcb_ranks(g, 4), 4)[2])(cases.poly_case(9)),
    "pitzDaily-rank-of-2": lambda: (lambda g: M.partition(g, M.rcb_ranks(g, 2), 2)[1])(
        cases.golden_mesh(os.path.join(GOLD, "pitzDaily.npz"))),
}


def build(host, s):
    h = host.fct_build(s.n, ip(s.ioffset), ip(s.ja), ip(s.diag), dp(s.xc), dp(s.yc), dp(s.zc))
    info = np.zeros(11, np.int32)
    host.fct_info(C.c_void_p(h), ip(info))
    return h, info


def run(host, h, s, mode, d, src, padd=0.0, small=1e-20):
    ref = np.zeros(s.n)
    host.fct_reference_sweep(mode, s.n, ip(s.ioffset), ip(s.ja), ip(s.diag), ip(s.tpos), dp(s.a), dp(d), dp(src), dp(ref),
                             C.c_double(small), C.c_double(padd))
    out = np.full(s.n, np.nan)
    bad = host.fct_sweep(C.c_void_p(h), mode, s.n, ip(s.ioffset), ip(s.diag), ip(s.tpos), dp(s.a), dp(d), dp(src), dp(out),
                         C.c_double(small), C.c_double(padd))
    return ref, out, bad


@pytest.mark.parametrize("name", list(MESHES))
def test_tiled_sweeps_equal_natural_order_sweeps(host, name):
    s = System(MESHES[name]())
    h, info = build(host, s)
    try:
        assert host.fct_ok(C.c_void_p(h)), host.fct_why(C.c_void_p(h)).decode()
        ntiles, cells, maxrows, lnlev, lloc, unlev, uloc, p2p_ok, maxprod, repaired, cost = info
        assert 1 <= maxrows <= 512 and ntiles >= 1 and 2 <= cells <= 512
        zero = np.zeros(s.n)
        for mode in (DIC, DIC_PAR, DILU):
            ref, out, bad = run(host, h, s, mode, zero, zero, padd=1e-20 if mode == DIC_PAR else 0.0)
            assert bad == 0, (name, mode, bad)
            assert np.array_equal(ref, out), (name, mode)
        d = ref                                   # the DILU diagonal
        rt, t, bad = run(host, h, s, FWD, d, s.r)
        assert bad == 0 and np.array_equal(rt, t)
        rz, z, bad = run(host, h, s, BWD, d, t)
        assert bad == 0 and np.array_equal(rz, z)
        assert np.all(np.isfinite(z))
    finally:
        host.fct_free(C.c_void_p(h))


def test_tile_levels_of_a_cube_are_the_tile_hyperplanes(host):
    """n = 32: 4 x 4 x 4 tiles of 8^3 cells -> 3*4-2 = 10 tile levels instead of 3*32-2 = 94 row levels, and
    3*8-2 = 22 local levels per tile."""
    s = System(cases.hex_case(32, 32, 32))
    h, info = build(host, s)
    ntiles, cells, maxrows, lnlev, lloc, unlev, uloc, p2p_ok, maxprod, repaired, cost = info
    host.fct_free(C.c_void_p(h))
    assert (ntiles, cells, maxrows) == (64, 8, 512)
    assert (lnlev, unlev) == (10, 10) and (lloc, uloc) == (22, 22)
    assert p2p_ok == 1 and maxprod == 3          # the three face neighbours on the low (high) side


def row_levels(s):
    """Dependency levels of the strict lower triangle = hand-overs of the one-level schedule."""
    lev = np.zeros(s.n, np.int64)
    for i in range(s.n):
        k0, k1 = s.ioffset[i], s.diag[i]
        if k1 > k0:
            lev[i] = lev[s.ja[k0:k1]].max() + 1
    return int(lev.max()) + 1


def test_a_numbering_that_is_not_monotone_across_the_bins_is_repaired_not_trusted(host):
    """Cell centres that do not follow the numbering at all: every bin depends on every other one in a circle.  The
    builder must not emit a schedule that deadlocks -- it cuts the circular bins into runs of consecutive rows (a
    chain) -- and the walk must still be exact.  (Whether such a chain is used is the library's decision from the
    cost estimate: at 8 000 rows it still beats 58 row levels, at 10 M rows it would not.)"""
    s = System(cases.hex_case(20, 20, 20))
    perm = np.random.default_rng(0).permutation(s.n)
    s.xc, s.yc, s.zc = s.xc[perm].copy(), s.yc[perm].copy(), s.zc[perm].copy()
    h, info = build(host, s)
    ntiles, cells, maxrows, lnlev, lloc, unlev, uloc, p2p_ok, maxprod, repaired, cost = info
    try:
        assert host.fct_ok(C.c_void_p(h))
        assert repaired > 0.9 * s.n and lnlev > 0.8 * ntiles      # (almost) one chain of runs
        zero = np.zeros(s.n)
        ref, out, bad = run(host, h, s, DILU, zero, zero)
        assert bad == 0 and np.array_equal(ref, out)
        rz, z, bad = run(host, h, s, BWD, ref, s.r)
        assert bad == 0 and np.array_equal(rz, z)
        assert cost >= 35 * ntiles                    # a chain: one hand-over per run, the estimate says so
    finally:
        host.fct_free(C.c_void_p(h))


def test_block_structured_and_jittered_meshes_tile_without_repair(host):
    """pitzDaily (five blocks, each numbered on its own) and the jittered polyhedral mesh (centres on the bin
    boundaries fall on either side at random): raw bins are circular there; after the monotone pass over the bin
    coordinates no bin has to be cut, and the sweeps need far fewer hand-overs than row levels."""
    s = System(cases.golden_mesh(os.path.join(GOLD, "pitzDaily.npz")))
    h, info = build(host, s)
    ntiles, cells, maxrows, lnlev, lloc, unlev, uloc, p2p_ok, maxprod, repaired, cost = info
    host.fct_free(C.c_void_p(h))
    nlev = row_levels(s)
    assert repaired == 0
    assert lnlev < nlev // 4 and cost < 0.7 * 42 * nlev
    s = System(cases.poly_case(16))
    h, info = build(host, s)
    host.fct_free(C.c_void_p(h))
    assert info[9] == 0 and info[3] < row_levels(s) // 3 and info[10] < 0.7 * 42 * row_levels(s)


MESHES["hex-16x12x10"] = lambda: cases.hex_case(16, 12, 10)
MESHES["slab-40x30x1"] = lambda: cases.hex_case(40, 30, 1, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry"))
# 512 host threads and a std::barrier per local level: keep the meshes to a few tiles (pitzDaily would take 40 s)
EMU_MESHES = ["hex-16x12x10", "slab-40x30x1", "poly-6", "skew", "hex-rank-of-3", "one-cell"]


@pytest.mark.parametrize("name", EMU_MESHES)
@pytest.mark.parametrize("p2p", [0, 1])
def test_kernel_source_run_on_host_threads_equals_natural_order_sweeps(host, name, p2p):
    """fc_tile_sweep.cuh itself (not a restatement): 512 host threads per CTA, std::barrier for __syncthreads, CTAs in
    ticket order, two launches in a row on the same counters.  All five modes, both register-prefetch widths."""
    if not host.fct_can_emulate():
        pytest.skip("this machine cannot run 512 threads at once")
    s = System(MESHES[name]())
    h, info = build(host, s)
    try:
        assert host.fct_ok(C.c_void_p(h))
        if p2p:
            assert info[7] == 1
        zero = np.zeros(s.n)

        def emu(mode, d, src, padd=0.0, pre8=0):
            ref = np.zeros(s.n)
            host.fct_reference_sweep(mode, s.n, ip(s.ioffset), ip(s.ja), ip(s.diag), ip(s.tpos), dp(s.a), dp(d), dp(src),
                                     dp(ref), C.c_double(1e-20), C.c_double(padd))
            out = np.zeros(s.n)
            rc = host.fct_emu_sweep(C.c_void_p(h), mode, pre8, p2p, 2, s.n, ip(s.ioffset), ip(s.diag), ip(s.tpos), dp(s.a),
                                    dp(d), dp(src), dp(out), C.c_double(1e-20), C.c_double(padd))
            assert rc == 0
            assert np.array_equal(ref, out), (name, mode, pre8, p2p)
            return out

        for pre8 in (0, 1):
            emu(DIC, zero, zero, pre8=pre8)
            emu(DIC_PAR, zero, zero, padd=1e-20, pre8=pre8)
            d = emu(DILU, zero, zero, pre8=pre8)
            t = emu(FWD, d, s.r, pre8=pre8)
            emu(BWD, d, t, pre8=pre8)
    finally:
        host.fct_free(C.c_void_p(h))


@pytest.mark.parametrize("name", EMU_MESHES + ["pitzDaily", "poly-rank-of-4", "hex-24x20x17"])
@pytest.mark.parametrize("hand", [1, 0])
def test_tile_walk_kernel_source_on_host_threads_equals_natural_order_sweeps(host, name, hand):
    """k_tile_walk of fc_tile_sweep.cuh (FC_TUNE_SWEEP_TILED = 4: 256 threads stage the tile in shared memory in
    ascending row order, fold the values of other tiles into the coefficients and retire; 64 threads walk the local
    levels branch-free; hand-over by producer flags (1) or tile-level counters (0)): the kernel source on host threads,
    two launches in a row."""
    s = System(MESHES[name]())
    h, info = build(host, s)
    try:
        assert host.fct_ok(C.c_void_p(h))
        zero = np.zeros(s.n)

        def emu(mode, d, src, padd=0.0, pre8=0):
            ref = np.zeros(s.n)
            host.fct_reference_sweep(mode, s.n, ip(s.ioffset), ip(s.ja), ip(s.diag), ip(s.tpos), dp(s.a), dp(d), dp(src),
                                     dp(ref), C.c_double(1e-20), C.c_double(padd))
            out = np.zeros(s.n)
            rc = host.fct_emu_walk(C.c_void_p(h), mode, pre8, hand, s.n, ip(s.ioffset), ip(s.diag), ip(s.tpos), dp(s.a),
                                   dp(d), dp(src), dp(out), C.c_double(1e-20), C.c_double(padd))
            assert rc == 0, (name, mode, pre8, rc)
            assert np.array_equal(ref, out), (name, mode, pre8)
            return out

        for pre8 in (0, 1, 2):   # registers / shared slots per row: 4, 8, 3 (rows longer than that take the slow path)
            emu(DIC, zero, zero, pre8=pre8)
            emu(DIC_PAR, zero, zero, padd=1e-20, pre8=pre8)
            d = emu(DILU, zero, zero, pre8=pre8)
            t = emu(FWD, d, s.r, pre8=pre8)
            emu(BWD, d, t, pre8=pre8)
    finally:
        host.fct_free(C.c_void_p(h))


@pytest.mark.parametrize("name", EMU_MESHES)
def test_value_as_flag_kernel_source_on_host_threads_equals_natural_order_sweeps(host, name):
    """k_tile_sweep_vf of fc_tile_sweep.cuh (FC_TUNE_SWEEP_TILED = 3: a row polls the VALUE of an out-of-tile dependency,
    no flags or counters): same emulation, all five modes, both prefetch widths; also checks that the forward sweep
    re-arms the backward sweep's target and the backward sweep its own input."""
    if not host.fct_can_emulate():
        pytest.skip("this machine cannot run 512 threads at once")
    s = System(MESHES[name]())
    h, info = build(host, s)
    try:
        assert host.fct_ok(C.c_void_p(h))
        zero = np.zeros(s.n)

        def emu(mode, d, src, padd=0.0, pre8=0):
            ref = np.zeros(s.n)
            host.fct_reference_sweep(mode, s.n, ip(s.ioffset), ip(s.ja), ip(s.diag), ip(s.tpos), dp(s.a), dp(d), dp(src),
                                     dp(ref), C.c_double(1e-20), C.c_double(padd))
            out = np.zeros(s.n)
            rc = host.fct_emu_sweep_vf(C.c_void_p(h), mode, pre8, s.n, ip(s.ioffset), ip(s.diag), ip(s.tpos), dp(s.a),
                                       dp(d), dp(src), dp(out), C.c_double(1e-20), C.c_double(padd))
            assert rc == 0, (name, mode, pre8, rc)
            assert np.array_equal(ref, out), (name, mode, pre8)
            return out

        for pre8 in (0, 1):
            emu(DIC, zero, zero, pre8=pre8)
            emu(DIC_PAR, zero, zero, padd=1e-20, pre8=pre8)
            d = emu(DILU, zero, zero, pre8=pre8)
            t = emu(FWD, d, s.r, pre8=pre8)
            emu(BWD, d, t, pre8=pre8)
    finally:
        host.fct_free(C.c_void_p(h))


@pytest.mark.parametrize("name", EMU_MESHES + ["pitzDaily", "poly-rank-of-4", "hex-24x20x17"])
def test_tile_walk_vf_kernel_source_on_host_threads_equals_natural_order_sweeps(host, name):
    """k_tile_walk_vf of fc_tile_sweep.cuh (FC_TUNE_SWEEP_TILED = 5: the staged walk of mode 4 with the value-as-flag
    hand-over of mode 3 -- 256 helper threads poll and fold the values of other tiles, 64 walkers wait per local level
    on a shared-memory count): the kernel source on host threads, all five modes, the three staging widths; also that
    the forward sweep re-arms the backward sweep's target and the backward sweep its own input."""
    s = System(MESHES[name]())
    h, info = build(host, s)
    try:
        assert host.fct_ok(C.c_void_p(h))
        zero = np.zeros(s.n)

        def emu(mode, d, src, padd=0.0, pre8=0):
            ref = np.zeros(s.n)
            host.fct_reference_sweep(mode, s.n, ip(s.ioffset), ip(s.ja), ip(s.diag), ip(s.tpos), dp(s.a), dp(d), dp(src),
                                     dp(ref), C.c_double(1e-20), C.c_double(padd))
            out = np.zeros(s.n)
            rc = host.fct_emu_walk_vf(C.c_void_p(h), mode, pre8, s.n, ip(s.ioffset), ip(s.diag), ip(s.tpos), dp(s.a),
                                      dp(d), dp(src), dp(out), C.c_double(1e-20), C.c_double(padd))
            assert rc == 0, (name, mode, pre8, rc)
            assert np.array_equal(ref, out), (name, mode, pre8)
            return out

        for pre8 in (0, 1, 2):   # staged entries per row: 4, 8, 3 (rows longer than that poll inside the walk)
            emu(DIC, zero, zero, pre8=pre8)
            emu(DIC_PAR, zero, zero, padd=1e-20, pre8=pre8)
            d = emu(DILU, zero, zero, pre8=pre8)
            t = emu(FWD, d, s.r, pre8=pre8)
            emu(BWD, d, t, pre8=pre8)
    finally:
        host.fct_free(C.c_void_p(h))


@pytest.mark.parametrize("name", ["poly-6", "poly-rank-of-4", "poly-12", "hex-24x20x17", "pitzDaily"])
@pytest.mark.parametrize("shrink", [1, 2, 3])
def test_narrower_bins_give_valid_and_exact_schedules(host, name, shrink):
    """fc_build_tile_schedule(min_shrink): narrower starting bins (FC_TILE_MIN_SHRINK in the library, a measurement
    knob: profiles/r02_poly_bins.txt).  Every bin width must give an acyclic tile graph, no row
    visited before its dependencies, values through global memory only from lower tile levels / named producers, and
    sweeps bit-identical to the natural-order ones -- also on a rank-local mesh of a partition."""
    mesh = cases.poly_case(12) if name == "poly-12" else MESHES[name]()
    s = System(mesh)
    h = host.fct_build_shrunk(s.n, ip(s.ioffset), ip(s.ja), ip(s.diag), dp(s.xc), dp(s.yc), dp(s.zc), shrink)
    try:
        assert host.fct_ok(C.c_void_p(h)), host.fct_why(C.c_void_p(h)).decode()
        info = np.zeros(11, np.int32)
        host.fct_info(C.c_void_p(h), ip(info))
        assert info[2] <= 512   # rows of the largest tile
        zero = np.zeros(s.n)
        for mode, padd in ((DIC, 0.0), (DIC_PAR, 1e-20), (DILU, 0.0)):
            ref, out, bad = run(host, h, s, mode, zero, zero, padd=padd)
            assert bad == 0 and np.array_equal(ref, out), (name, shrink, mode, bad)
        d = out
        ref, t, bad = run(host, h, s, FWD, d, s.r)
        assert bad == 0 and np.array_equal(ref, t), (name, shrink, "fwd", bad)
        ref, z, bad = run(host, h, s, BWD, d, t)
        assert bad == 0 and np.array_equal(ref, z), (name, shrink, "bwd", bad)
    finally:
        host.fct_free(C.c_void_p(h))


def renumbered(s, perm):
    """The same matrix with rows / columns renumbered: new row perm[i] = old row i; columns ascending again."""
    n = s.n
    row = np.repeat(np.arange(n), np.diff(s.ioffset))
    r2, c2 = perm[row], perm[s.ja]
    order = np.lexsort((c2, r2))
    t = System.__new__(System)
    t.n = n
    t.ja = c2[order].astype(np.int32)
    t.a = s.a[order].copy()
    rr = r2[order]
    t.ioffset = np.zeros(n + 1, np.int32)
    np.add.at(t.ioffset, rr + 1, 1)
    t.ioffset = np.cumsum(t.ioffset).astype(np.int32)
    t.diag = np.flatnonzero(t.ja == rr).astype(np.int32)
    key = rr.astype(np.int64) * n + t.ja
    tkey = t.ja.astype(np.int64) * n + rr
    t.tpos = np.searchsorted(key, tkey).astype(np.int32)          # key is sorted: rows ascending, columns ascending
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)
    t.xc, t.yc, t.zc = s.xc[inv].copy(), s.yc[inv].copy(), s.zc[inv].copy()
    t.r = s.r[inv].copy()
    return t


@pytest.mark.parametrize("seed", range(8))
def test_any_numbering_gives_a_valid_and_exact_schedule(host, seed):
    """Numberings between 'lexicographic' and 'random': blocks of cells renumbered at random, slabs reversed, a few
    cells swapped.  Whatever the bins look like afterwards (circular, oversized, cut into runs), the walk must visit
    every row after its dependencies and reproduce the natural-order sweep of the renumbered matrix bit for bit."""
    rng = np.random.default_rng(seed)
    base = System(cases.hex_case(14, 11, 9) if seed % 2 == 0 else cases.poly_case(7))
    n = base.n
    perm = np.arange(n)
    kind = seed % 4
    if kind == 0:                                   # random renumbering of one third of the cells
        idx = rng.choice(n, n // 3, replace=False)
        perm[idx] = perm[rng.permutation(idx)]
    elif kind == 1:                                 # reverse the numbering of the upper half
        perm[n // 2:] = perm[n // 2:][::-1]
    elif kind == 2:                                 # blocks of 50 consecutive cells in random order
        blocks = [np.arange(i, min(i + 50, n)) for i in range(0, n, 50)]
        perm = np.empty(n, np.int64)
        perm[np.concatenate([blocks[b] for b in rng.permutation(len(blocks))])] = np.arange(n)
    else:                                           # completely random
        perm = rng.permutation(n)
    s = renumbered(base, perm)
    assert np.array_equal(s.ja[s.diag], np.arange(n)) and np.array_equal(s.ja[s.tpos], np.repeat(np.arange(n), np.diff(s.ioffset)))
    h, info = build(host, s)
    try:
        assert host.fct_ok(C.c_void_p(h)), host.fct_why(C.c_void_p(h)).decode()
        assert info[2] <= 512
        zero = np.zeros(n)
        d, out, bad = run(host, h, s, DILU, zero, zero)
        assert bad == 0 and np.array_equal(d, out)
        rt, t, bad = run(host, h, s, FWD, d, s.r)
        assert bad == 0 and np.array_equal(rt, t)
        rz, z, bad = run(host, h, s, BWD, d, t)
        assert bad == 0 and np.array_equal(rz, z)
    finally:
        host.fct_free(C.c_void_p(h))
