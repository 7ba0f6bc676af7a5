"""GPU tests of the kernel variants behind the C ABI: the TMA-staged SpMV pipeline vs the CSR-stream
kernel, the persistent cooperative DPCG kernel vs one launch per vector operation, ragged and
over-long rows, and -- at BASELINE's full 216^3 size -- size-independent properties that need no
oracle run (symmetry of the assembled operator, the true residual of the converged solve).
"""
import os

import numpy as np
import pytest

from freecappuccino_b200 import cases, mesh as M
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fc():
    from freecappuccino_b200 import lib
    return lib


def banded_csr(n, half_width, rng, ragged=True):
    """1-based CSR of a symmetric positive definite banded matrix; rows keep a random subset of the band."""
    rows, cols = [], []
    for i in range(n):
        lo, hi = max(0, i - half_width), min(n - 1, i + half_width)
        for j in range(lo, hi + 1):
            if j == i or not ragged or ((i * 2654435761 + j * 40503) ^ (j * 2654435761 + i * 40503)) % 3 != 0:
                rows.append(i); cols.append(j)
    rows, cols = np.array(rows), np.array(cols)
    val = np.where(rows == cols, 0.0, -rng.random(rows.size) - 0.1)
    # symmetrise the values and make the matrix diagonally dominant
    key = {}
    for k, (i, j) in enumerate(zip(rows, cols)):
        key[(i, j)] = k
    for (i, j), k in key.items():
        if i < j and (j, i) in key:
            val[key[(j, i)]] = val[k]
    rowsum = np.zeros(n)
    np.add.at(rowsum, rows, np.abs(val))
    val[rows == cols] = rowsum + 1.0
    ioffset = np.zeros(n + 1, np.int32)
    np.add.at(ioffset, rows + 1, 1)
    ioffset = (np.cumsum(ioffset) + 1).astype(np.int32)
    diag = np.array([key[(i, i)] + 1 for i in range(n)], np.int32)
    return ioffset, (cols + 1).astype(np.int32), diag, val


@pytest.mark.parametrize("n,half_width", [(5, 1), (333, 3), (1000, 12), (700, 60)])
@pytest.mark.parametrize("solver", ["dpcg", "iccg", "bicgstab"])
def test_explicit_csr_ragged_rows_all_kernel_variants(fc, n, half_width, solver):
    """fc_solve_csr on ragged banded systems: row counts that are not multiples of 32 or 256, rows of
    very different length, and (half_width 60) 256-row chunks that exceed every staging buffer, which
    takes the pipeline's straight-from-global fallback.  All kernel variants must agree with the oracle."""
    rng = np.random.default_rng(n + half_width)
    ioffset, ja, diag, a = banded_csr(n, half_width, rng)
    b = rng.standard_normal(n)
    csr = oracle.Csr(ioffset, ja, diag)
    xo = np.zeros(n)
    res0_o, resl_o, iters_o, _ = oracle.solve(solver, csr, a, b, xo, sor=1e-10, nsw=500)
    opts = fc.solver_opts(1e-10, 500)
    for spmv_kernel in (0, 1):
        for persistent in ((0, 1) if solver == "dpcg" else (0,)):
            ctx = fc.Context(0)
            ctx.set_tuning(fc.TUNE_SPMV_KERNEL, spmv_kernel)
            ctx.set_tuning(fc.TUNE_DPCG_PERSISTENT, persistent)
            x = np.zeros(n)
            rep = ctx.solve_csr(solver, ioffset, ja, diag, a, b, x, opts)
            assert abs(rep.iters - iters_o) <= 1, (spmv_kernel, persistent, rep.iters, iters_o)
            assert rep.res0 == pytest.approx(res0_o, rel=1e-12)
            assert cases.rel_l2(x, xo) < 1e-8, (spmv_kernel, persistent)
            if persistent:
                assert ctx.timings().persist_iters == rep.iters, "the persistent kernel did not run"
            ctx.close()


@pytest.mark.parametrize("name", ["hex", "skew"])
def test_spmv_variants_bit_identical(fc, name):
    mesh = cases.hex_case(37, 29, 13) if name == "hex" else cases.skew_case(21, 17, 9)
    ref = oracle.create_csr(mesh)
    rng = np.random.default_rng(5)
    a = rng.standard_normal(mesh.nnz)
    x = rng.standard_normal(mesh.numTotal)
    want = oracle.spmv(ref, a, x)
    for kernel in (0, 1):
        for geo in (0, 1, 2, 3):
            ctx = fc.Context(0)
            ctx.set_mesh(mesh)
            ctx.create_csr(download=False)
            ctx.set_tuning(fc.TUNE_SPMV_KERNEL, kernel)
            ctx.set_tuning(fc.TUNE_PIPE_GEOMETRY, geo)
            ctx.upload("A", a)
            ctx.upload("PP", x)
            ctx.spmv("PP", "SCRATCH_T")
            assert np.array_equal(ctx.download("SCRATCH_T")[:mesh.numCells], want), (kernel, geo)
            ctx.close()
            if kernel == 0:
                break


def test_dpcg_persistent_equals_multi_kernel(fc):
    """Same system, both DPCG paths: identical iteration count, residual histories equal to round-off
    (the reductions group the same terms differently), same solution."""
    mesh = cases.hex_case(40, 36, 20, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry"))
    out = {}
    for persistent in (0, 1):
        ctx = fc.Context(0)
        ctx.set_mesh(mesh)
        ctx.create_csr(download=False)
        ctx.set_tuning(fc.TUNE_DPCG_PERSISTENT, persistent)
        ctx.upload("APU", -np.ones(mesh.numCells))
        ctx.upload("SU", cases.poisson_rhs(mesh))
        ctx.fill("PP", 0.0)
        ctx.laplacian("APU", "PP")
        rep = ctx.solve("dpcg", "PP", fc.solver_opts(1e-9, 5000))
        out[persistent] = (rep.iters, rep.res0, rep.resl, ctx.download("PP")[:mesh.numCells], ctx.timings().persist_iters)
        ctx.close()
    assert out[1][4] == out[1][0] and out[0][4] == 0
    assert out[0][0] == out[1][0]
    assert out[0][1] == pytest.approx(out[1][1], rel=1e-13)
    assert out[0][2] == pytest.approx(out[1][2], rel=1e-6)
    assert cases.rel_l2(out[0][3], out[1][3]) < 1e-10


@pytest.mark.parametrize("name", ["hex", "hex_big", "poly", "skew", "shuffled"])
def test_dpcg_coded_columns_are_bit_identical(fc, name):
    """FC_TUNE_JA_CODED: the persistent kernel reads one-byte column codes (ja = row + offset[code]) instead of `ja`
    when the pattern has at most 256 distinct column offsets.  Same columns, same order, same grid: iteration count,
    residuals, solution and residual vector must not differ by a bit.  A numbering with more than 256 offsets (rows
    shuffled at random) keeps `ja`."""
    mesh = {"hex": lambda: cases.hex_case(40, 36, 20, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry")),
            "hex_big": lambda: cases.hex_case(96, 64, 50),
            "poly": lambda: cases.poly_case(9), "skew": lambda: cases.skew_case(14, 12, 10),
            "shuffled": lambda: cases.hex_case(12, 11, 10)}[name]()
    su = np.random.default_rng(3).standard_normal(mesh.numCells) * mesh.vol[:mesh.numCells]
    out = {}
    for coded in (0, 1):
        ctx = fc.Context(0)
        ctx.set_tuning(fc.TUNE_JA_CODED, coded)
        if name == "shuffled":   # explicit CSR of a randomly renumbered Laplacian: ~n distinct offsets
            n = mesh.numCells
            ioffset, ja, diag, a = oracle_csr_shuffled(mesh)
            x = np.zeros(n)
            rep = ctx.solve_csr("dpcg", ioffset, ja, diag, a, su[:n], x, fc.solver_opts(1e-9, 5000))
            res = np.zeros(0)
        else:
            ctx.set_mesh(mesh)
            ctx.create_csr(download=False)
            ctx.upload("APU", -np.ones(mesh.numCells))
            ctx.upload("SU", su)
            ctx.upload("PP", 0.01 * np.cos(np.arange(mesh.numTotal)))
            ctx.laplacian("APU", "PP")
            rep = ctx.solve("dpcg", "PP", fc.solver_opts(1e-9, 5000))
            x, res = ctx.download("PP"), ctx.download("RES")
        t = ctx.timings()
        assert t.persist_iters == rep.iters, "the persistent kernel did not run"
        assert rep.iters > 5
        if name == "shuffled":
            assert t.column_offsets == 0 and t.persist_index_bytes == 4
        elif name in ("hex", "hex_big"):
            assert t.column_offsets == 7 and t.persist_index_bytes == (1 if coded else 4)
        else:
            assert t.persist_index_bytes == (1 if coded and t.column_offsets else 4)
        out[coded] = (rep.iters, rep.res0, rep.resl, x, res)
        ctx.close()
    assert out[0][:3] == out[1][:3], (name, out[0][:3], out[1][:3])
    assert np.array_equal(out[0][3], out[1][3]) and np.array_equal(out[0][4], out[1][4]), name


@pytest.mark.parametrize("name", ["hex", "poly", "skew"])
def test_dpcg_eager_x_update_is_bit_identical(fc, name):
    """FC_TUNE_DPCG_EAGER: fi += alf*pk behind the beta reduction and q = res/a_ii handed from the x/r update to the
    p-update.  Same operands and rounding, same grid: iteration count, residuals, solution and final residual vector
    must not differ by a bit -- also with nsw = 0, one iteration, and a stop after a few iterations."""
    mesh = {"hex": lambda: cases.hex_case(40, 36, 20, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry")),
            "poly": lambda: cases.poly_case(9), "skew": lambda: cases.skew_case(14, 12, 10)}[name]()
    su = np.random.default_rng(3).standard_normal(mesh.numCells) * mesh.vol[:mesh.numCells]
    for sor, nsw in ((1e-9, 5000), (1e-30, 1), (1e-30, 7), (1e-9, 0)):
        out = {}
        for eager in (0, 1):
            ctx = fc.Context(0)
            ctx.set_mesh(mesh)
            ctx.create_csr(download=False)
            ctx.set_tuning(fc.TUNE_DPCG_EAGER, eager)
            ctx.upload("APU", -np.ones(mesh.numCells))
            ctx.upload("SU", su)
            ctx.upload("PP", 0.01 * np.cos(np.arange(mesh.numTotal)))
            ctx.laplacian("APU", "PP")
            rep = ctx.solve("dpcg", "PP", fc.solver_opts(sor, nsw))
            assert ctx.timings().persist_iters == rep.iters
            out[eager] = (rep.iters, rep.res0, rep.resl, ctx.download("PP"), ctx.download("RES"))
            ctx.close()
        assert out[0][:3] == out[1][:3], (name, sor, nsw, out[0][:3], out[1][:3])
        assert np.array_equal(out[0][3], out[1][3]) and np.array_equal(out[0][4], out[1][4]), (name, sor, nsw)


def oracle_csr_shuffled(mesh):
    """CSR (1-based ioffset / ja / diag + values) of the mesh's unit Laplacian with the rows renumbered at random."""
    n = mesh.numCells
    perm = np.random.default_rng(11).permutation(n)
    own = perm[mesh.owner[:mesh.numInnerFaces] - 1]
    nei = perm[mesh.neighbour[:mesh.numInnerFaces] - 1]
    rows = np.concatenate([np.arange(n), own, nei])
    cols = np.concatenate([np.arange(n), nei, own])
    vals = np.concatenate([np.zeros(n), -np.ones(own.size), -np.ones(nei.size)])
    np.add.at(vals, own, 1.0)
    np.add.at(vals, nei, 1.0)
    vals[:n] += 0.1
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    ioffset = np.zeros(n + 1, np.int32)
    np.add.at(ioffset, rows + 1, 1)
    ioffset = np.cumsum(ioffset).astype(np.int32)
    diag = np.flatnonzero(rows == cols).astype(np.int32)
    return ioffset + 1, (cols + 1).astype(np.int32), diag + 1, vals


@pytest.mark.parametrize("name", ["hex", "poly", "skew"])
def test_dpcg_fused_p_scheme_is_bit_identical(fc, name):
    """FC_TUNE_DPCG_FUSED: the persistent kernel without a p-update phase (the product gathers p = q + bet*pold, the
    x/r update leaves q = res/a_ii behind).  Same operands and rounding in every expression and the same grid, hence
    the same partial sums: iteration count, residuals, solution and final residual vector must not differ by a bit.
    Also nsw = 0 and an early exit after one iteration (the deferred x update of the last iteration)."""
    mesh = {"hex": lambda: cases.hex_case(40, 36, 20, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry")),
            "poly": lambda: cases.poly_case(9), "skew": lambda: cases.skew_case(14, 12, 10)}[name]()
    su = np.random.default_rng(3).standard_normal(mesh.numCells) * mesh.vol[:mesh.numCells]
    for sor, nsw in ((1e-9, 5000), (1e-30, 1), (1e-30, 7), (1e-9, 0)):
        out = {}
        for fused in (0, 1):
            ctx = fc.Context(0)
            ctx.set_mesh(mesh)
            ctx.create_csr(download=False)
            ctx.set_tuning(fc.TUNE_DPCG_FUSED, fused)
            ctx.upload("APU", -np.ones(mesh.numCells))
            ctx.upload("SU", su)
            ctx.upload("PP", 0.01 * np.cos(np.arange(mesh.numTotal)))
            ctx.laplacian("APU", "PP")
            rep = ctx.solve("dpcg", "PP", fc.solver_opts(sor, nsw))
            assert ctx.timings().persist_iters == rep.iters
            out[fused] = (rep.iters, rep.res0, rep.resl, ctx.download("PP"), ctx.download("RES"))
            ctx.close()
        assert out[0][:3] == out[1][:3], (name, sor, nsw, out[0][:3], out[1][:3])
        assert np.array_equal(out[0][3], out[1][3]) and np.array_equal(out[0][4], out[1][4]), (name, sor, nsw)


def test_early_return_when_already_converged(fc):
    """dpcg.f90:66-70: res0 < tol returns before the first iteration (persistent and multi-kernel path)."""
    mesh = cases.hex_case(8, 8, 8)
    for persistent in (0, 1):
        ctx = fc.Context(0)
        ctx.set_mesh(mesh)
        ctx.create_csr(download=False)
        ctx.set_tuning(fc.TUNE_DPCG_PERSISTENT, persistent)
        ctx.upload("APU", -np.ones(mesh.numCells))
        ctx.fill("SU", 0.0)
        ctx.fill("PP", 0.0)
        ctx.laplacian("APU", "PP")
        rep = ctx.solve("dpcg", "PP", fc.solver_opts(1e-8, 100))
        assert rep.iters == 0 and rep.res0 == 0.0
        assert np.all(ctx.download("PP") == 0.0)
        ctx.close()


def test_full_size_216_properties(fc):
    """BASELINE config 4 at its full size (10 077 696 cells): the oracle would need minutes, so the checks are
    size-independent properties -- the assembled p' operator is symmetric (x.Ay == y.Ax), its rows sum to the
    boundary part only (A 1 = 0 in a closed box), and the converged DPCG solution satisfies the system:
    |su - A pp|_1 / res0 < sor, with the TRUE residual computed by an independent SpMV launch."""
    n = 216
    mesh = M.hex_mesh(n, n, n)
    f = cases.config4_fields(mesh)
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    for k, name in (("u", "U"), ("v", "V"), ("w", "W"), ("p", "P"), ("den", "DEN"), ("apu", "APU"), ("apv", "APV"),
                    ("apw", "APW")):
        ctx.upload(name, f[k])
    ctx.grad_gauss("P", "DPDXI", 1)
    opts = fc.calcp_opts(solver="dpcg", const_mflux=True, sor=1e-8, nsw=100000, pRefCell=1, urf_p=0.3)
    ctx.calcp_assemble(opts)
    su = ctx.download("SU")
    rng = np.random.default_rng(0)
    x = rng.standard_normal(mesh.numTotal)
    y = rng.standard_normal(mesh.numTotal)
    nc = mesh.numCells
    ctx.upload("USER0", x)
    ctx.spmv("USER0", "SCRATCH_T")
    ax = ctx.download("SCRATCH_T")[:nc]
    ctx.upload("USER0", y)
    ctx.spmv("USER0", "SCRATCH_T")
    ay = ctx.download("SCRATCH_T")[:nc]
    assert abs(np.dot(y[:nc], ax) - np.dot(x[:nc], ay)) <= 1e-10 * (np.linalg.norm(ax) * np.linalg.norm(y[:nc]))
    ctx.fill("USER0", 1.0)
    ctx.spmv("USER0", "SCRATCH_T")
    assert np.max(np.abs(ctx.download("SCRATCH_T")[:nc])) <= 1e-9 * np.max(np.abs(ax))
    # conservation: the mass-imbalance source of a closed box sums to zero
    assert abs(su.sum()) <= 1e-9 * np.abs(su).sum()
    ctx.fill("PP", 0.0)
    rep = ctx.solve("dpcg", "PP", fc.solver_opts(1e-8, 100000))
    assert 1000 < rep.iters < 2500 and rep.resl / rep.res0 < 1e-8
    assert ctx.timings().persist_iters == rep.iters
    ctx.spmv("PP", "SCRATCH_T")
    r = su - ctx.download("SCRATCH_T")[:nc]
    assert rep.res0 == pytest.approx(np.abs(su).sum(), rel=1e-12)
    assert np.abs(r).sum() / rep.res0 < 1.05e-8
    ctx.close()


@pytest.mark.parametrize("solver", ["dpcg", "iccg"])
def test_config3_full_size_poisson_against_the_oracle(fc, solver):
    """BASELINE config 3 at its full size (100^3 Poisson, rsm < 1e-8, nsw = 10000): iteration count within +-1 of the
    oracle's, solution within the solver tolerance, L-infinity error against sin(2 pi x) sin(2 pi y) second order.
    (The oracle needs a few seconds per solver.)"""
    from oracle import oracle
    mesh = cases.hex_case(100, 100, 100, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry"))
    n = mesh.numCells
    csr = oracle.create_csr(mesh)
    su = cases.poisson_rhs(mesh)
    su_ref = su.copy()
    a = oracle.laplacian(mesh, csr, -np.ones(n), np.zeros(mesh.numTotal), su_ref)
    fi_ref = np.zeros(mesh.numTotal)
    res0, resl, iters, _ = oracle.solve(solver, csr, a, su_ref, fi_ref, sor=1e-8, nsw=10000)
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    ctx.upload("APU", -np.ones(n))
    ctx.upload("SU", su)
    ctx.fill("PP", 0.0)
    ctx.laplacian("APU", "PP")
    rep = ctx.solve(solver, "PP", fc.solver_opts(1e-8, 10000))
    got = ctx.download("PP")[:n]
    print(f"\n[config 3] {solver}: {rep.iters} iterations (oracle {iters}), {ctx.timings().solve_ms:.1f} ms")
    assert abs(rep.iters - iters) <= 1, (rep.iters, iters)
    assert rep.res0 == pytest.approx(res0, rel=1e-11)
    assert cases.rel_l2(got, fi_ref[:n]) < 1e-6
    exact = np.sin(2 * np.pi * mesh.xc[:n]) * np.sin(2 * np.pi * mesh.yc[:n])
    assert np.max(np.abs(got - exact)) < 1e-3
    ctx.close()


def test_config4_at_108_cubed_iteration_parity_with_the_oracle(fc):
    """BASELINE config 4 at 108^3 (1.26 M cells -- what one rank of the 8-GPU 216^3 run holds): the whole `calcp`
    (assembly, DPCG to rsm < 1e-8, corrections) against the serial oracle on the same inputs: matrix and right-hand side
    bit for bit, the same iteration count (north star: within +-1), pp and p to 1e-10.  (bench.py makes the same
    comparison at 216^3 on every rank count, in its `parity` object.)"""
    from oracle import oracle, oracle_par
    n = 108
    mesh = M.hex_mesh(n, n, n)
    f = cases.config4_fields(mesh)
    csr = oracle.create_csr(mesh)
    of = oracle.Fields(mesh, csr.nnz)
    for k in ("u", "v", "w", "p", "den", "apu", "apv", "apw"):
        getattr(of, k)[:] = f[k]
    of.dPdxi[:] = oracle.grad_gauss(mesh, of.p, 1)
    kw = dict(solver="dpcg", const_mflux=True, sor=1e-8, nsw=100000, pRefCell=1, urf_p=0.3)
    oracle_par.set_threads2(1, min(8, len(os.sched_getaffinity(0))))   # row loops only: bit-identical to one thread
    rep_o = oracle.calcp(mesh, csr, of, oracle.calcp_opts(**kw))
    oracle_par.set_threads2(1, 1)
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    for k, name in (("u", "U"), ("v", "V"), ("w", "W"), ("p", "P"), ("den", "DEN"), ("apu", "APU"), ("apv", "APV"),
                    ("apw", "APW")):
        ctx.upload(name, f[k])
    ctx.grad_gauss("P", "DPDXI", 1)
    rep = ctx.calcp(fc.calcp_opts(**kw))
    print(f"\n[config 4, 108^3] DPCG {rep.rep[0].iters} iterations (oracle {rep_o.rep[0].iters})")
    assert np.array_equal(ctx.download("A"), of.a) and np.array_equal(ctx.download("SU"), of.su)
    assert abs(rep.rep[0].iters - rep_o.rep[0].iters) <= 1, (rep.rep[0].iters, rep_o.rep[0].iters)
    assert rep.rep[0].res0 == pytest.approx(rep_o.rep[0].res0, rel=1e-12)
    nc = mesh.numCells
    assert cases.rel_l2(ctx.download("PP")[:nc], of.pp[:nc]) < 1e-10
    assert cases.rel_l2(ctx.download("P")[:nc], of.p[:nc]) < 1e-10
    assert cases.rel_l2(ctx.download("U")[:nc], of.u[:nc]) < 1e-10
    ctx.close()
