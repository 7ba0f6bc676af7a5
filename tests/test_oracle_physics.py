"""Pins for the parts of the oracle the reference ships no stored outputs for (assembly,
gradients, Laplacian): analytic solutions and exact-arithmetic identities."""
import os

import numpy as np
import pytest

from freecappuccino_b200 import cases, mesh as M
from oracle import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_csr_pattern_invariants():
    m = cases.skew_case(7, 6, 5)
    c = oracle.create_csr(m)
    n = m.numCells
    assert c.ioffset[0] == 1 and c.ioffset[-1] == m.nnz + 1
    for i in range(n):
        row = c.ja[c.ioffset[i] - 1:c.ioffset[i + 1] - 1]
        assert np.all(np.diff(row) > 0)                      # sorted, no duplicates
        assert c.ja[c.diag[i] - 1] == i + 1                  # diag points at (i,i)
    own, nb = m.owner[:m.numInnerFaces], m.neighbour
    assert np.array_equal(c.ja[c.icell_jcell - 1], nb)       # a(icell_jcell(f)) is a(owner, neighbour)
    assert np.array_equal(c.ja[c.jcell_icell - 1], own)


def test_poisson_second_order_convergence():
    """poisson.f90 on the unit square (n x n x 1 cells): L_inf error ~ h^2, as in
    examples/unitSquare/convergence.png (6e-3 at 40^2, 1.55e-3 at 80^2)."""
    errs = []
    for n in (20, 40, 80):
        m = M.hex_mesh(n, n, 1, (1.0, 1.0, 1.0 / n), ("wall", "wall", "wall", "wall", "symmetry", "symmetry"))
        c = oracle.create_csr(m)
        su = cases.poisson_rhs(m)
        a = oracle.laplacian(m, c, -np.ones(m.numCells), np.zeros(m.numTotal), su)
        p = np.zeros(m.numTotal)
        oracle.solve("iccg", c, a, su, p, sor=1e-12, nsw=2000)
        exact = np.sin(2 * np.pi * m.xc) * np.sin(2 * np.pi * m.yc)
        errs.append(np.max(np.abs(p[:m.numCells] - exact)))
    # second-order slope, the property examples/unitSquare/convergence.png documents (its absolute values,
    # ~6e-3 at h=1/40, are read off a picture of runs on meshes that are not shipped -- not a pin)
    assert 3.7 < errs[0] / errs[1] < 4.3 and 3.7 < errs[1] / errs[2] < 4.3
    assert errs[1] < 6.0e-3


def test_gauss_gradient_exact_for_linear_field_on_hex():
    m = cases.hex_case(6, 5, 4)
    n = m.numCells
    phi = np.zeros(m.numTotal)
    gx, gy, gz = 1.5, -2.0, 0.25
    phi[:n] = gx * m.xc + gy * m.yc + gz * m.zc
    for kind in M.KINDS:
        fs, sl = m.boundary_faces(kind), m.boundary_slots(kind)
        phi[sl] = gx * m.xf[fs] + gy * m.yf[fs] + gz * m.zf[fs]
    g = oracle.grad_gauss(m, phi, 1)
    assert np.allclose(g, [gx, gy, gz], atol=1e-12)


def test_gauss_gradient_iterations_converge_on_skewed_mesh():
    """The nigrad fixed-point passes reduce the error of a linear field's gradient on a skewed mesh."""
    m = cases.skew_case(8, 8, 8, kinds=("wall",) * 6, jitter=0.15)
    n = m.numCells
    phi = np.zeros(m.numTotal)
    phi[:n] = 1.0 * m.xc + 2.0 * m.yc - 1.0 * m.zc
    for kind in M.KINDS:
        fs, sl = m.boundary_faces(kind), m.boundary_slots(kind)
        phi[sl] = 1.0 * m.xf[fs] + 2.0 * m.yf[fs] - 1.0 * m.zf[fs]
    e = [np.abs(oracle.grad_gauss(m, phi, k) - [1.0, 2.0, -1.0]).max() for k in (1, 2, 4)]
    assert e[1] < 0.7 * e[0] and e[2] < 0.7 * e[0]   # the skewness correction pays; it then saturates


def test_laplacian_rows_sum_to_zero_away_from_walls():
    m = cases.skew_case(6, 6, 6, kinds=("symmetry",) * 6)
    c = oracle.create_csr(m)
    su = np.zeros(m.numCells)
    a = oracle.laplacian(m, c, -np.ones(m.numCells), np.zeros(m.numTotal), su)
    rows = np.add.reduceat(a, c.ioffset[:-1] - 1)
    assert np.allclose(rows, 0.0, atol=1e-10)
    # symmetric matrix: a(i,j) == a(j,i)
    assert np.array_equal(a[c.icell_jcell - 1], a[c.jcell_icell - 1])


def test_calcp_source_is_the_net_mass_flux():
    """su(P) = -sum of outgoing face fluxes: sum(su) telescopes to the boundary fluxes (calcp :56-75, 107)."""
    m = cases.skew_case()
    c = oracle.create_csr(m)
    f = cases.flow_fields(m)
    fmi, flomas = cases.inlet_fluxes(m, f)
    of = oracle.Fields(m, c.nnz)
    for k in ("u", "v", "w", "p", "den", "apu", "apv", "apw"):
        getattr(of, k)[:] = f[k]
    of.fmi[:fmi.size] = fmi
    of.dPdxi[:] = oracle.grad_gauss(m, of.p, 1)
    oracle.calcp_assemble(m, c, of, oracle.calcp_opts(flomas=flomas))
    nout = m.count("outlet")
    assert of.su.sum() == pytest.approx(-(fmi.sum() + of.fmo[:nout].sum()), abs=1e-12)
    assert of.fmo[:nout].sum() == pytest.approx(flomas, rel=1e-12)      # outlet scaled to the inflow
    rows = np.add.reduceat(of.a, c.ioffset[:-1] - 1)
    assert np.allclose(rows, 0.0, atol=1e-9 * np.abs(of.a).max())       # pure Neumann p' operator


@pytest.mark.parametrize("solver", ["dpcg", "iccg", "bicgstab"])
def test_solvers_reduce_the_true_residual(solver):
    m = cases.hex_case(10, 9, 8)
    c = oracle.create_csr(m)
    su = cases.poisson_rhs(m)
    a = oracle.laplacian(m, c, -np.ones(m.numCells), np.zeros(m.numTotal), su)
    x = np.zeros(m.numTotal)
    res0, resl, iters, res = oracle.solve(solver, c, a, su, x, sor=1e-10, nsw=1000)
    true = su - oracle.spmv(c, a, x)
    assert np.abs(true).sum() < 1e-9 * res0
    assert np.allclose(res, true, atol=1e-12 * res0)


@pytest.mark.parametrize("case,cells,inner,faces", [("cavity", 400, 760, 1640), ("pitzDaily", 12225, 24170, 49180)])
def test_shipped_example_meshes(case, cells, inner, faces):
    """Configs 1 and 2: geometry of the reference's example meshes (sizes of SURVEY 8; closed cells)."""
    m = cases.golden_mesh(os.path.join(GOLD, case + ".npz"))
    assert (m.numCells, m.numInnerFaces, m.numFaces) == (cells, inner, faces)
    assert np.all(m.vol > 0) and np.all((m.facint > 0) & (m.facint < 1))
    # every cell is closed: sum of outward area vectors vanishes
    s = np.zeros((m.numCells, 3))
    ar = np.stack([m.arx, m.ary, m.arz], axis=1)
    np.add.at(s, m.owner - 1, ar)
    np.add.at(s, m.neighbour - 1, -ar[:m.numInnerFaces])
    assert np.abs(s).max() < 1e-12 * np.abs(ar).max() * 10
    c = oracle.create_csr(m)
    assert c.nnz == cells + 2 * inner
