"""CPU checks of the oracle's least-squares gradients and slope limiters."""
import numpy as np
import pytest

from freecappuccino_b200 import cases
from freecappuccino_b200 import mesh as M
from oracle import oracle


def linear_field(mesh, a=(0.3, -0.7, 1.1), c=0.25):
    nt, n = mesh.numTotal, mesh.numCells
    x, y, z = np.zeros(nt), np.zeros(nt), np.zeros(nt)
    x[:n], y[:n], z[:n] = mesh.xc[:n], mesh.yc[:n], mesh.zc[:n]
    for kind in M.KINDS:
        fs, sl = mesh.boundary_faces(kind), mesh.boundary_slots(kind)
        x[sl], y[sl], z[sl] = mesh.xf[fs], mesh.yf[fs], mesh.zf[fs]
    return c + a[0] * x + a[1] * y + a[2] * z


@pytest.fixture(autouse=True)
def reset_gradient():
    yield
    oracle.set_gradient("gauss", "no-limit")


@pytest.mark.parametrize("make", [lambda: cases.hex_case(6, 5, 4, kinds=("inlet", "outlet", "wall", "wall", "symmetry", "prOutlet")),
                                  lambda: cases.skew_case(6, 5, 4)])
def test_lsq_qr_is_exact_for_linear_fields(make):
    mesh = make()
    D, bad = oracle.lsq_qr_matrix(mesh)
    assert bad == 0
    g = oracle.grad_lsq_qr(mesh, D, linear_field(mesh))
    n = mesh.numCells
    assert np.allclose(g[:n], np.array([0.3, -0.7, 1.1]), rtol=0, atol=1e-11)


def test_lsq_qr_matrix_is_the_pseudo_inverse():
    """R^-1 Q^T of the 6 x 3 distance matrix = (D^T D)^-1 D^T, whatever QR algorithm produced it (the reference
    calls LAPACK DGEQRF; the oracle restates DGEQR2)."""
    mesh = cases.skew_case(5, 4, 3)
    Dq, bad = oracle.lsq_qr_matrix(mesh)
    assert bad == 0
    n, F = mesh.numCells, mesh.numInnerFaces
    rows = [[] for _ in range(n)]
    for f in range(F):
        p, q = mesh.owner[f] - 1, mesh.neighbour[f] - 1
        rows[p].append([mesh.xc[q] - mesh.xc[p], mesh.yc[q] - mesh.yc[p], mesh.zc[q] - mesh.zc[p]])
        rows[q].append([mesh.xc[p] - mesh.xc[q], mesh.yc[p] - mesh.yc[q], mesh.zc[p] - mesh.zc[q]])
    for kind in M.KINDS:
        for f in mesh.boundary_faces(kind):
            p = mesh.owner[f] - 1
            rows[p].append([mesh.xf[f] - mesh.xc[p], mesh.yf[f] - mesh.yc[p], mesh.zf[f] - mesh.zc[p]])
    for c in range(n):
        A = np.array(rows[c])
        assert np.allclose(Dq[c].T, np.linalg.pinv(A), rtol=1e-9, atol=1e-9 * np.abs(np.linalg.pinv(A)).max())


def test_poly_cells_are_outside_lsq_qr():
    D, bad = oracle.lsq_qr_matrix(cases.poly_case(3))
    assert bad > 0      # 14-faced cells: D(3,6,numCells) cannot hold them (grad_lsq_qr.f90:31)


@pytest.mark.parametrize("weighted", [False, True])
def test_lsq_normal_equations_x_and_z_exact_y_as_written(weighted):
    """grad_lsq / grad_lsq_dm: the x and z components reproduce a linear field; the y component is what
    grad_lsq.f90:303 computes, b1*Dmat(4) - b2*Dmat(5) - b3*Dmat(6), which is not the least-squares one."""
    mesh = cases.hex_case(5, 4, 3, lengths=(1.0, 0.7, 0.4))
    fi = linear_field(mesh)
    dmat = oracle.lsq_matrix(mesh, weighted)
    g = oracle.grad_lsq(mesh, weighted, dmat, fi)
    n = mesh.numCells
    if not weighted:
        assert np.allclose(g[:n, 0], 0.3, atol=1e-11) and np.allclose(g[:n, 2], 1.1, atol=1e-11)
    # orthogonal mesh: the normal matrix is diagonal, so as written y = b1 / d22 = a_x d11 / d22 (unweighted)
    assert not np.allclose(g[:n, 1], -0.7, atol=1e-3)
    if not weighted:
        hx, hy = 1.0 / 5, 0.7 / 4
        interior = np.ones(n, bool)        # a cell away from all boundaries: d11 = 2 hx^2, d22 = 2 hy^2
        ijk = np.indices((3, 4, 5)).reshape(3, -1)
        interior = (ijk[2] > 0) & (ijk[2] < 4) & (ijk[1] > 0) & (ijk[1] < 3) & (ijk[0] > 0) & (ijk[0] < 2)
        assert np.allclose(g[:n, 1][interior], 0.3 * hx * hx / (hy * hy), rtol=1e-10)


@pytest.mark.parametrize("which", ["Barth-Jespersen", "Venkatakrishnan", "mVenkatakrishnan"])
def test_limiters_scale_the_gradient_by_a_factor_in_0_1(which):
    mesh = cases.skew_case(6, 5, 4)
    csr = oracle.create_csr(mesh)
    phi = cases.flow_fields(mesh)["p"]
    g0 = oracle.grad_gauss(mesh, phi, 1)
    g = g0.copy()
    oracle.slope_limiter(mesh, csr, which, phi, g)
    n = mesh.numCells
    big = np.abs(g0[:n]).max(axis=1) > 1e-12
    k = np.argmax(np.abs(g0[:n]), axis=1)
    s = g[np.arange(n), k][big] / g0[np.arange(n), k][big]
    lo = -1e-12 if which == "mVenkatakrishnan" else -np.inf   # B-J / Venkatakrishnan as written have no lower clamp
    assert np.all(s <= 1.0 + 1e-12) and np.all(s >= lo)
    assert np.allclose(g[:n][big], g0[:n][big] * s[:, None], rtol=1e-12, atol=1e-14)
    assert (s < 0.999).any()          # something is actually limited


def test_dispatcher_drives_calcp():
    """With lstsq_qr + mVenkatakrishnan configured (the pitzDaily `input`), calcp's velocity gradients are the
    limited least-squares ones; with the default they are the Gauss ones."""
    mesh = cases.skew_case(6, 5, 4)
    f = cases.flow_fields(mesh)
    csr = oracle.create_csr(mesh)

    def run():
        of = oracle.Fields(mesh, csr.nnz)
        for k in ("u", "v", "w", "p", "den", "apu", "apv", "apw"):
            getattr(of, k)[:] = f[k]
        oracle.calcp_assemble(mesh, csr, of, oracle.calcp_opts(const_mflux=True))
        return of

    a = run()
    assert np.array_equal(a.dUdxi, oracle.grad_gauss(mesh, f["u"], 1))
    oracle.set_gradient("lstsq_qr", "mVenkatakrishnan", mesh)
    b = run()
    D, _ = oracle.lsq_qr_matrix(mesh)
    g = oracle.grad_lsq_qr(mesh, D, f["u"])
    oracle.slope_limiter(mesh, csr, "mVenkatakrishnan", f["u"], g)
    assert np.array_equal(b.dUdxi, g)
    assert not np.array_equal(a.su, b.su)


def test_simple_loop_with_lsq_qr_gradients_reproduces_ghia():
    """The QR least-squares gradient (DGEQR2 restatement) drives the whole SIMPLE loop to the same benchmark solution
    as the Gauss gradient; the normal-equation variants, whose y-component is computed as written in
    grad_lsq.f90:303, do not -- evidence that the restated quirk is what the formula does, not a transcription slip."""
    import test_oracle_ghia as G
    err = {}
    for method in ("lstsq_qr", "lstsq"):
        mesh = G.cavity_mesh(20)
        oracle.set_gradient(method, "no-limit", mesh)
        mesh, of, it = G.simple_to_convergence(20)
        err[method], _ = G.centreline_error(mesh, 20, of.u)
        oracle.set_gradient("gauss", "no-limit")
    assert err["lstsq_qr"] < 0.025
    assert err["lstsq"] > 0.05


# ---- several ranks: the dispatcher of src-parallel/gradients.f90 (fco_par_grad) ----
def _pitz_two_ranks():
    import os
    from freecappuccino_b200 import mesh as M
    gold = os.path.join(os.path.dirname(__file__), "golden")
    g = cases.golden_mesh(os.path.join(gold, "pitzDaily.npz"))
    cr = np.load(os.path.join(gold, "pitzDaily_par_cells.npz"))["cell_rank"].astype(np.int64)
    return g, M.partition(g, cr, 2)


def _smooth(g):
    n = g.numCells
    x = np.zeros(g.numTotal)
    from freecappuccino_b200 import mesh as M
    x[:n] = np.sin(40 * g.xc[:n]) + 30 * g.yc[:n]
    for k in M.KINDS:
        fs, sl = g.boundary_faces(k), g.boundary_slots(k)
        x[sl] = np.sin(40 * g.xf[fs]) + 30 * g.yf[fs]
    return x


def test_two_rank_lstsq_qr_gradient_is_the_serial_one():
    """The shipped 2-rank pitzDaily decomposition: a cell next to the cut has the same six neighbours as in the serial
    mesh (the halo cell instead of the remote one), only in another order -- the least-squares gradient is the same
    up to round-off; the halo copies of the gradient are the neighbour rank's values."""
    from freecappuccino_b200 import mesh as M
    from oracle import oracle_par
    g, parts = _pitz_two_ranks()
    x = _smooth(g)
    csr = oracle.create_csr(g)
    D, bad = oracle.lsq_qr_matrix(g)
    assert bad == 0
    ref = oracle.grad_lsq_qr(g, D, x)
    pc = oracle_par.ParCase(parts)
    out = pc.grad([M.scatter_total(g, p, x) for p in parts], "lstsq_qr")
    got = M.gather_cells(g, parts, [o[:p.numCells] for o, p in zip(out, parts)])
    assert np.allclose(got, ref[:g.numCells], rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    for o, p in zip(out, parts):
        assert np.allclose(o[p.numCells:], ref[p.halo_global], rtol=1e-9, atol=1e-9 * np.abs(ref).max())


@pytest.mark.parametrize("limiter", ["Barth-Jespersen", "Venkatakrishnan", "mVenkatakrishnan"])
def test_two_rank_limiters_only_shorten_the_gradient(limiter):
    from freecappuccino_b200 import mesh as M
    from oracle import oracle_par
    g, parts = _pitz_two_ranks()
    x = _smooth(g)
    pc = oracle_par.ParCase(parts)
    phis = [M.scatter_total(g, p, x) for p in parts]
    free = pc.grad([p.copy() for p in phis], "gauss")
    lim = pc.grad([p.copy() for p in phis], "gauss", limiter)
    for a, b, p in zip(free, lim, parts):
        n = p.numCells
        na, nb_ = np.linalg.norm(a[:n], axis=1), np.linalg.norm(b[:n], axis=1)
        assert np.all(nb_ <= na * (1 + 1e-12))
        cos = np.einsum("ij,ij->i", a[:n], b[:n])
        assert np.all(cos >= -1e-12 * na * na)          # same direction: a scalar factor in [0, 1]
        if limiter == "mVenkatakrishnan":   # the other two compare with the GLOBAL extrema and rarely bite on a smooth field
            assert np.any(nb_ < 0.999 * na)
