"""Physical pin of the whole SIMPLE loop (calcuvw + calcp with their gradients, BiCGStab and ICCG): the lid-driven
cavity at Re = 100 against the benchmark table of Ghia, Ghia & Shin, J. Comput. Phys. 48 (1982), Table I
(u along the vertical line through the geometric centre).  The reference stores no outputs of these routines; a
published benchmark is the strongest check available for the restatement.  CPU only (the GPU twin is in
tests/test_gpu_zz1_momentum.py)."""
import numpy as np

from freecappuccino_b200 import mesh as M
from oracle import oracle

GHIA_Y = np.array([0.9766, 0.9688, 0.9609, 0.9531, 0.8516, 0.7344, 0.6172, 0.5, 0.4531, 0.2813, 0.1719, 0.1016, 0.0703,
                   0.0625, 0.0547])
GHIA_U = np.array([0.84123, 0.78871, 0.73722, 0.68717, 0.23151, 0.00332, -0.13641, -0.20581, -0.21090, -0.15662,
                   -0.10150, -0.06434, -0.04775, -0.04192, -0.03717])


def cavity_mesh(n):
    return M.hex_mesh(n, n, 1, (1.0, 1.0, 1.0 / n), ("wall", "wall", "wall", "wall", "symmetry", "symmetry"))


def lid_slots(mesh):
    fs, sl = mesh.boundary_faces("wall"), mesh.boundary_slots("wall")
    return sl[mesh.ary[fs] > 0.0]


def centreline_error(mesh, n, u):
    nc = mesh.numCells
    U = u[:nc].reshape(n, n)
    yc = mesh.yc[:nc].reshape(n, n)[:, 0]
    ucl = 0.5 * (U[:, n // 2 - 1] + U[:, n // 2])
    return float(np.abs(np.interp(GHIA_Y, yc, ucl) - GHIA_U).max()), float(ucl.min())


def simple_to_convergence(n, re=100.0, sormax=1e-7, maxit=3000):
    mesh = cavity_mesh(n)
    csr = oracle.create_csr(mesh)
    of = oracle.Fields(mesh, csr.nnz)
    of.u[lid_slots(mesh)] = 1.0
    nu = 1.0 / re
    x = oracle.UvwFields(mesh, of, nu)
    oo = oracle.uvw_opts(scheme="muscl-f", urf=(0.7,) * 3, sor=(1e-2,) * 3, nsw=(20,) * 3, bdf=True, btime=0.0,
                         timestep=1e20, viscos=nu)     # the shipped cavity settings, examples/cavity/input
    po = oracle.calcp_opts(solver="iccg", sor=1e-2, nsw=100, urf_p=0.3, pRefCell=1)
    for it in range(1, maxit + 1):
        ru = oracle.calcuvw(mesh, csr, of, x, oo)
        rp = oracle.calcp(mesh, csr, of, po)
        if max(ru.rep[0].res0, ru.rep[1].res0, rp.rep[0].res0) < sormax:    # main.f90:157-169
            break
    return mesh, of, it


def test_lid_driven_cavity_re100_matches_ghia():
    errs = {}
    for n in (20, 40):
        mesh, of, it = simple_to_convergence(n)
        assert it < 3000
        errs[n], umin = centreline_error(mesh, n, of.u)
        assert abs(umin - (-0.2109)) < (0.01 if n == 20 else 0.002)
    assert errs[20] < 0.02 and errs[40] < 0.006
    assert errs[20] / errs[40] > 3.0        # second-order convergence towards the benchmark


def poiseuille(nx, ny, maxit=4000):
    """Plane Poiseuille flow, Re = 10: parabolic inlet profile, outlet, two walls.  Exercises the inlet / outlet
    branches of calcuvw (facefluxuvw_boundary) and calcp (adjustMassFlow, outlet extrapolation and scaling)."""
    lx, h, nu = 3.0, 1.0, 0.1
    mesh = M.hex_mesh(nx, ny, 1, (lx, h, h / ny), ("inlet", "outlet", "wall", "wall", "symmetry", "symmetry"))
    nc = mesh.numCells
    csr = oracle.create_csr(mesh)
    of = oracle.Fields(mesh, csr.nnz)
    fs, sl = mesh.boundary_faces("inlet"), mesh.boundary_slots("inlet")
    of.u[sl] = 6.0 * mesh.yf[fs] * (1.0 - mesh.yf[fs])
    of.u[:nc] = 1.0
    of.fmi[:len(fs)] = of.den[sl] * of.u[sl] * mesh.arx[fs]          # bcin.f90: fmi = den (U . S), negative = inflow
    flomas = float(-of.fmi[:len(fs)].sum())
    F = mesh.numInnerFaces
    o, nb = mesh.owner[:F] - 1, mesh.neighbour - 1
    of.flmass[:] = 0.5 * (of.u[o] + of.u[nb]) * mesh.arx[:F]
    x = oracle.UvwFields(mesh, of, nu)
    oo = oracle.uvw_opts(scheme="muscl-f", urf=(0.7,) * 3, sor=(1e-2,) * 3, nsw=(20,) * 3, bdf=True, timestep=1e20, viscos=nu)
    po = oracle.calcp_opts(solver="iccg", sor=1e-2, nsw=100, urf_p=0.3, pRefCell=1, flomas=flomas)
    for it in range(1, maxit + 1):
        ru = oracle.calcuvw(mesh, csr, of, x, oo)
        rp = oracle.calcp(mesh, csr, of, po)
        if max(ru.rep[0].res0, ru.rep[1].res0, rp.rep[0].res0) < 1e-8:
            break
    U, P = of.u[:nc].reshape(ny, nx), of.p[:nc].reshape(ny, nx)
    yc, xc = mesh.yc[:nc].reshape(ny, nx)[:, 0], mesh.xc[:nc].reshape(ny, nx)[0, :]
    i = nx // 2
    err = float(np.abs(U[:, i] - 6.0 * yc * (1.0 - yc)).max())
    dpdx = float((P[ny // 2, i + 2] - P[ny // 2, i - 2]) / (xc[i + 2] - xc[i - 2]))
    return it, err, dpdx


def test_plane_poiseuille_flow_is_second_order():
    """Analytic solution: u = 6 y (1 - y), dp/dx = -12 nu U / H^2 = -1.2."""
    it1, e1, g1 = poiseuille(15, 10)
    it2, e2, g2 = poiseuille(30, 20)
    assert it1 < 4000 and it2 < 4000
    assert e1 < 0.025 and e2 < 0.007 and e1 / e2 > 3.0
    assert abs(g1 + 1.2) < 0.04 and abs(g2 + 1.2) < 0.01
