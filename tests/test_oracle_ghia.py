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
