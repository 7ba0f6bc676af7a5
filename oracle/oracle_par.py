"""ctypes front-end of the lock-step multi-rank oracle (TEST INFRASTRUCTURE ONLY).

src-parallel semantics (exchange.f90, global_sum_mpi.f90, the parallel twins of calcp / dpcg /
iccg / bicgstab / grad_gauss / laplacian) with the R ranks advanced inside one process; used to
check the NCCL path and the gloo tests.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np

from . import oracle as O


def set_threads(n: int) -> int:
    """Host threads for the rank loops of the lock-step solver (1 = serial).  Returns the number in effect."""
    if not O.lib().fco_par_openmp():
        return 1
    O.lib().fco_par_set_threads(int(n))
    return max(1, int(n))


def set_threads2(outer: int, inner: int):
    """`outer` host threads over the ranks x `inner` threads inside every rank's row loops (SpMV, AXPY; the inner
    products stay sequential, so all results are bit-identical to the one-thread run).  Returns (outer, inner) in
    effect."""
    if not O.lib().fco_par_openmp():
        return 1, 1
    O.lib().fco_par_set_threads2(int(outer), int(inner))
    return max(1, int(outer)), max(1, int(inner))


class FcoRank(C.Structure):
    _fields_ = [("g", O.FcoMesh), ("m", O.FcoCsr), ("f", O.FcoFields), ("apr", O.dp), ("fmpro", O.dp),
                ("numConnections", C.c_int), ("neighbProcNo", O.ip), ("neighbProcOffset", O.ip)]


class ParCase:
    """Per-rank meshes (from freecappuccino_b200.mesh.partition) + CSR + fields."""

    def __init__(self, meshes):
        self.meshes = list(meshes)
        self.nr = len(self.meshes)
        self.csr = [O.create_csr(m) for m in self.meshes]
        self.fields = [O.Fields(m, c.nnz) for m, c in zip(self.meshes, self.csr)]
        self.apr = [np.zeros(max(m.npro, 1)) for m in self.meshes]
        self.fmpro = [np.zeros(max(m.npro, 1)) for m in self.meshes]
        self._keep = []
        self.R = (FcoRank * self.nr)()
        for r, m in enumerate(self.meshes):
            ms = O.mesh_struct(m)
            self._keep.append(ms)
            nb = np.ascontiguousarray(m.neighbProcNo if m.neighbProcNo is not None else np.zeros(0), dtype=np.int32)
            off = np.ascontiguousarray(m.neighbProcOffset if m.neighbProcOffset is not None else np.ones(1),
                                       dtype=np.int32)
            self._keep += [nb, off]
            self.R[r].g = ms
            self.R[r].m = self.csr[r].c()
            self.R[r].f = self.fields[r].c()
            self.R[r].apr = O._d(self.apr[r])
            self.R[r].fmpro = O._d(self.fmpro[r])
            self.R[r].numConnections = nb.size
            self.R[r].neighbProcNo = O._i(nb)
            self.R[r].neighbProcOffset = O._i(off)

    def _ptrs(self, arrays: List[np.ndarray]):
        p = (O.dp * self.nr)()
        for r, a in enumerate(arrays):
            assert a.dtype == np.float64 and a.flags.c_contiguous
            p[r] = a.ctypes.data_as(O.dp)
        return p

    def exchange(self, arrays: List[np.ndarray], stride: int = 1):
        O.lib().fco_par_exchange(self.R, self.nr, self._ptrs(arrays), stride)

    def grad_gauss(self, phis: List[np.ndarray], nigrad: int = 1) -> List[np.ndarray]:
        out = [np.zeros((m.numCells + m.npro, 3)) for m in self.meshes]
        O.lib().fco_par_grad_gauss(self.R, self.nr, self._ptrs(phis), nigrad, self._ptrs(out))
        return out

    def laplacian(self, mus: List[np.ndarray], phis: List[np.ndarray]):
        """fills fields[r].a, apr[r]; updates fields[r].su."""
        O.lib().fco_par_laplacian(self.R, self.nr, self._ptrs(mus), self._ptrs(phis))

    def solve(self, name: str, fis: List[np.ndarray], sor: float, nsw: int, small: float = O.SMALL,
              tol: float = O.TOL, history: bool = False):
        o = O.FcoSolverOpts(sor, nsw, small, tol, 1)
        rep = O.FcoReport()
        hist = np.zeros(nsw) if history else None
        rc = O.lib().fco_par_solve(self.R, self.nr, O.SOLVERS[name], self._ptrs(fis), C.byref(o), C.byref(rep),
                                   O._d(hist))
        assert rc == 0
        return (rep, hist[:rep.iters]) if history else rep

    def calcp_assemble(self, opts: O.FcoCalcpOpts):
        O.lib().fco_par_calcp_assemble(self.R, self.nr, C.byref(opts))

    def calcp(self, opts: O.FcoCalcpOpts) -> O.FcoCalcpReport:
        rep = O.FcoCalcpReport()
        rc = O.lib().fco_par_calcp(self.R, self.nr, C.byref(opts), C.byref(rep))
        assert rc == 0
        return rep

    # ---- momentum predictor, src-parallel/calcuvw.f90 (fc_oracle_par_uvw.c) ----
    def uvw_fields(self, viscos: float = 0.0):
        """One UvwFields per rank (apu/apv/apw alias the rank's Fields arrays, numCells+npro)."""
        self.uvw = [O.UvwFields(m, f, viscos) for m, f in zip(self.meshes, self.fields)]
        self.X = (O.FcoUvw * self.nr)()
        for r, x in enumerate(self.uvw):
            self.X[r] = x.c()
        return self.uvw

    def calcuvw_assemble(self, opts: "O.FcoUvwOpts"):
        O.lib().fco_par_calcuvw_assemble(self.R, self.nr, self.X, C.byref(opts))

    def calcuvw_component(self, opts: "O.FcoUvwOpts", comp: int) -> "O.FcoReport":
        rep = O.FcoReport()
        rc = O.lib().fco_par_calcuvw_component(self.R, self.nr, self.X, C.byref(opts), comp, C.byref(rep))
        assert rc == 0, rc
        return rep

    def calcuvw(self, opts: "O.FcoUvwOpts") -> "O.FcoUvwReport":
        rep = O.FcoUvwReport()
        rc = O.lib().fco_par_calcuvw(self.R, self.nr, self.X, C.byref(opts), C.byref(rep))
        assert rc == 0, rc
        return rep

    # ---- PISO / PIMPLE on several ranks (fc_oracle_par_piso.c) ----
    def piso(self, opts: "O.FcoPisoOpts") -> "O.FcoPisoReport":
        """src-parallel/PISO_multiple_correction.f90 (PIMPLE with opts.pimple); uvw_fields() must have been called and
        fields[r].a must hold the momentum matrix calcuvw left behind (apr[r] its processor coefficients)."""
        rep = O.FcoPisoReport()
        hs = [np.zeros(c.nnz) for c in self.csr]
        rc = O.lib().fco_par_piso(self.R, self.nr, self.X, C.byref(opts), self._ptrs(hs), C.byref(rep))
        assert rc == 0, rc
        return rep

    def get_rAU_x_UEqnH(self, opts: "O.FcoPisoOpts", hs: List[np.ndarray]) -> None:
        """src-parallel/get_rAU_x_UEqnH.f90 on every rank; hs[r] = the backed-up momentum matrix of rank r."""
        O.lib().fco_par_get_rAU_x_UEqnH(self.R, self.nr, self.X, C.byref(opts), self._ptrs(hs))

    # ---- the `grad` dispatcher on several ranks (src-parallel/gradients.f90:95-160) ----
    def grad(self, phis: List[np.ndarray], method: str = "gauss", limiter: str = "no-limit", nigrad: int = 1,
             small: float = O.SMALL) -> List[np.ndarray]:
        """gauss or lstsq_qr (+ optional limiter) of every rank's field; returns the (numCells+npro, 3) gradients with
        current halos."""
        mth, lim = O.GRAD_METHODS[method], O.LIMITERS[limiter]
        Ds = [np.zeros(1) for _ in self.meshes]
        if mth == 2:
            Ds = []
            for m in self.meshes:
                ms = O.mesh_struct(m)
                D = np.zeros(18 * m.numCells)
                bad = O.lib().fco_lsq_qr_matrix(C.byref(ms), O._d(D))
                assert bad == 0, f"{bad} cells without exactly 6 neighbours"
                Ds.append(D)
        out = [np.zeros((m.numCells + m.npro, 3)) for m in self.meshes]
        rc = O.lib().fco_par_grad(self.R, self.nr, mth, lim, C.c_double(small), self._ptrs(Ds), self._ptrs(phis), nigrad,
                                  self._ptrs(out))
        assert rc == 0, rc
        return out

    def set_gradient(self, method: str = "gauss", limiter: str = "no-limit", small: float = O.SMALL) -> None:
        """The `grad` options (input file: lstsq_qr / gauss + limiter) used by calcp / piso / calcuvw of this case from
        now on; call set_gradient() again with the defaults to switch back."""
        rc = O.lib().fco_par_set_gradient(self.R, self.nr, O.GRAD_METHODS[method], O.LIMITERS[limiter], C.c_double(small))
        assert rc == 0, rc
