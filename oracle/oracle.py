"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY, see fc_oracle.h).

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference
legs of bench.py may import this module.  The product package
``freecappuccino_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[C.CDLL] = None

SMALL = float(np.float32(1e-20))      # parameters: small = 1e-20 (single literal), modules_allocatable.f90:27
SMALL_TEST = float(np.float32(1e-30)) # tests/test_sparse_solvers.f90:36
TOL = float(np.float32(1e-13))        # dpcg.f90:37


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libfc_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("fc_oracle.c", "fc_oracle_uvw.c", "fc_oracle_piso.c", "fc_oracle_grad.c", "fc_oracle_par.c", "fc_oracle_par_uvw.c",
                                           "fc_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libfc_oracle.so"])
    return so


dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class FcoMesh(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "numCells", "numInnerFaces", "numFaces", "numTotal", "npro",
        "ninl", "nout", "nsym", "nwal", "npru", "noc",
        "iProcFacesStart", "iInletFacesStart", "iOutletFacesStart", "iSymmetryFacesStart",
        "iWallFacesStart", "iPressOutletFacesStart", "iOCFacesStart")] + [
        ("owner", ip), ("neighbour", ip),
        ("xc", dp), ("yc", dp), ("zc", dp), ("vol", dp),
        ("arx", dp), ("ary", dp), ("arz", dp), ("xf", dp), ("yf", dp), ("zf", dp),
        ("facint", dp), ("fpro", dp),
        ("ijl", ip), ("ijr", ip), ("ijlFace", ip), ("foc", dp)]


class FcoCsr(C.Structure):
    _fields_ = [("n", C.c_int), ("nnz", C.c_int), ("ioffset", ip), ("ja", ip), ("diag", ip),
                ("icell_jcell", ip), ("jcell_icell", ip)]


class FcoSolverOpts(C.Structure):
    _fields_ = [("sor", C.c_double), ("nsw", C.c_int), ("small", C.c_double), ("tol", C.c_double),
                ("parallel", C.c_int)]


class FcoReport(C.Structure):
    _fields_ = [("res0", C.c_double), ("resl", C.c_double), ("iters", C.c_int)]


class FcoStrips(C.Structure):
    _fields_ = [("noc", C.c_int), ("ijl", ip), ("ijr", ip), ("al", dp), ("ar", dp),
                ("npro", C.c_int), ("pown", ip), ("apr", dp), ("iProcStart", C.c_int)]


class FcoFields(C.Structure):
    _fields_ = [(n, dp) for n in (
        "u", "v", "w", "p", "pp", "den", "flmass", "fmi", "fmo", "fmoc",
        "dUdxi", "dVdxi", "dWdxi", "dPdxi", "apu", "apv", "apw", "a", "su", "res", "al", "ar")]


class FcoCalcpOpts(C.Structure):
    _fields_ = [("npcor", C.c_int), ("nigrad", C.c_int), ("nipgrad", C.c_int), ("pRefCell", C.c_int),
                ("urf_p", C.c_double), ("solver", C.c_int), ("const_mflux", C.c_int), ("flomas", C.c_double),
                ("lsq_flag", C.c_int), ("flux_variant", C.c_int), ("sol", FcoSolverOpts)]


class FcoCalcpReport(C.Structure):
    _fields_ = [("rep", FcoReport * 8), ("sumLocalContErr", C.c_double), ("globalContErr", C.c_double)]


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.fco_create_csr.restype = C.c_int
        for f in ("fco_dpcg", "fco_iccg", "fco_bicgstab", "fco_calcp", "fco_calcuvw", "fco_calcuvw_assemble",
                  "fco_calcuvw_component", "fco_piso", "fco_lsq_qr_matrix"):
            getattr(_LIB, f).restype = C.c_int
    return _LIB


def _d(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(dp)


def _i(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(ip)


SOLVERS = {"dpcg": 0, "iccg": 1, "bicgstab": 2}


class Csr:
    """Integer arrays of ``create_CSR_matrix_from_mesh_data`` (1-based)."""

    def __init__(self, ioffset, ja, diag, icell_jcell=None, jcell_icell=None):
        self.ioffset, self.ja, self.diag = ioffset, ja, diag
        self.icell_jcell, self.jcell_icell = icell_jcell, jcell_icell
        self.n = ioffset.size - 1
        self.nnz = ja.size

    def c(self) -> FcoCsr:
        return FcoCsr(self.n, self.nnz, _i(self.ioffset), _i(self.ja), _i(self.diag),
                      _i(self.icell_jcell), _i(self.jcell_icell))


def create_csr(mesh) -> Csr:
    n, F = mesh.numCells, mesh.numInnerFaces
    nnz = n + 2 * F
    ioffset = np.zeros(n + 1, np.int32)
    ja = np.zeros(nnz, np.int32)
    diag = np.zeros(n, np.int32)
    ij = np.zeros(F, np.int32)
    ji = np.zeros(F, np.int32)
    rc = lib().fco_create_csr(n, F, _i(mesh.owner), _i(mesh.neighbour), _i(ioffset), _i(ja), _i(diag), _i(ij), _i(ji))
    assert rc == 0
    return Csr(ioffset, ja, diag, ij, ji)


def mesh_struct(m) -> FcoMesh:
    keep = []
    def d(a):
        a = np.ascontiguousarray(a, dtype=np.float64) if a is not None else None
        keep.append(a)
        return _d(a)
    s = FcoMesh()
    s.numCells, s.numInnerFaces, s.numFaces, s.numTotal, s.npro = m.numCells, m.numInnerFaces, m.numFaces, m.numTotal, m.npro
    s.ninl, s.nout, s.nsym, s.nwal, s.npru, s.noc = (m.count("inlet"), m.count("outlet"), m.count("symmetry"),
                                                   m.count("wall"), m.count("prOutlet"), m.noc)
    s.iProcFacesStart = m.iProcFacesStart
    s.iInletFacesStart, s.iOutletFacesStart = m.faces_start("inlet"), m.faces_start("outlet")
    s.iSymmetryFacesStart, s.iWallFacesStart = m.faces_start("symmetry"), m.faces_start("wall")
    s.iPressOutletFacesStart, s.iOCFacesStart = m.faces_start("prOutlet"), m.iOCFacesStart
    s.owner, s.neighbour = _i(m.owner), _i(m.neighbour)
    s.xc, s.yc, s.zc, s.vol = d(m.xc), d(m.yc), d(m.zc), d(m.vol)
    s.arx, s.ary, s.arz, s.xf, s.yf, s.zf = d(m.arx), d(m.ary), d(m.arz), d(m.xf), d(m.yf), d(m.zf)
    s.facint, s.fpro = d(m.facint), d(m.fpro)
    s.ijl, s.ijr, s.ijlFace, s.foc = _i(m.ijl), _i(m.ijr), _i(m.ijlFace), d(m.foc)
    s._keep = keep
    return s


def spmv(csr: Csr, a: np.ndarray, x: np.ndarray) -> np.ndarray:
    y = np.zeros(csr.n)
    cs = csr.c()
    lib().fco_spmv(C.byref(cs), _d(a), _d(x), _d(y))
    return y


def solve(name: str, csr: Csr, a: np.ndarray, su: np.ndarray, fi: np.ndarray, sor: float, nsw: int,
          small: float = SMALL, tol: float = TOL, parallel: bool = False, history: bool = False):
    """Run dpcg / iccg / bicgstab in place on ``fi``.  Returns (res0, resl, iters, res[, hist])."""
    res = np.zeros(csr.n)
    hist = np.zeros(nsw) if history else None
    o = FcoSolverOpts(sor, nsw, small, tol, int(parallel))
    rep = FcoReport()
    cs = csr.c()
    fn = getattr(lib(), "fco_" + name)
    rc = fn(C.byref(cs), _d(a), _d(su), _d(fi), _d(res), None, C.byref(o), C.byref(rep), _d(hist))
    assert rc == 0
    out = (rep.res0, rep.resl, rep.iters, res)
    return out + (hist[:rep.iters],) if history else out


def laplacian(mesh, csr: Csr, mu: np.ndarray, phi: np.ndarray, su: np.ndarray) -> np.ndarray:
    """``call laplacian(mu,phi)``: returns ``a``; updates ``su`` in place (wall BC)."""
    a = np.zeros(csr.nnz)
    ms, cs = mesh_struct(mesh), csr.c()
    lib().fco_laplacian(C.byref(ms), C.byref(cs), _d(mu), _d(phi), _d(a), _d(su), None, None)
    return a


def grad_gauss(mesh, u: np.ndarray, nigrad: int = 1) -> np.ndarray:
    """Returns dPhidxi as an [numCells(+npro), 3] C array == Fortran (3,numPCells)."""
    g = np.zeros((mesh.numCells + mesh.npro, 3))
    ms = mesh_struct(mesh)
    lib().fco_grad_gauss(C.byref(ms), _d(u), nigrad, _d(g))
    return g


def grad_gauss_corrected(mesh, u: np.ndarray, seed: np.ndarray) -> np.ndarray:
    g = np.ascontiguousarray(seed, dtype=np.float64).copy()
    ms = mesh_struct(mesh)
    lib().fco_grad_gauss_corrected(C.byref(ms), _d(u), _d(g))
    return g


def bpres(mesh, p: np.ndarray, dPdxi: np.ndarray, istage: int) -> None:
    ms = mesh_struct(mesh)
    lib().fco_bpres(C.byref(ms), _d(p), _d(dPdxi), istage)


class Fields:
    """The arrays of ``module variables`` / ``sparse_matrix`` calcp touches."""

    NAMES = ("u", "v", "w", "p", "pp", "den", "flmass", "fmi", "fmo", "fmoc",
             "dUdxi", "dVdxi", "dWdxi", "dPdxi", "apu", "apv", "apw", "a", "su", "res", "al", "ar")

    def __init__(self, mesh, nnz: int):
        nt, n = mesh.numTotal, mesh.numCells
        z = np.zeros
        self.u, self.v, self.w, self.p, self.pp = z(nt), z(nt), z(nt), z(nt), z(nt)
        self.den = np.ones(nt)
        self.flmass = z(mesh.numInnerFaces)
        self.fmi, self.fmo, self.fmoc = z(max(mesh.count("inlet"), 1)), z(max(mesh.count("outlet"), 1)), z(max(mesh.noc, 1))
        npc = n + mesh.npro   # dPhidxi(3,numPCells) in src-parallel
        self.dUdxi, self.dVdxi, self.dWdxi, self.dPdxi = z((npc, 3)), z((npc, 3)), z((npc, 3)), z((npc, 3))
        self.apu, self.apv, self.apw = z(n + mesh.npro), z(n + mesh.npro), z(n + mesh.npro)
        self.a, self.su, self.res = z(nnz), z(n), z(n)
        self.al, self.ar = z(max(mesh.noc, 1)), z(max(mesh.noc, 1))

    def c(self) -> FcoFields:
        return FcoFields(*[_d(getattr(self, k)) for k in self.NAMES])

    def copy(self) -> "Fields":
        import copy
        return copy.deepcopy(self)


def calcp_opts(npcor=1, nigrad=1, pRefCell=1, urf_p=0.3, solver="iccg", const_mflux=False, flomas=0.0,
               lsq_flag=False, flux_variant=0, sor=1e-2, nsw=100, small=SMALL, tol=TOL) -> FcoCalcpOpts:
    return FcoCalcpOpts(npcor, nigrad, 2, pRefCell, urf_p, SOLVERS[solver], int(const_mflux), flomas,
                        int(lsq_flag), flux_variant, FcoSolverOpts(sor, nsw, small, tol, 0))


def calcp_assemble(mesh, csr: Csr, f: Fields, opts: FcoCalcpOpts) -> None:
    ms, cs, fs = mesh_struct(mesh), csr.c(), f.c()
    lib().fco_calcp_assemble(C.byref(ms), C.byref(cs), C.byref(fs), C.byref(opts))


def calcp(mesh, csr: Csr, f: Fields, opts: FcoCalcpOpts) -> FcoCalcpReport:
    ms, cs, fs = mesh_struct(mesh), csr.c(), f.c()
    rep = FcoCalcpReport()
    rc = lib().fco_calcp(C.byref(ms), C.byref(cs), C.byref(fs), C.byref(opts), C.byref(rep))
    assert rc == 0
    return rep


# ---- momentum predictor calcuvw (fc_oracle_uvw.c) ----
SCHEMES = {"central": (0, 7), "cds-corrected": (1, 7), "central-f": (2, 7), "linear-f": (3, 7), "muscl-f": (4, 7),
           "smart": (5, 0), "avl-smart": (5, 1), "muscl": (5, 2), "umist": (5, 3), "koren": (5, 4), "charm": (5, 5),
           "ospre": (5, 6), "linear": (5, 7)}   # read_input.f90:97-133 -> (face_value branch, limiter)


class FcoUvw(C.Structure):
    _fields_ = [(n, dp) for n in ("vis", "uo", "vo", "wo", "uoo", "voo", "woo", "t", "sv", "sw", "spu", "spv", "sp",
                                  "apu", "apv", "apw")]


class FcoUvwOpts(C.Structure):
    _fields_ = [("nigrad", C.c_int), ("nipgrad", C.c_int), ("scheme", C.c_int), ("limiter", C.c_int),
                ("gds", C.c_double), ("urf", C.c_double * 3), ("sor", C.c_double * 3), ("nsw", C.c_int * 3),
                ("bdf", C.c_int), ("btime", C.c_double), ("timestep", C.c_double), ("cn", C.c_int),
                ("const_mflux", C.c_int), ("gradPcmf", C.c_double),
                ("lbuoy", C.c_int), ("boussinesq", C.c_int), ("beta", C.c_double), ("tref", C.c_double),
                ("densit", C.c_double), ("gravx", C.c_double), ("gravy", C.c_double), ("gravz", C.c_double),
                ("viscos", C.c_double), ("sol", FcoSolverOpts)]


class FcoUvwReport(C.Structure):
    _fields_ = [("rep", FcoReport * 3)]


class UvwFields:
    """The extra arrays of ``module variables`` / ``sparse_matrix`` calcuvw touches (vis, old time levels,
    sv, sw, spu, spv, sp); apu/apv/apw alias the arrays of a ``Fields`` so that calcp sees the update."""

    NAMES = ("vis", "uo", "vo", "wo", "uoo", "voo", "woo", "t", "sv", "sw", "spu", "spv", "sp", "apu", "apv", "apw")

    def __init__(self, mesh, f: "Fields", viscos: float = 0.0):
        nt, n = mesh.numTotal, mesh.numCells
        z = np.zeros
        self.vis = np.full(nt, float(viscos))
        self.uo, self.vo, self.wo, self.uoo, self.voo, self.woo = z(nt), z(nt), z(nt), z(nt), z(nt), z(nt)
        self.t = z(nt)
        self.sv, self.sw, self.spu, self.spv, self.sp = z(n), z(n), z(n), z(n), z(n)
        self.apu, self.apv, self.apw = f.apu, f.apv, f.apw

    def c(self) -> FcoUvw:
        return FcoUvw(*[_d(getattr(self, k)) for k in self.NAMES])


def uvw_opts(scheme="muscl-f", gds=1.0, urf=(0.7, 0.7, 0.7), sor=(1e-2, 1e-2, 1e-2), nsw=(20, 20, 20), nigrad=1,
             bdf=False, btime=0.0, timestep=1e20, cn=False, const_mflux=False, gradPcmf=0.0, lbuoy=False,
             boussinesq=True, beta=0.0, tref=0.0, densit=1.0, grav=(0.0, 0.0, 0.0), viscos=0.01, small=SMALL,
             tol=TOL) -> FcoUvwOpts:
    sc, lim = SCHEMES[scheme]
    return FcoUvwOpts(nigrad, 2, sc, lim, gds, (C.c_double * 3)(*urf), (C.c_double * 3)(*sor), (C.c_int * 3)(*nsw),
                      int(bdf), btime, timestep, int(cn), int(const_mflux), gradPcmf, int(lbuoy), int(boussinesq),
                      beta, tref, densit, grav[0], grav[1], grav[2], viscos, FcoSolverOpts(0.0, 0, small, tol, 0))


def calcuvw_assemble(mesh, csr: Csr, f: Fields, x: UvwFields, opts: FcoUvwOpts) -> None:
    ms, cs, fs, xs = mesh_struct(mesh), csr.c(), f.c(), x.c()
    rc = lib().fco_calcuvw_assemble(C.byref(ms), C.byref(cs), C.byref(fs), C.byref(xs), C.byref(opts))
    assert rc == 0, rc


def calcuvw_component(mesh, csr: Csr, f: Fields, x: UvwFields, opts: FcoUvwOpts, comp: int) -> FcoReport:
    ms, cs, fs, xs = mesh_struct(mesh), csr.c(), f.c(), x.c()
    rep = FcoReport()
    rc = lib().fco_calcuvw_component(C.byref(ms), C.byref(cs), C.byref(fs), C.byref(xs), C.byref(opts), comp,
                                     C.byref(rep))
    assert rc == 0, rc
    return rep


def calcuvw(mesh, csr: Csr, f: Fields, x: UvwFields, opts: FcoUvwOpts) -> FcoUvwReport:
    ms, cs, fs, xs = mesh_struct(mesh), csr.c(), f.c(), x.c()
    rep = FcoUvwReport()
    rc = lib().fco_calcuvw(C.byref(ms), C.byref(cs), C.byref(fs), C.byref(xs), C.byref(opts), C.byref(rep))
    assert rc == 0, rc
    return rep


# ---- PISO / PIMPLE pressure equation (fc_oracle_piso.c) ----
class FcoPisoOpts(C.Structure):
    _fields_ = [("ncorr", C.c_int), ("npcor", C.c_int), ("nigrad", C.c_int), ("nipgrad", C.c_int),
                ("pRefCell", C.c_int), ("pimple", C.c_int), ("urf_p", C.c_double), ("const_mflux", C.c_int),
                ("flomas", C.c_double), ("bdf", C.c_int), ("btime", C.c_double), ("timestep", C.c_double),
                ("cn", C.c_int), ("lbuoy", C.c_int), ("boussinesq", C.c_int), ("beta", C.c_double),
                ("tref", C.c_double), ("densit", C.c_double), ("gravx", C.c_double), ("gravy", C.c_double),
                ("gravz", C.c_double), ("sol", FcoSolverOpts)]


class FcoPisoReport(C.Structure):
    _fields_ = [("rep", FcoReport * 16), ("nsolves", C.c_int), ("sumLocalContErr", C.c_double),
                ("globalContErr", C.c_double)]


def piso_opts(ncorr=1, npcor=1, nigrad=1, pRefCell=1, pimple=False, urf_p=0.3, const_mflux=False, flomas=0.0,
              bdf=True, btime=0.0, timestep=1e-3, cn=False, lbuoy=False, boussinesq=True, beta=0.0, tref=0.0,
              densit=1.0, grav=(0.0, 0.0, 0.0), sor=1e-2, nsw=100, small=SMALL, tol=TOL) -> FcoPisoOpts:
    return FcoPisoOpts(ncorr, npcor, nigrad, 2, pRefCell, int(pimple), urf_p, int(const_mflux), flomas, int(bdf), btime,
                       timestep, int(cn), int(lbuoy), int(boussinesq), beta, tref, densit, grav[0], grav[1], grav[2],
                       FcoSolverOpts(sor, nsw, small, tol, 0))


def get_rAU_x_UEqnH(mesh, csr: Csr, f: Fields, x: UvwFields, opts: FcoPisoOpts, h: np.ndarray) -> None:
    ms, cs, fs, xs = mesh_struct(mesh), csr.c(), f.c(), x.c()
    lib().fco_get_rAU_x_UEqnH(C.byref(ms), C.byref(cs), C.byref(fs), C.byref(xs), C.byref(opts), _d(h))


def piso(mesh, csr: Csr, f: Fields, x: UvwFields, opts: FcoPisoOpts) -> FcoPisoReport:
    """PISO_multiple_correction / PIMPLE_multiple_correction; f.a must hold the momentum matrix of calcuvw."""
    ms, cs, fs, xs = mesh_struct(mesh), csr.c(), f.c(), x.c()
    rep = FcoPisoReport()
    h = np.zeros(csr.nnz)
    rc = lib().fco_piso(C.byref(ms), C.byref(cs), C.byref(fs), C.byref(xs), C.byref(opts), _d(h), C.byref(rep))
    assert rc == 0, rc
    return rep


# ---- least-squares gradients + slope limiters (fc_oracle_grad.c) ----
GRAD_METHODS = {"gauss": 0, "lstsq": 1, "lstsq_qr": 2, "lstsq_dm": 3}
LIMITERS = {"no-limit": 0, "Barth-Jespersen": 1, "Venkatakrishnan": 2, "mVenkatakrishnan": 3}


class FcoGradientCfg(C.Structure):
    _fields_ = [("method", C.c_int), ("limiter", C.c_int), ("dmat", dp), ("dmatqr", dp), ("small", C.c_double)]


def lsq_matrix(mesh, weighted: bool) -> np.ndarray:
    ms = mesh_struct(mesh)
    dmat = np.zeros((mesh.numCells, 9))
    lib().fco_lsq_matrix(C.byref(ms), int(weighted), _d(dmat))
    return dmat


def lsq_qr_matrix(mesh):
    """Returns (D (numCells,6,3) = R^-1 Q^T stored like Fortran's D(3,6,numCells), cells the routine is undefined for)."""
    ms = mesh_struct(mesh)
    D = np.zeros((mesh.numCells, 6, 3))
    bad = lib().fco_lsq_qr_matrix(C.byref(ms), _d(D))
    return D, bad


def grad_lsq(mesh, weighted: bool, dmat: np.ndarray, fi: np.ndarray) -> np.ndarray:
    ms = mesh_struct(mesh)
    out = np.zeros((mesh.numCells + mesh.npro, 3))
    lib().fco_grad_lsq(C.byref(ms), int(weighted), _d(dmat), _d(fi), _d(out))
    return out


def grad_lsq_qr(mesh, D: np.ndarray, fi: np.ndarray) -> np.ndarray:
    ms = mesh_struct(mesh)
    out = np.zeros((mesh.numCells + mesh.npro, 3))
    lib().fco_grad_lsq_qr(C.byref(ms), _d(D), _d(fi), _d(out))
    return out


def slope_limiter(mesh, csr: Csr, which: str, phi: np.ndarray, grad: np.ndarray, small: float = SMALL) -> None:
    ms, cs = mesh_struct(mesh), csr.c()
    lib().fco_slope_limiter(C.byref(ms), C.byref(cs), LIMITERS[which], _d(phi), _d(grad), C.c_double(small))


_GRAD_KEEP = []


def set_gradient(method: str = "gauss", limiter: str = "no-limit", mesh=None, small: float = SMALL):
    """Configure the oracle's `grad` dispatcher (process-wide, like the reference's lstsq / lstsq_qr / lstsq_dm /
    gauss flags and `limiter` string); builds the geometric matrices the method needs (create_lsq_gradients_matrix)."""
    _GRAD_KEEP.clear()
    dmat = dmatqr = None
    if method in ("lstsq", "lstsq_dm"):
        dmat = lsq_matrix(mesh, method == "lstsq_dm")
    elif method == "lstsq_qr":
        dmatqr, bad = lsq_qr_matrix(mesh)
        assert bad == 0, f"lstsq_qr is defined for cells with exactly 6 neighbours ({bad} cells have another count)"
    _GRAD_KEEP.extend([dmat, dmatqr])
    cfg = FcoGradientCfg(GRAD_METHODS[method], LIMITERS[limiter], _d(dmat), _d(dmatqr), small)
    lib().fco_set_gradient(C.byref(cfg))


def grad(mesh, csr: Csr, phi: np.ndarray, nigrad: int = 1) -> np.ndarray:
    ms, cs = mesh_struct(mesh), csr.c()
    out = np.zeros((mesh.numCells + mesh.npro, 3))
    lib().fco_grad(C.byref(ms), C.byref(cs), _d(phi), nigrad, _d(out))
    return out
