"""ctypes binding of ``libfcapp_cuda.so`` (C ABI: ``include/fcapp.h``).

This is the Python stand-in for the Fortran ISO_C_BINDING shim (``fortran/fcapp_shim.f90``):
it passes exactly what gfortran would pass -- 1-based int32 index arrays, float64 arrays,
plain pointers -- and nothing else.  There is no CPU path: if the shared library has not
been built (``python -c "import __graft_entry__ as g; g.build()"``) loading fails loudly,
and without a CUDA device ``fc_create`` returns ``FC_ERR_NODEVICE``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfcapp_cuda.so")

FC_OK, FC_ERR_ARG, FC_ERR_CUDA, FC_ERR_NCCL, FC_ERR_UNSUPPORTED, FC_ERR_NODEVICE = range(6)
DPCG, ICCG, BICGSTAB = 0, 1, 2
SOLVERS = {"dpcg": DPCG, "iccg": ICCG, "bicgstab": BICGSTAB}

FIELDS = ("U", "V", "W", "P", "PP", "DEN", "FLMASS", "APU", "APV", "APW", "DUDXI", "DVDXI", "DWDXI", "DPDXI",
          "A", "SU", "RES", "FMI", "FMO", "APR", "FMPRO", "SCRATCH_T", "USER0", "USER1", "USER2", "USER3",
          # momentum predictor (fc_calcuvw), allocated on first use
          "VIS", "UO", "VO", "WO", "UOO", "VOO", "WOO", "T", "SV", "SW", "SPU", "SPV", "SP")
F = {name: i for i, name in enumerate(FIELDS)}

SMALL = float(np.float32(1e-20))   # `small` of module parameters is a default-real literal (modules_allocatable.f90:27)
TOL = float(np.float32(1e-13))     # dpcg.f90:37

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)

# every symbol include/fcapp.h declares (checked by tests/test_abi.py)
SYMBOLS = (
    "fc_create", "fc_destroy", "fc_last_error", "fc_version", "fc_comm_unique_id", "fc_comm_init", "fc_set_mesh",
    "fc_create_csr", "fc_field_size", "fc_upload", "fc_download", "fc_fill", "fc_synchronize", "fc_spmv",
    "fc_grad_gauss", "fc_grad_gauss_corrected", "fc_bpres", "fc_laplacian", "fc_solve", "fc_solve_host",
    "fc_solve_csr", "fc_calcp_assemble", "fc_calcp", "fc_calcp_host", "fc_exchange", "fc_global_sum",
    "fc_get_timings", "fc_time_spmv", "fc_stream", "fc_copy", "fc_set_spmv_sampling", "fc_comm_p2p_blob",
    "fc_comm_p2p_open", "fc_set_tuning", "fc_calcuvw_assemble", "fc_calcuvw_component", "fc_calcuvw",
    "fc_calcuvw_host", "fc_piso", "fc_set_gradient", "fc_grad", "fc_dpcg", "fc_iccg", "fc_bicgstab",
    "fc_calcp_correct", "fc_sweep_schedule_info",
)
GRAD_METHODS = {"gauss": 0, "lstsq": 1, "lstsq_qr": 2, "lstsq_dm": 3}
LIMITERS = {"no-limit": 0, "Barth-Jespersen": 1, "Venkatakrishnan": 2, "mVenkatakrishnan": 3}
TUNE_SPMV_KERNEL, TUNE_DPCG_PERSISTENT, TUNE_CTAS_PER_SM, TUNE_PIPE_GEOMETRY, TUNE_SWEEP_P2P = 0, 1, 2, 3, 4
TUNE_SWEEP_TILED = 5
TUNE_FUSED_GRAD = 6
TUNE_TILE_CTAS = 7
TUNE_SWEEP_CHECK = 8
TUNE_L2_KEEP = 9
TUNE_DPCG_FUSED = 10
TUNE_FACE_OCC = 11
TUNE_MAT_KEEP = 12
TUNE_JA_CODED = 13
TUNE_X_PREFETCH = 14
TUNE_DPCG_EAGER = 15
TUNE_KEYS = {"spmv_kernel": TUNE_SPMV_KERNEL, "dpcg_persistent": TUNE_DPCG_PERSISTENT, "ctas_per_sm": TUNE_CTAS_PER_SM,
             "pipe_geometry": TUNE_PIPE_GEOMETRY, "sweep_p2p": TUNE_SWEEP_P2P, "sweep_tiled": TUNE_SWEEP_TILED,
             "fused_grad": TUNE_FUSED_GRAD, "tile_ctas": TUNE_TILE_CTAS, "sweep_check": TUNE_SWEEP_CHECK, "l2_keep": TUNE_L2_KEEP, "dpcg_fused": TUNE_DPCG_FUSED, "face_occ": TUNE_FACE_OCC, "mat_keep": TUNE_MAT_KEEP, "ja_coded": TUNE_JA_CODED, "x_prefetch": TUNE_X_PREFETCH, "dpcg_eager": TUNE_DPCG_EAGER}


class MeshDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "numCells", "numInnerFaces", "numFaces", "numTotal", "npro",
        "ninl", "nout", "nsym", "nwal", "npru", "noc",
        "iProcFacesStart", "iInletFacesStart", "iOutletFacesStart", "iSymmetryFacesStart",
        "iWallFacesStart", "iPressOutletFacesStart", "iOCFacesStart")] + [
        ("owner", ip), ("neighbour", ip),
        ("xc", dp), ("yc", dp), ("zc", dp), ("vol", dp),
        ("arx", dp), ("ary", dp), ("arz", dp), ("xf", dp), ("yf", dp), ("zf", dp),
        ("facint", dp), ("fpro", dp),
        ("numConnections", C.c_int), ("neighbProcNo", ip), ("neighbProcOffset", ip), ("gloCells", C.c_int)]


class SolverOpts(C.Structure):
    _fields_ = [("sor", C.c_double), ("nsw", C.c_int), ("small", C.c_double), ("tol", C.c_double),
                ("parallel", C.c_int)]


class SolverReport(C.Structure):
    _fields_ = [("res0", C.c_double), ("resl", C.c_double), ("iters", C.c_int)]


class CalcpOpts(C.Structure):
    _fields_ = [("npcor", C.c_int), ("nigrad", C.c_int), ("nipgrad", C.c_int), ("pRefCell", C.c_int),
                ("urf_p", C.c_double), ("solver", C.c_int), ("const_mflux", C.c_int), ("flomas", C.c_double),
                ("lsq_flag", C.c_int), ("flux_variant", C.c_int), ("sol", SolverOpts)]


class CalcpReport(C.Structure):
    _fields_ = [("rep", SolverReport * 8), ("sumLocalContErr", C.c_double), ("globalContErr", C.c_double)]


class Timings(C.Structure):
    _fields_ = [("solve_ms", C.c_double), ("assemble_ms", C.c_double), ("correct_ms", C.c_double),
                ("spmv_ms", C.c_double), ("spmv_samples", C.c_int), ("sweep_tiles", C.c_int), ("launches", C.c_longlong),
                ("persist_ms", C.c_double), ("persist_pupdate_ms", C.c_double), ("persist_spmv_ms", C.c_double),
                ("persist_update_ms", C.c_double), ("persist_iters", C.c_int), ("persist_grid", C.c_int),
                ("persist_mail_ms", C.c_double), ("uvw_assemble_ms", C.c_double), ("uvw_solve_ms", C.c_double),
                ("persist_index_bytes", C.c_int), ("column_offsets", C.c_int)]


# convective schemes of read_input.f90:97-133 -> (face_value branch, limiter) of fc_calcuvw_opts
SCHEMES = {"central": (0, 7), "cds-corrected": (1, 7), "central-f": (2, 7), "linear-f": (3, 7), "muscl-f": (4, 7),
           "smart": (5, 0), "avl-smart": (5, 1), "muscl": (5, 2), "umist": (5, 3), "koren": (5, 4), "charm": (5, 5),
           "ospre": (5, 6), "linear": (5, 7)}


class CalcuvwOpts(C.Structure):
    _fields_ = [("nigrad", C.c_int), ("nipgrad", C.c_int), ("scheme", C.c_int), ("limiter", C.c_int),
                ("gds", C.c_double), ("urf", C.c_double * 3), ("sor", C.c_double * 3), ("nsw", C.c_int * 3),
                ("bdf", C.c_int), ("btime", C.c_double), ("timestep", C.c_double), ("cn", C.c_int),
                ("const_mflux", C.c_int), ("gradPcmf", C.c_double),
                ("lbuoy", C.c_int), ("boussinesq", C.c_int), ("beta", C.c_double), ("tref", C.c_double),
                ("densit", C.c_double), ("gravx", C.c_double), ("gravy", C.c_double), ("gravz", C.c_double),
                ("viscos", C.c_double), ("sol", SolverOpts)]


class CalcuvwReport(C.Structure):
    _fields_ = [("rep", SolverReport * 3)]


def calcuvw_opts(scheme="muscl-f", gds=1.0, urf=(0.7, 0.7, 0.7), sor=(1e-2, 1e-2, 1e-2), nsw=(20, 20, 20), nigrad=1,
                 bdf=False, btime=0.0, timestep=1e20, cn=False, const_mflux=False, gradPcmf=0.0, lbuoy=False,
                 boussinesq=True, beta=0.0, tref=0.0, densit=1.0, grav=(0.0, 0.0, 0.0), viscos=0.01, small=SMALL,
                 tol=TOL) -> CalcuvwOpts:
    """Options of ``call calcuvw`` with the names of the reference's ``input`` file / module parameters."""
    sc, lim = SCHEMES[scheme]
    return CalcuvwOpts(nigrad, 2, sc, lim, gds, (C.c_double * 3)(*urf), (C.c_double * 3)(*sor), (C.c_int * 3)(*nsw),
                       int(bdf), btime, timestep, int(cn), int(const_mflux), gradPcmf, int(lbuoy), int(boussinesq),
                       beta, tref, densit, grav[0], grav[1], grav[2], viscos, SolverOpts(0.0, 0, small, tol, 0))


class PisoOpts(C.Structure):
    _fields_ = [("ncorr", C.c_int), ("npcor", C.c_int), ("nigrad", C.c_int), ("nipgrad", C.c_int),
                ("pRefCell", C.c_int), ("pimple", C.c_int), ("urf_p", C.c_double), ("const_mflux", C.c_int),
                ("flomas", C.c_double), ("bdf", C.c_int), ("btime", C.c_double), ("timestep", C.c_double),
                ("cn", C.c_int), ("lbuoy", C.c_int), ("boussinesq", C.c_int), ("beta", C.c_double),
                ("tref", C.c_double), ("densit", C.c_double), ("gravx", C.c_double), ("gravy", C.c_double),
                ("gravz", C.c_double), ("sol", SolverOpts)]


class PisoReport(C.Structure):
    _fields_ = [("rep", SolverReport * 16), ("nsolves", C.c_int), ("sumLocalContErr", C.c_double),
                ("globalContErr", C.c_double)]


def piso_opts(ncorr=1, npcor=1, nigrad=1, pRefCell=1, pimple=False, urf_p=0.3, const_mflux=False, flomas=0.0,
              bdf=True, btime=0.0, timestep=1e-3, cn=False, lbuoy=False, boussinesq=True, beta=0.0, tref=0.0,
              densit=1.0, grav=(0.0, 0.0, 0.0), sor=1e-2, nsw=100, small=SMALL, tol=TOL) -> PisoOpts:
    """Options of ``PISO_multiple_correction`` / ``PIMPLE_multiple_correction`` (pimple=True)."""
    return PisoOpts(ncorr, npcor, nigrad, 2, pRefCell, int(pimple), urf_p, int(const_mflux), flomas, int(bdf), btime,
                    timestep, int(cn), int(lbuoy), int(boussinesq), beta, tref, densit, grav[0], grav[1], grav[2],
                    SolverOpts(sor, nsw, small, tol, 0))


class FcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libfcapp_cuda error {code}: {msg}")
        self.code = code


_LIB: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the CUDA library.  Raises if it has not been built -- never falls back."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `make -C freecappuccino_b200/csrc` "
                "(or __graft_entry__.build()).  freecappuccino_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.fc_last_error.restype = C.c_char_p
        lib.fc_sweep_schedule_info.restype = C.c_char_p
        lib.fc_sweep_schedule_info.argtypes = [C.c_void_p]
        lib.fc_last_error.argtypes = [C.c_void_p]
        lib.fc_stream.restype = C.c_void_p
        lib.fc_stream.argtypes = [C.c_void_p]
        for name in SYMBOLS:
            fn = getattr(lib, name)
            if name not in ("fc_last_error", "fc_stream", "fc_sweep_schedule_info"):
                fn.restype = C.c_int
        _LIB = lib
    return _LIB


def _d(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous, "float64 C-contiguous array expected"
    return a.ctypes.data_as(dp)


def _i(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous, "int32 C-contiguous array expected"
    return a.ctypes.data_as(ip)


def solver_opts(sor: float, nsw: int, small: float = SMALL, tol: float = TOL, parallel: bool = False) -> SolverOpts:
    return SolverOpts(sor, nsw, small, tol, int(parallel))


def calcp_opts(npcor=1, nigrad=1, pRefCell=1, urf_p=0.3, solver="iccg", const_mflux=False, flomas=0.0,
               lsq_flag=False, flux_variant=0, sor=1e-2, nsw=100, small=SMALL, tol=TOL, parallel=False) -> CalcpOpts:
    return CalcpOpts(npcor, nigrad, 2, pRefCell, urf_p, SOLVERS[solver], int(const_mflux), flomas,
                     int(lsq_flag), flux_variant, SolverOpts(sor, nsw, small, tol, int(parallel)))


class Context:
    """One ``fc_context`` = one rank / one GPU."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.fc_create(device, C.byref(self.h))
        if rc != FC_OK:
            raise FcError(rc, (self.lib.fc_last_error(None) or b"").decode())
        self.mesh = None
        self._keep = []
        # A/B switch for measurements without touching the caller: FCAPP_TUNE="sweep_tiled=2,fused_grad=1"
        for item in filter(None, os.environ.get("FCAPP_TUNE", "").replace(" ", "").split(",")):
            name, _, value = item.partition("=")
            self.set_tuning(TUNE_KEYS[name.lower()], int(value))

    # -- plumbing ---------------------------------------------------------
    def _ck(self, rc: int):
        if rc != FC_OK:
            raise FcError(rc, (self.lib.fc_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            self.lib.fc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, rank: int, nranks: int, uid: bytes):
        self._ck(self.lib.fc_comm_init(self.h, rank, nranks, C.c_char_p(uid)))

    P2P_BLOB_BYTES = 512

    def p2p_blob(self) -> bytes:
        buf = C.create_string_buffer(self.P2P_BLOB_BYTES)
        self._ck(self.lib.fc_comm_p2p_blob(self.h, buf))
        return buf.raw

    def p2p_open(self, blobs):
        raw = b"".join(blobs)
        self._ck(self.lib.fc_comm_p2p_open(self.h, C.c_char_p(raw), len(blobs)))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = load().fc_comm_unique_id(buf)
        if rc != FC_OK:
            raise FcError(rc, "ncclGetUniqueId failed")
        return buf.raw

    # -- mesh / pattern ---------------------------------------------------
    def set_mesh(self, m):
        """``m``: :class:`freecappuccino_b200.mesh.Mesh` (arrays exactly as the Fortran host holds them)."""
        s = MeshDesc()
        s.numCells, s.numInnerFaces, s.numFaces, s.numTotal, s.npro = (m.numCells, m.numInnerFaces, m.numFaces,
                                                                       m.numTotal, m.npro)
        s.ninl, s.nout, s.nsym, s.nwal, s.npru, s.noc = (m.count("inlet"), m.count("outlet"), m.count("symmetry"),
                                                         m.count("wall"), m.count("prOutlet"), m.noc)
        s.iProcFacesStart = m.iProcFacesStart
        s.iInletFacesStart, s.iOutletFacesStart = m.faces_start("inlet"), m.faces_start("outlet")
        s.iSymmetryFacesStart, s.iWallFacesStart = m.faces_start("symmetry"), m.faces_start("wall")
        s.iPressOutletFacesStart, s.iOCFacesStart = m.faces_start("prOutlet"), m.iOCFacesStart
        keep = []

        def d(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            return _d(a)

        def i(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.int32)
            keep.append(a)
            return _i(a)

        s.owner, s.neighbour = i(m.owner), i(m.neighbour)
        s.xc, s.yc, s.zc, s.vol = d(m.xc), d(m.yc), d(m.zc), d(m.vol)
        s.arx, s.ary, s.arz, s.xf, s.yf, s.zf = d(m.arx), d(m.ary), d(m.arz), d(m.xf), d(m.yf), d(m.zf)
        s.facint, s.fpro = d(m.facint), d(m.fpro)
        s.numConnections = 0 if m.neighbProcNo is None else int(len(m.neighbProcNo))
        s.neighbProcNo, s.neighbProcOffset = i(m.neighbProcNo), i(m.neighbProcOffset)
        s.gloCells = m.gloCells or m.numCells
        self._ck(self.lib.fc_set_mesh(self.h, C.byref(s)))
        self.mesh = m

    def create_csr(self, download: bool = True):
        """``create_CSR_matrix_from_mesh_data``: returns the 1-based (ioffset, ja, diag, icell_jcell, jcell_icell)."""
        m = self.mesh
        if not download:
            self._ck(self.lib.fc_create_csr(self.h, None, None, None, None, None))
            return None
        ioffset = np.zeros(m.numCells + 1, np.int32)
        ja = np.zeros(m.nnz, np.int32)
        diag = np.zeros(m.numCells, np.int32)
        ij = np.zeros(m.numInnerFaces, np.int32)
        ji = np.zeros(m.numInnerFaces, np.int32)
        self._ck(self.lib.fc_create_csr(self.h, _i(ioffset), _i(ja), _i(diag), _i(ij), _i(ji)))
        return ioffset, ja, diag, ij, ji

    # -- fields -----------------------------------------------------------
    def field_size(self, name: str) -> int:
        n = C.c_size_t()
        self._ck(self.lib.fc_field_size(self.h, F[name], C.byref(n)))
        return n.value

    def upload(self, name: str, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        self._ck(self.lib.fc_upload(self.h, F[name], _d(a), C.c_size_t(a.size)))

    def download(self, name: str, n: Optional[int] = None) -> np.ndarray:
        n = self.field_size(name) if n is None else n
        out = np.zeros(n)
        self._ck(self.lib.fc_download(self.h, F[name], _d(out), C.c_size_t(n)))
        return out

    def fill(self, name: str, value: float):
        self._ck(self.lib.fc_fill(self.h, F[name], C.c_double(value)))

    def copy(self, src: str, dst: str):
        self._ck(self.lib.fc_copy(self.h, F[src], F[dst]))

    def set_tuning(self, key: int, value: int):
        """Kernel selection for A/B measurements (TUNE_* keys, include/fcapp.h)."""
        self._ck(self.lib.fc_set_tuning(self.h, int(key), int(value)))

    def sweep_schedule_info(self) -> str:
        """Which schedule the triangular sweeps of the last iccg / bicgstab solve used (fc_sweep_schedule_info)."""
        return (self.lib.fc_sweep_schedule_info(self.h) or b"").decode()

    def set_spmv_sampling(self, max_samples: int):
        self._ck(self.lib.fc_set_spmv_sampling(self.h, max_samples))

    def synchronize(self):
        self._ck(self.lib.fc_synchronize(self.h))

    # -- operators --------------------------------------------------------
    def spmv(self, x: str, y: str):
        self._ck(self.lib.fc_spmv(self.h, F[x], F[y]))

    def time_spmv(self, x: str, y: str, reps: int) -> float:
        ms = C.c_double()
        self._ck(self.lib.fc_time_spmv(self.h, F[x], F[y], reps, C.byref(ms)))
        return ms.value

    def grad_gauss(self, phi: str, grad: str, nigrad: int = 1):
        self._ck(self.lib.fc_grad_gauss(self.h, F[phi], F[grad], nigrad))

    def grad_gauss_corrected(self, phi: str, grad: str, zero_seed: bool = False):
        self._ck(self.lib.fc_grad_gauss_corrected(self.h, F[phi], F[grad], int(zero_seed)))

    def bpres(self, p: str, istage: int):
        self._ck(self.lib.fc_bpres(self.h, F[p], istage))

    def laplacian(self, mu: str, phi: str):
        self._ck(self.lib.fc_laplacian(self.h, F[mu], F[phi]))

    def solve(self, solver: str, fi: str, opts: SolverOpts) -> SolverReport:
        rep = SolverReport()
        self._ck(self.lib.fc_solve(self.h, SOLVERS[solver], F[fi], C.byref(opts), C.byref(rep)))
        return rep

    def dpcg(self, fi: str, opts: SolverOpts) -> SolverReport:
        """``call dpcg(fi,ifi)`` on a device-resident field (fc_dpcg)."""
        rep = SolverReport()
        self._ck(self.lib.fc_dpcg(self.h, F[fi], C.byref(opts), C.byref(rep)))
        return rep

    def iccg(self, fi: str, opts: SolverOpts) -> SolverReport:
        rep = SolverReport()
        self._ck(self.lib.fc_iccg(self.h, F[fi], C.byref(opts), C.byref(rep)))
        return rep

    def bicgstab(self, fi: str, opts: SolverOpts) -> SolverReport:
        rep = SolverReport()
        self._ck(self.lib.fc_bicgstab(self.h, F[fi], C.byref(opts), C.byref(rep)))
        return rep

    def solve_host(self, solver: str, a: np.ndarray, su: np.ndarray, fi: np.ndarray, opts: SolverOpts,
                   res: Optional[np.ndarray] = None) -> SolverReport:
        rep = SolverReport()
        self._ck(self.lib.fc_solve_host(self.h, SOLVERS[solver], _d(a), _d(su), _d(fi), _d(res), C.byref(opts),
                                        C.byref(rep)))
        return rep

    def solve_csr(self, solver: str, ioffset, ja, diag, a, su, fi, opts: SolverOpts, history: bool = False):
        rep = SolverReport()
        hist = np.zeros(max(opts.nsw, 1)) if history else None
        self._ck(self.lib.fc_solve_csr(self.h, SOLVERS[solver], int(diag.size), int(ja.size), _i(ioffset), _i(ja),
                                       _i(diag), _d(a), _d(su), _d(fi), C.byref(opts), C.byref(rep), _d(hist)))
        return (rep, hist[:rep.iters]) if history else rep

    def calcp_assemble(self, opts: CalcpOpts):
        self._ck(self.lib.fc_calcp_assemble(self.h, C.byref(opts)))

    def calcp(self, opts: CalcpOpts) -> CalcpReport:
        rep = CalcpReport()
        self._ck(self.lib.fc_calcp(self.h, C.byref(opts), C.byref(rep)))
        return rep

    def calcp_correct(self, opts: CalcpOpts, ipcorr: int) -> CalcpReport:
        """Post-solve half of corrector ``ipcorr`` (FC_PP = the solved correction); the report's continuity errors are
        filled after the last corrector."""
        rep = CalcpReport()
        self._ck(self.lib.fc_calcp_correct(self.h, C.byref(opts), int(ipcorr), C.byref(rep)))
        return rep

    def calcp_host(self, opts: CalcpOpts, u, v, w, p, pp, apu, apv, apw, flmass) -> CalcpReport:
        rep = CalcpReport()
        self._ck(self.lib.fc_calcp_host(self.h, C.byref(opts), _d(u), _d(v), _d(w), _d(p), _d(pp), _d(apu), _d(apv),
                                        _d(apw), _d(flmass), C.byref(rep)))
        return rep

    # -- momentum predictor ----------------------------
    def calcuvw_assemble(self, opts: CalcuvwOpts):
        self._ck(self.lib.fc_calcuvw_assemble(self.h, C.byref(opts)))

    def calcuvw_component(self, opts: CalcuvwOpts, comp: int) -> SolverReport:
        rep = SolverReport()
        self._ck(self.lib.fc_calcuvw_component(self.h, C.byref(opts), comp, C.byref(rep)))
        return rep

    def calcuvw(self, opts: CalcuvwOpts) -> CalcuvwReport:
        rep = CalcuvwReport()
        self._ck(self.lib.fc_calcuvw(self.h, C.byref(opts), C.byref(rep)))
        return rep

    def calcuvw_host(self, opts: CalcuvwOpts, u, v, w, p, vis, flmass, apu, apv, apw) -> CalcuvwReport:
        rep = CalcuvwReport()
        self._ck(self.lib.fc_calcuvw_host(self.h, C.byref(opts), _d(u), _d(v), _d(w), _d(p), _d(vis), _d(flmass),
                                          _d(apu), _d(apv), _d(apw), C.byref(rep)))
        return rep

    def set_gradient(self, method: str = "gauss", limiter: str = "no-limit", small: float = SMALL):
        """The `grad` dispatcher's scheme (lstsq / lstsq_qr / lstsq_dm / gauss + limiter of the input file)."""
        self._ck(self.lib.fc_set_gradient(self.h, GRAD_METHODS[method], LIMITERS[limiter], C.c_double(small)))

    def grad(self, phi: str, grad: str, nigrad: int = 1):
        self._ck(self.lib.fc_grad(self.h, F[phi], F[grad], nigrad))

    def piso(self, opts: PisoOpts) -> PisoReport:
        rep = PisoReport()
        self._ck(self.lib.fc_piso(self.h, C.byref(opts), C.byref(rep)))
        return rep

    def exchange(self, field: str):
        self._ck(self.lib.fc_exchange(self.h, F[field]))

    def global_sum(self, value: float) -> float:
        v = C.c_double(value)
        self._ck(self.lib.fc_global_sum(self.h, C.byref(v)))
        return v.value

    def timings(self) -> Timings:
        t = Timings()
        self._ck(self.lib.fc_get_timings(self.h, C.byref(t)))
        return t
