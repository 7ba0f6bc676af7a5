"""Index logic and arithmetic order of the momentum-predictor kernel bodies, checked WITHOUT a GPU.

freecappuccino_b200/csrc/fc_momentum_body.cuh holds the per-index bodies the CUDA kernels of
fc_momentum.cu call.  tests/kernel_bodies_host/fcm_host.cpp compiles that header with g++
(-ffp-contract=off, the host twin of nvcc -fmad=false) and runs the bodies in plain loops over the
device data layout (0-based indices, cell-to-face map) built here in numpy.  Everything must equal
the oracle's calcuvw bit for bit.  This is test infrastructure: the product has no host path
(tests/test_abi.py asserts the package never references it); the GPU twin of this test is
tests/test_gpu_zz1_momentum.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from freecappuccino_b200 import cases
from freecappuccino_b200 import mesh as M
from oracle import oracle

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernel_bodies_host")
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


class Geom(C.Structure):
    _fields_ = [("owner", ip), ("neigh", ip)] + [(k, dp) for k in (
        "xc", "yc", "zc", "vol", "arx", "ary", "arz", "xf", "yf", "zf", "facint")] + [("n", C.c_int), ("F", C.c_int)]


class C2f(C.Structure):
    _fields_ = [(k, ip) for k in ("off", "face", "other", "pos")]


class Slots(C.Structure):
    _fields_ = [("slot", C.c_int * 5), ("face", C.c_int * 5), ("count", C.c_int * 5)]


class Flow(C.Structure):
    _fields_ = [(k, dp) for k in ("u", "v", "w", "p", "den", "vis", "flmass", "fmi", "fmo", "dU", "dV", "dW", "dP",
                                  "uo", "vo", "wo", "uoo", "voo", "woo", "t")]


class Opts(C.Structure):
    _fields_ = [("scheme", C.c_int), ("limiter", C.c_int), ("gds", C.c_double), ("bdf", C.c_int),
                ("btime", C.c_double), ("timestep", C.c_double), ("cn", C.c_int), ("const_mflux", C.c_int),
                ("gradPcmf", C.c_double), ("lbuoy", C.c_int), ("boussinesq", C.c_int)] + [
        (k, C.c_double) for k in ("beta", "tref", "densit", "gravx", "gravy", "gravz", "viscos")]


class Faces(C.Structure):
    _fields_ = [(k, dp) for k in ("can", "cap", "sup", "svp", "swp", "fie")]


class Rows(C.Structure):
    _fields_ = [(k, dp) for k in ("a", "su", "sv", "sw", "spu", "spv", "sp")]


class Comp(C.Structure):
    _fields_ = [("ioffset", ip), ("diag", ip)] + [(k, dp) for k in ("a", "s", "spc", "su", "ap", "phi", "phio", "den")] + [
        (k, C.c_double) for k in ("urfrs", "urfms", "small", "timestep")] + [("cn", C.c_int), ("zero_diag", C.c_int),
                                                                              ("parallel", C.c_int), ("npro", C.c_int),
                                                                              ("apr", dp)]


class Proc(C.Structure):
    _fields_ = [("npro", C.c_int), ("pface0", C.c_int)] + [(k, dp) for k in ("fpro", "fmpro", "apr", "sup", "svp", "swp",
                                                                             "fie")]


class Hbya(C.Structure):
    _fields_ = [("ioffset", ip), ("diag", ip)] + [(k, dp) for k in (
        "h", "u", "v", "w", "uo", "vo", "wo", "uoo", "voo", "woo", "t", "den", "su", "sv", "sw")] + [
        ("bdf", C.c_int), ("btime", C.c_double), ("timestep", C.c_double), ("cn", C.c_int), ("lbuoy", C.c_int),
        ("boussinesq", C.c_int)] + [(k, C.c_double) for k in ("beta", "tref", "densit", "gravx", "gravy", "gravz")] + [
        ("apr", dp), ("npro", C.c_int)]


@pytest.fixture(scope="module")
def host():
    so = os.path.join(HERE, "libfcm_host.so")
    src = os.path.join(HERE, "fcm_host.cpp")
    hdrs = [os.path.join(HERE, "..", "..", "freecappuccino_b200", "csrc", h) for h in ("fc_momentum_body.cuh",
                                                                                         "fc_piso_body.cuh",
                                                                                         "fc_grad_body.cuh")]
    if not os.path.exists(so) or max(os.path.getmtime(f) for f in [src] + hdrs) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so, src])
    lib = C.CDLL(so)
    for k, st in enumerate((Geom, C2f, Slots, Flow, Opts, Faces, Rows, Comp, Hbya, Proc)):
        assert lib.fcm_host_sizes(k) == C.sizeof(st), (k, st)
    return lib


def d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(dp)


def i32(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(ip)


def device_layout(mesh, csr):
    """What fc_set_mesh / fc_create_csr keep on the device: 0-based owner / neighbour / CSR and
    the cell-to-face map in the reference's loop order (fc_csr.cu k_c2f_fill / k_c2f_sort)."""
    n, F, NF = mesh.numCells, mesh.numInnerFaces, mesh.numFaces
    owner = (mesh.owner - 1).astype(np.int32)
    neigh = (mesh.neighbour - 1).astype(np.int32)
    ioffset = (csr.ioffset - 1).astype(np.int32)
    ja = (csr.ja - 1).astype(np.int32)
    diag = (csr.diag - 1).astype(np.int32)
    ent = [[] for _ in range(n)]
    for f in range(F):
        ent[owner[f]].append((f, f, int(neigh[f])))
        ent[neigh[f]].append((f, f | 0x80000000, int(owner[f])))
    for i in range(mesh.npro):                 # processor faces: halo slot n + i, between inner and boundary entries
        f = mesh.iProcFacesStart + i
        ent[owner[f]].append((F + i, f, n + i))
    slots = Slots()
    for b, kind in enumerate(M.KINDS):
        fs, sl = mesh.boundary_faces(kind), mesh.boundary_slots(kind)
        slots.count[b] = len(fs)
        slots.face[b] = int(fs[0]) if len(fs) else 0
        slots.slot[b] = int(sl[0]) if len(sl) else 0
        for f, s in zip(fs, sl):
            ent[owner[f]].append((F + int(s) - n, int(f), int(s)))
    off = np.zeros(n + 1, np.int32)
    face, other, pos = [], [], []
    for c in range(n):
        ent[c].sort(key=lambda e: e[0])
        off[c + 1] = off[c] + len(ent[c])
        for _, fe, o in ent[c]:
            face.append(fe)
            other.append(o)
            if (fe & 0x7fffffff) < F:
                row = ja[ioffset[c]:ioffset[c + 1]]
                pos.append(int(ioffset[c] + np.nonzero(row == o)[0][0]))
            else:
                pos.append(-1)
    face = np.array(face, dtype=np.uint32).view(np.int32).copy()
    return dict(owner=owner, neigh=neigh, ioffset=ioffset, ja=ja, diag=diag, off=off, face=face,
                other=np.array(other, np.int32), pos=np.array(pos, np.int32), slots=slots)


def fluxes(mesh, of):
    F = mesh.numInnerFaces
    o, nb = mesh.owner[:F] - 1, mesh.neighbour - 1
    fx = mesh.facint
    df = of.den[o] * (1 - fx) + of.den[nb] * fx
    fl = df * ((of.u[o] * (1 - fx) + of.u[nb] * fx) * mesh.arx[:F] + (of.v[o] * (1 - fx) + of.v[nb] * fx) * mesh.ary[:F]
               + (of.w[o] * (1 - fx) + of.w[nb] * fx) * mesh.arz[:F])
    fl[::3] *= -1.0   # both flow directions on every kind of face
    return fl


def run_case(host, mesh, f, scheme, **kw):
    rng = np.random.default_rng(3)
    csr = oracle.create_csr(mesh)
    n, F, nt = mesh.numCells, mesh.numInnerFaces, mesh.numTotal
    of = oracle.Fields(mesh, csr.nnz)
    for k in ("u", "v", "w", "p", "den"):
        getattr(of, k)[:] = f[k]
    of.flmass[:] = fluxes(mesh, of)
    fmi, _ = cases.inlet_fluxes(mesh, f)
    of.fmi[:fmi.size] = fmi
    fs, sl = mesh.boundary_faces("outlet"), mesh.boundary_slots("outlet")
    of.fmo[:len(fs)] = f["den"][sl] * (f["u"][sl] * mesh.arx[fs] + f["v"][sl] * mesh.ary[fs] + f["w"][sl] * mesh.arz[fs])
    of.a[:] = rng.standard_normal(csr.nnz)   # stale matrix of the previous solve (the U diagonal sum reads it)
    x = oracle.UvwFields(mesh, of, 0.0)
    x.vis[:] = 0.01 * (1.0 + 0.3 * rng.random(nt))
    for k in ("uo", "vo", "wo", "uoo", "voo", "woo", "t"):
        getattr(x, k)[:] = rng.standard_normal(nt)
    opts = oracle.uvw_opts(scheme=scheme, urf=(0.7, 0.8, 0.6), sor=(1e-30,) * 3, nsw=(2, 2, 2), **kw)

    # ---- the kernel bodies on the device layout, fed with the oracle's gradients / boundary pressure ----
    L = device_layout(mesh, csr)
    g = dict(u=of.u.copy(), v=of.v.copy(), w=of.w.copy(), p=of.p.copy(), a=of.a.copy())
    dU = oracle.grad_gauss(mesh, g["u"], 1)
    dV = oracle.grad_gauss(mesh, g["v"], 1)
    dW = oracle.grad_gauss(mesh, g["w"], 1)
    dP = of.dPdxi.copy()
    for istage in (1, 2):
        oracle.bpres(mesh, g["p"], dP, istage)
        dP = oracle.grad_gauss(mesh, g["p"], 1)
    geo = {k: np.ascontiguousarray(getattr(mesh, k), dtype=np.float64) for k in
           ("xc", "yc", "zc", "vol", "arx", "ary", "arz", "xf", "yf", "zf", "facint")}
    G = Geom(i32(L["owner"]), i32(L["neigh"]), *[d(geo[k]) for k in ("xc", "yc", "zc", "vol", "arx", "ary", "arz",
                                                                    "xf", "yf", "zf", "facint")], n, F)
    Mp = C2f(i32(L["off"]), i32(L["face"]), i32(L["other"]), i32(L["pos"]))
    den = np.ascontiguousarray(of.den)
    FL = Flow(d(g["u"]), d(g["v"]), d(g["w"]), d(g["p"]), d(den), d(x.vis), d(of.flmass), d(of.fmi), d(of.fmo),
              d(dU), d(dV), d(dW), d(dP), d(x.uo), d(x.vo), d(x.wo), d(x.uoo), d(x.voo), d(x.woo), d(x.t))
    O = Opts(opts.scheme, opts.limiter, opts.gds, opts.bdf, opts.btime, opts.timestep, opts.cn, opts.const_mflux,
             opts.gradPcmf, opts.lbuoy, opts.boussinesq, opts.beta, opts.tref, opts.densit, opts.gravx, opts.gravy,
             opts.gravz, opts.viscos)
    fa = [np.zeros(max(F, 1)) for _ in range(6)]
    rows = {k: np.zeros(n) for k in ("su", "sv", "sw", "spu", "spv", "sp")}
    FA = Faces(*[d(a) for a in fa])
    R = Rows(d(g["a"]), *[d(rows[k]) for k in ("su", "sv", "sw", "spu", "spv", "sp")])
    one = np.zeros(1)
    P0 = Proc(0, 0, *[d(one)] * 7)
    host.fcm_host_assemble(C.byref(G), C.byref(Mp), C.byref(L["slots"]), C.byref(FL), C.byref(O), C.byref(FA),
                           C.byref(P0), C.byref(R), csr.nnz)

    # ---- oracle ----
    oracle.calcuvw_assemble(mesh, csr, of, x, opts)
    assert np.array_equal(g["p"], of.p)
    assert np.array_equal(dP, of.dPdxi) and np.array_equal(dU, of.dUdxi)
    for k, ref in (("su", of.su), ("sv", x.sv), ("sw", x.sw), ("spu", x.spu), ("spv", x.spv), ("sp", x.sp)):
        assert np.array_equal(rows[k], ref), (scheme, k, np.abs(rows[k] - ref).max())
    assert np.array_equal(g["a"], of.a)

    # ---- per component: diagonal / under-relaxation / ap*, then the oracle's solve moves phi ----
    ap = [np.zeros(n) for _ in range(3)]
    su_rhs = rows["su"]
    for comp, (sk, spk, phik, oldk) in enumerate((("su", "spu", "u", "uo"), ("sv", "spv", "v", "vo"),
                                                  ("sw", "sp", "w", "wo"))):
        K = Comp(i32(L["ioffset"]), i32(L["diag"]), d(g["a"]), d(rows[sk]), d(rows[spk]), d(su_rhs), d(ap[comp]),
                 d(g[phik]), d(getattr(x, oldk)), d(den), 1.0 / opts.urf[comp], 1.0 - opts.urf[comp], opts.sol.small,
                 opts.timestep, opts.cn, 1 if comp else 0, 0, 0, d(one))
        host.fcm_host_component(C.byref(G), C.byref(Mp), C.byref(K))
        a_before = g["a"].copy()
        # oracle: same step + BiCGStab; capture its matrix / rhs through a zero-sweep solve
        o0 = oracle.uvw_opts(scheme=scheme, urf=(0.7, 0.8, 0.6), sor=(1e-30,) * 3, nsw=(0, 0, 0), **kw)
        oracle.calcuvw_component(mesh, csr, of, x, o0, comp)
        assert np.array_equal(a_before, of.a), (scheme, comp)
        assert np.array_equal(su_rhs, of.su), (scheme, comp)
        assert np.array_equal(ap[comp], (x.apu, x.apv, x.apw)[comp][:n]), (scheme, comp)
        assert np.array_equal(rows[sk], (of.su, x.sv, x.sw)[comp]) or comp == 0
        assert np.array_equal(rows[spk], (x.spu, x.spv, x.sp)[comp])
    return True


MESHES = {
    "skew": lambda: cases.skew_case(6, 5, 4),
    "hex_mixed_bc": lambda: cases.hex_case(5, 4, 6, kinds=("inlet", "outlet", "wall", "symmetry", "wall", "prOutlet")),
    "poly": lambda: cases.poly_case(3),
}


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("scheme", ["muscl-f", "central", "cds-corrected", "central-f", "linear-f", "smart", "koren",
                                    "charm", "ospre", "linear"])
def test_bodies_equal_oracle_steady(host, name, scheme):
    mesh = MESHES[name]()
    assert run_case(host, mesh, cases.flow_fields(mesh), scheme)


@pytest.mark.parametrize("kw", [dict(bdf=True, btime=0.0, timestep=0.02), dict(bdf=True, btime=1.0, timestep=0.02),
                                dict(bdf=True, btime=1.0, timestep=0.02, cn=True),
                                dict(const_mflux=True, gradPcmf=0.37),
                                dict(lbuoy=True, boussinesq=True, beta=0.3, tref=0.1, densit=1.1, grav=(0.1, -9.81, 0.2)),
                                dict(lbuoy=True, boussinesq=False, densit=1.1, grav=(0.1, -9.81, 0.2))])
def test_bodies_equal_oracle_sources(host, kw):
    mesh = MESHES["skew"]()
    assert run_case(host, mesh, cases.flow_fields(mesh), "muscl-f", **kw)


# ---- PISO / PIMPLE bodies (fc_piso_body.cuh) ----
@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("kw", [dict(bdf=True, btime=0.0, timestep=0.01), dict(bdf=True, btime=1.0, timestep=0.01, cn=True),
                                dict(bdf=False, lbuoy=True, boussinesq=False, densit=1.1, grav=(0.1, -9.81, 0.2)),
                                dict(bdf=True, btime=1.0, timestep=0.01, lbuoy=True, beta=0.3, tref=0.1, densit=1.1,
                                     grav=(0.1, -9.81, 0.2))])
def test_hbya_body_equals_oracle(host, name, kw):
    mesh = MESHES[name]()
    rng = np.random.default_rng(11)
    csr = oracle.create_csr(mesh)
    n, F, nt = mesh.numCells, mesh.numInnerFaces, mesh.numTotal
    f = cases.flow_fields(mesh)
    of = oracle.Fields(mesh, csr.nnz)
    for k in ("u", "v", "w", "p", "den"):
        getattr(of, k)[:] = f[k]
    of.apu[:], of.apv[:], of.apw[:] = f["apu"], f["apv"], f["apw"]
    x = oracle.UvwFields(mesh, of, 0.01)
    for k in ("uo", "vo", "wo", "uoo", "voo", "woo", "t"):
        getattr(x, k)[:] = rng.standard_normal(nt)
    h = rng.standard_normal(csr.nnz)
    L = device_layout(mesh, csr)
    geo = {k: np.ascontiguousarray(getattr(mesh, k), dtype=np.float64) for k in
           ("xc", "yc", "zc", "vol", "arx", "ary", "arz", "xf", "yf", "zf", "facint")}
    G = Geom(i32(L["owner"]), i32(L["neigh"]), *[d(geo[k]) for k in ("xc", "yc", "zc", "vol", "arx", "ary", "arz",
                                                                    "xf", "yf", "zf", "facint")], n, F)
    Mp = C2f(i32(L["off"]), i32(L["face"]), i32(L["other"]), i32(L["pos"]))
    u, v, w = of.u.copy(), of.v.copy(), of.w.copy()
    su, sv, sw = np.zeros(n), np.zeros(n), np.zeros(n)
    po = oracle.piso_opts(**kw)
    den = np.ascontiguousarray(of.den)
    K = Hbya(i32(L["ioffset"]), i32(L["diag"]), d(h), d(u), d(v), d(w), d(x.uo), d(x.vo), d(x.wo), d(x.uoo), d(x.voo),
             d(x.woo), d(x.t), d(den), d(su), d(sv), d(sw), po.bdf, po.btime, po.timestep, po.cn, po.lbuoy,
             po.boussinesq, po.beta, po.tref, po.densit, po.gravx, po.gravy, po.gravz, None, 0)
    apu, apv, apw = (np.ascontiguousarray(a[:n]) for a in (of.apu, of.apv, of.apw))
    host.fcp_host_hbya(C.byref(G), C.byref(Mp), C.byref(K), d(apu), d(apv), d(apw), d(u), d(v), d(w))
    oracle.get_rAU_x_UEqnH(mesh, csr, of, x, po, h)
    for got, ref in ((su, of.su), (sv, x.sv), (sw, x.sw), (u, of.u), (v, of.v), (w, of.w)):
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("kw", [dict(bdf=True, btime=1.0, timestep=0.01), dict(bdf=True, cn=True, timestep=0.02),
                                dict(bdf=True, btime=1.0, timestep=0.01, lbuoy=True, beta=0.3, tref=0.1, densit=1.1,
                                     grav=(0.1, -9.81, 0.2))])
@pytest.mark.parametrize("name", ["skew", "poly"])
def test_hbya_body_on_a_partitioned_mesh_equals_the_lock_step_oracle(host, name, kw):
    """src-parallel/get_rAU_x_UEqnH.f90: processor-face terms with apr, all three components' terms landing in su."""
    from freecappuccino_b200 import mesh as M
    from oracle import oracle_par
    g = MESHES[name]()
    nr = 3
    parts = M.partition(g, M.rcb_ranks(g, nr) if name == "poly" else M.slab_ranks(g.numCells, nr), nr)
    rng = np.random.default_rng(12)
    pc = oracle_par.ParCase(parts)
    xs = pc.uvw_fields(0.01)
    po = oracle.piso_opts(**kw)
    hs, devs = [], []
    for r, mesh in enumerate(parts):
        nt, n = mesh.numTotal, mesh.numCells
        fr = pc.fields[r]
        for k in ("u", "v", "w", "den"):
            getattr(fr, k)[:] = rng.standard_normal(nt) if k != "den" else 1.0 + 0.1 * rng.random(nt)
        for k in ("uo", "vo", "wo", "uoo", "voo", "woo", "t"):
            getattr(xs[r], k)[:] = rng.standard_normal(nt)
        xs[r].apu[:], xs[r].apv[:], xs[r].apw[:] = (rng.random(n + mesh.npro) + 0.5 for _ in range(3))
        pc.apr[r][:] = rng.standard_normal(pc.apr[r].size)
        hs.append(rng.standard_normal(pc.csr[r].nnz))
    for r, mesh in enumerate(parts):
        n, F = mesh.numCells, mesh.numInnerFaces
        fr, x, csr = pc.fields[r], xs[r], pc.csr[r]
        L = device_layout(mesh, csr)
        geo = {k: np.ascontiguousarray(getattr(mesh, k), dtype=np.float64) for k in
               ("xc", "yc", "zc", "vol", "arx", "ary", "arz", "xf", "yf", "zf", "facint")}
        G = Geom(i32(L["owner"]), i32(L["neigh"]), *[d(geo[k]) for k in ("xc", "yc", "zc", "vol", "arx", "ary", "arz",
                                                                        "xf", "yf", "zf", "facint")], n, F)
        Mp = C2f(i32(L["off"]), i32(L["face"]), i32(L["other"]), i32(L["pos"]))
        u, v, w = fr.u.copy(), fr.v.copy(), fr.w.copy()
        su, sv, sw = np.zeros(n), np.zeros(n), np.zeros(n)
        den = np.ascontiguousarray(fr.den)
        apr = np.ascontiguousarray(pc.apr[r])
        K = Hbya(i32(L["ioffset"]), i32(L["diag"]), d(hs[r]), d(u), d(v), d(w), d(x.uo), d(x.vo), d(x.wo), d(x.uoo),
                 d(x.voo), d(x.woo), d(x.t), d(den), d(su), d(sv), d(sw), po.bdf, po.btime, po.timestep, po.cn, po.lbuoy,
                 po.boussinesq, po.beta, po.tref, po.densit, po.gravx, po.gravy, po.gravz, d(apr), mesh.npro)
        apu, apv, apw = (np.ascontiguousarray(a[:n]) for a in (x.apu, x.apv, x.apw))
        host.fcp_host_hbya(C.byref(G), C.byref(Mp), C.byref(K), d(apu), d(apv), d(apw), d(u), d(v), d(w))
        devs.append((su, sv, sw, u, v, w))
    pc.get_rAU_x_UEqnH(po, hs)
    for r in range(nr):
        fr, x = pc.fields[r], xs[r]
        n = parts[r].numCells
        for got, ref, nm in zip(devs[r], (fr.su, x.sv, x.sw, fr.u, fr.v, fr.w), ("su", "sv", "sw", "u", "v", "w")):
            assert np.array_equal(got[:n], ref[:n]), (r, nm)


def test_piso_tail_bodies(host):
    """pin row / flux correction from the matrix / velocity correction / pressure relaxation against the same
    expressions evaluated with numpy (element-wise IEEE operations in the reference's order)."""
    mesh = MESHES["skew"]()
    rng = np.random.default_rng(2)
    csr = oracle.create_csr(mesh)
    n, F, nt = mesh.numCells, mesh.numInnerFaces, mesh.numTotal
    L = device_layout(mesh, csr)
    geo = {k: np.ascontiguousarray(getattr(mesh, k), dtype=np.float64) for k in
           ("xc", "yc", "zc", "vol", "arx", "ary", "arz", "xf", "yf", "zf", "facint")}
    G = Geom(i32(L["owner"]), i32(L["neigh"]), *[d(geo[k]) for k in ("xc", "yc", "zc", "vol", "arx", "ary", "arz",
                                                                    "xf", "yf", "zf", "facint")], n, F)
    a, su, pp, p = rng.standard_normal(csr.nnz), rng.standard_normal(n), rng.standard_normal(nt), rng.standard_normal(nt)
    fl, dP = rng.standard_normal(F), rng.standard_normal((n, 3))
    u, v, w = rng.standard_normal(nt), rng.standard_normal(nt), rng.standard_normal(nt)
    apu, apv, apw = rng.random(n), rng.random(n), rng.random(n)
    icj = (csr.icell_jcell - 1).astype(np.int32)
    pref, urf = 17, 0.3
    a0, su0, fl0, u0, v0, w0, p0 = a.copy(), su.copy(), fl.copy(), u.copy(), v.copy(), w.copy(), p.copy()
    host.fcp_host_tail(C.byref(G), i32(L["ioffset"]), i32(L["diag"]), i32(icj), d(a), d(su), d(p0.copy()), pref, d(pp),
                       d(fl), d(apu), d(apv), d(apw), d(dP), d(u), d(v), d(w), C.c_double(urf), d(p))
    a0[L["ioffset"][pref]:L["ioffset"][pref + 1]] = 0.0
    a0[L["diag"][pref]] = 1.0
    su0[pref] = p0[pref]
    assert np.array_equal(a, a0) and np.array_equal(su, su0)
    o, nb = L["owner"][:F], L["neigh"]
    assert np.array_equal(fl, fl0 + a0[icj] * (pp[nb] - pp[o]))
    vol = mesh.vol[:n]
    assert np.array_equal(u[:n], u0[:n] - apu * dP[:, 0] * vol) and np.array_equal(u[n:], u0[n:])
    assert np.array_equal(v[:n], v0[:n] - apv * dP[:, 1] * vol)
    assert np.array_equal(w[:n], w0[:n] - apw * dP[:, 2] * vol)
    assert np.array_equal(p[:n], p0[:n] + urf * (pp[:n] - p0[:n])) and np.array_equal(p[n:], p0[n:])


# ---- least-squares gradients and slope limiters (fc_grad_body.cuh) ----
def geom_and_map(mesh, csr):
    L = device_layout(mesh, csr)
    geo = {k: np.ascontiguousarray(getattr(mesh, k), dtype=np.float64) for k in
           ("xc", "yc", "zc", "vol", "arx", "ary", "arz", "xf", "yf", "zf", "facint")}
    G = Geom(i32(L["owner"]), i32(L["neigh"]), *[d(geo[k]) for k in ("xc", "yc", "zc", "vol", "arx", "ary", "arz",
                                                                    "xf", "yf", "zf", "facint")],
             mesh.numCells, mesh.numInnerFaces)
    Mp = C2f(i32(L["off"]), i32(L["face"]), i32(L["other"]), i32(L["pos"]))
    return L, G, Mp, geo


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("weighted", [False, True])
def test_lsq_bodies_equal_oracle(host, name, weighted):
    mesh = MESHES[name]()
    csr = oracle.create_csr(mesh)
    L, G, Mp, _keep = geom_and_map(mesh, csr)
    n = mesh.numCells
    fi = cases.flow_fields(mesh)["p"]
    dmat, out = np.zeros((n, 9)), np.zeros((n, 3))
    host.fcg_host_lsq(C.byref(G), C.byref(Mp), C.byref(L["slots"]), int(weighted), d(dmat), d(fi), d(out))
    ref_m = oracle.lsq_matrix(mesh, weighted)
    assert np.array_equal(dmat, ref_m)
    assert np.array_equal(out, oracle.grad_lsq(mesh, weighted, ref_m, fi)[:n])


@pytest.mark.parametrize("name", ["skew", "hex_mixed_bc"])
def test_lsq_qr_bodies_equal_oracle(host, name):
    mesh = MESHES[name]()
    csr = oracle.create_csr(mesh)
    L, G, Mp, _keep = geom_and_map(mesh, csr)
    n = mesh.numCells
    fi = cases.flow_fields(mesh)["u"]
    D, out = np.zeros((n, 6, 3)), np.zeros((n, 3))
    bad = host.fcg_host_lsq_qr(C.byref(G), C.byref(Mp), d(D), d(fi), d(out))
    ref_D, ref_bad = oracle.lsq_qr_matrix(mesh)
    assert bad == ref_bad == 0
    assert np.array_equal(D, ref_D)
    assert np.array_equal(out, oracle.grad_lsq_qr(mesh, ref_D, fi)[:n])


def test_lsq_qr_body_flags_cells_without_six_neighbours(host):
    mesh = MESHES["poly"]()
    csr = oracle.create_csr(mesh)
    L, G, Mp, _keep = geom_and_map(mesh, csr)
    n = mesh.numCells
    D, out = np.ones((n, 6, 3)), np.zeros((n, 3))
    bad = host.fcg_host_lsq_qr(C.byref(G), C.byref(Mp), d(D), d(np.zeros(mesh.numTotal)), d(out))
    assert bad == oracle.lsq_qr_matrix(mesh)[1] > 0


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("which", ["Barth-Jespersen", "Venkatakrishnan", "mVenkatakrishnan"])
def test_limiter_bodies_equal_oracle(host, name, which):
    mesh = MESHES[name]()
    csr = oracle.create_csr(mesh)
    L, G, Mp, _keep = geom_and_map(mesh, csr)
    n = mesh.numCells
    phi = cases.flow_fields(mesh)["v"]
    g0 = oracle.grad_gauss(mesh, phi, 1)
    got = np.ascontiguousarray(g0[:n].copy())
    host.fcg_host_limiter(C.byref(G), i32(L["ioffset"]), i32(L["ja"]), i32(L["diag"]), oracle.LIMITERS[which], d(phi),
                          d(got), C.c_double(phi[:n].min()), C.c_double(phi[:n].max()), C.c_double(oracle.SMALL))
    ref = g0.copy()
    oracle.slope_limiter(mesh, csr, which, phi, ref)
    assert np.array_equal(got, ref[:n])


def gaussn(host, G, Mp, npro, fpro, fields, n, nigrad):
    nf = len(fields)
    out = [np.zeros((n, 3)) for _ in range(nf)]
    old = [np.zeros((n, 3)) for _ in range(nf)]
    dpp = C.POINTER(C.c_double) * nf
    host.fcg_host_gaussn(C.byref(G), C.byref(Mp), npro, d(fpro), nf, dpp(*[d(v) for v in fields]),
                         dpp(*[d(v) for v in out]), dpp(*[d(v) for v in old]), nigrad)
    return out


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("nigrad", [1, 2])
@pytest.mark.parametrize("keys", [("u", "v", "w"), ("u", "v", "w", "p")])
def test_fused_gauss_gradient_body_equals_single_field_oracle_passes(host, name, nigrad, keys):
    """fcg_gaussn_row (FC_TUNE_FUSED_GRAD): three or four gradients in one walk over the cell-to-face map, each
    bit-identical to grad_gauss of that field alone."""
    mesh = MESHES[name]()
    csr = oracle.create_csr(mesh)
    L, G, Mp, _keep = geom_and_map(mesh, csr)
    n = mesh.numCells
    f = cases.flow_fields(mesh)
    fields = [np.ascontiguousarray(f[k], dtype=np.float64) for k in keys]
    for got, k in zip(gaussn(host, G, Mp, 0, np.zeros(1), fields, n, nigrad), keys):
        assert np.array_equal(got, oracle.grad_gauss(mesh, f[k], nigrad)[:n]), (name, k)


@pytest.mark.parametrize("name,nranks", [("skew", 3), ("poly", 2)])
def test_fused_gauss_gradient_body_on_processor_faces(host, name, nranks):
    from oracle import oracle_par as OP
    mesh = cases.skew_case(6, 5, 7) if name == "skew" else MESHES["poly"]()
    f = cases.flow_fields(mesh)
    parts = M.partition(mesh, M.rcb_ranks(mesh, nranks) if name == "poly" else M.slab_ranks(mesh.numCells, nranks), nranks)
    pc = OP.ParCase(parts)
    keys = ("u", "v", "w", "p")
    phis = {k: [np.ascontiguousarray(M.scatter_total(mesh, p, f[k])) for p in parts] for k in keys}
    ref = {k: pc.grad_gauss(phis[k], 1) for k in keys}     # exchanges the halo of phi first
    for r, part in enumerate(parts):
        L, G, Mp, _keep = geom_and_map(part, pc.csr[r])
        n, npro = part.numCells, part.npro
        fpro = np.ascontiguousarray(part.fpro, dtype=np.float64) if npro else np.zeros(1)
        for got, k in zip(gaussn(host, G, Mp, npro, fpro, [phis[k][r] for k in keys], n, 1), keys):
            assert np.array_equal(got, np.asarray(ref[k][r]).reshape(-1, 3)[:n]), (name, r, k)


# ---- several ranks: processor faces, running-subtraction diagonal (src-parallel/calcuvw.f90) ----
@pytest.mark.parametrize("kw", [dict(), dict(bdf=True, btime=1.0, timestep=0.02, cn=True)])
@pytest.mark.parametrize("name,nranks", [("skew", 2), ("skew", 3), ("poly", 2)])
def test_bodies_equal_parallel_oracle(host, name, nranks, kw):
    """Every rank's kernel bodies (inner + processor faces, rows, the three component steps with the parallel
    diagonal) against the lock-step multi-rank oracle, bit for bit.  Halos, gradients and the boundary pressure are
    taken from the oracle (on the GPU the verified halo exchange / gradient kernels provide them)."""
    from oracle import oracle_par as OP
    mesh = cases.skew_case(6, 5, 7) if name == "skew" else MESHES["poly"]()
    f = cases.flow_fields(mesh)
    rng = np.random.default_rng(8)
    nt = mesh.numTotal
    cell_rank = M.rcb_ranks(mesh, nranks) if name == "poly" else M.slab_ranks(mesh.numCells, nranks)
    parts = M.partition(mesh, cell_rank, nranks)
    pc = OP.ParCase(parts)
    xs = pc.uvw_fields(0.0)
    gl = dict(vis=0.01 * (1.0 + 0.3 * rng.random(nt)))
    for k in ("uo", "vo", "wo", "uoo", "voo", "woo", "t"):
        gl[k] = rng.standard_normal(nt)
    fl_g = rng.standard_normal(mesh.numInnerFaces) * 1e-2
    for r, part in enumerate(parts):
        fr = pc.fields[r]
        for k in ("u", "v", "w", "p", "den"):
            getattr(fr, k)[:] = M.scatter_total(mesh, part, f[k])
        gf = part.face_global
        fr.flmass[:] = fl_g[gf[:part.numInnerFaces]]
        pf = gf[part.iProcFacesStart:part.iProcFacesStart + part.npro]
        pc.fmpro[r][:part.npro] = fl_g[pf] * np.sign(part.arx[part.iProcFacesStart:] * mesh.arx[pf] +
                                                     part.ary[part.iProcFacesStart:] * mesh.ary[pf] +
                                                     part.arz[part.iProcFacesStart:] * mesh.arz[pf])
        fr.fmi[:] = rng.standard_normal(fr.fmi.size) * 1e-2
        fr.fmo[:] = rng.standard_normal(fr.fmo.size) * 1e-2
        fr.a[:] = rng.standard_normal(fr.a.size)
        for k in gl:
            getattr(xs[r], k)[:] = M.scatter_total(mesh, part, gl[k])
    opts = oracle.uvw_opts(scheme="muscl-f", urf=(0.7, 0.8, 0.6), nsw=(0, 0, 0), **kw)
    a_in = [fr.a.copy() for fr in pc.fields]
    pc.calcuvw_assemble(opts)
    keep = []
    for r, part in enumerate(parts):
        fr, x, csr = pc.fields[r], xs[r], pc.csr[r]
        n, F, npro = part.numCells, part.numInnerFaces, part.npro
        L, G, Mp, geo = geom_and_map(part, csr)
        den = np.ascontiguousarray(fr.den)
        FL = Flow(d(fr.u), d(fr.v), d(fr.w), d(fr.p), d(den), d(x.vis), d(fr.flmass), d(fr.fmi), d(fr.fmo), d(fr.dUdxi),
                  d(fr.dVdxi), d(fr.dWdxi), d(fr.dPdxi), d(x.uo), d(x.vo), d(x.wo), d(x.uoo), d(x.voo), d(x.woo), d(x.t))
        O = Opts(opts.scheme, opts.limiter, opts.gds, opts.bdf, opts.btime, opts.timestep, opts.cn, opts.const_mflux,
                 opts.gradPcmf, opts.lbuoy, opts.boussinesq, opts.beta, opts.tref, opts.densit, opts.gravx, opts.gravy,
                 opts.gravz, opts.viscos)
        fa = [np.zeros(max(F, 1)) for _ in range(6)]
        pa = [np.zeros(max(npro, 1)) for _ in range(5)]     # apr, sup, svp, swp, fie
        rows = {k: np.zeros(n) for k in ("su", "sv", "sw", "spu", "spv", "sp")}
        a = a_in[r].copy()
        fpro = np.ascontiguousarray(part.fpro, dtype=np.float64) if npro else np.zeros(1)
        P = Proc(npro, part.iProcFacesStart, d(fpro), d(pc.fmpro[r]), *[d(v) for v in pa])
        host.fcm_host_assemble(C.byref(G), C.byref(Mp), C.byref(L["slots"]), C.byref(FL), C.byref(O),
                               C.byref(Faces(*[d(v) for v in fa])), C.byref(P),
                               C.byref(Rows(d(a), *[d(rows[k]) for k in ("su", "sv", "sw", "spu", "spv", "sp")])), csr.nnz)
        for k, ref in (("su", fr.su), ("sv", x.sv), ("sw", x.sw), ("spu", x.spu), ("spv", x.spv), ("sp", x.sp)):
            assert np.array_equal(rows[k], ref), (r, k, np.abs(rows[k] - ref).max())
        assert np.array_equal(a, fr.a)
        assert np.array_equal(pa[0][:npro], pc.apr[r][:npro])
        keep.append((L, G, Mp, geo, den, rows, a, pa, fpro))
    # component steps: all ranks of one component, then the next (the oracle's zero-sweep solve leaves u, v, w alone)
    for comp, (sk, spk, phik, oldk, apk) in enumerate((("su", "spu", "u", "uo", "apu"), ("sv", "spv", "v", "vo", "apv"),
                                                       ("sw", "sp", "w", "wo", "apw"))):
        got = []
        for r, part in enumerate(parts):
            L, G, Mp, geo, den, rows, a, pa, fpro = keep[r]
            fr, x = pc.fields[r], xs[r]
            n = part.numCells
            ap = np.zeros(n)
            K = Comp(i32(L["ioffset"]), i32(L["diag"]), d(a), d(rows[sk]), d(rows[spk]), d(rows["su"]), d(ap),
                     d(getattr(fr, phik)), d(getattr(x, oldk)), d(den), 1.0 / opts.urf[comp], 1.0 - opts.urf[comp],
                     opts.sol.small, opts.timestep, opts.cn, 1 if comp else 0, 1, part.npro, d(pa[0]))
            host.fcm_host_component(C.byref(G), C.byref(Mp), C.byref(K))
            got.append(ap)
        pc.calcuvw_component(opts, comp)
        for r, part in enumerate(parts):
            L, G, Mp, geo, den, rows, a, pa, fpro = keep[r]
            n = part.numCells
            assert np.array_equal(a, pc.fields[r].a), (comp, r)
            assert np.array_equal(rows["su"], pc.fields[r].su), (comp, r)
            assert np.array_equal(got[r], getattr(xs[r], apk)[:n]), (comp, r)


# ---- edge cases: smallest mesh the predictor accepts, one-cell-thick slab, a single boundary kind ----
@pytest.mark.parametrize("make", [lambda: cases.hex_case(3, 1, 1), lambda: cases.hex_case(20, 20, 1, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry")),
                                  lambda: cases.hex_case(4, 3, 2, kinds=("symmetry",) * 6), lambda: cases.hex_case(1, 1, 1)])
def test_bodies_equal_oracle_edge_meshes(host, make):
    mesh = make()
    f = cases.flow_fields(mesh)
    if mesh.numCells < 3 and mesh.numInnerFaces > 0:
        pytest.skip("outside the reference's df(ijp,3) addressing")
    assert run_case(host, mesh, f, "muscl-f", bdf=True, btime=1.0, timestep=0.05)
