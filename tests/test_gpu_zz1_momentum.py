"""GPU parity of the momentum predictor `call calcuvw` through the C ABI.

The kernels' per-index bodies are also checked bit for bit against the oracle on the CPU
(tests/test_momentum_bodies.py).  First hardware run: profiles/r01_momentum_gpu_tests.txt (37 passed).

Bars: explicit sources / coefficients bit-exact (same summation order, no FMA); after the BiCGStab solves
iteration counts within +-1 and fields within 1e-10 relative L2 when the counts agree.
"""
import os

import numpy as np
import pytest

from freecappuccino_b200 import cases
from oracle import oracle

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MESHES = {
    "skew": lambda: cases.skew_case(),
    "hex_mixed_bc": lambda: cases.hex_case(9, 7, 11, kinds=("inlet", "outlet", "wall", "symmetry", "wall", "prOutlet")),
    "poly": lambda: cases.poly_case(5),
    "cavity": lambda: cases.golden_mesh(os.path.join(GOLD, "cavity.npz")),
    "pitzDaily": lambda: cases.golden_mesh(os.path.join(GOLD, "pitzDaily.npz")),
}


@pytest.fixture(scope="module")
def fc():
    from freecappuccino_b200 import lib
    return lib


def linear_fluxes(mesh, of, flip=True):
    F = mesh.numInnerFaces
    o, nb = mesh.owner[:F] - 1, mesh.neighbour - 1
    fx = mesh.facint
    df = of.den[o] * (1 - fx) + of.den[nb] * fx
    fl = df * ((of.u[o] * (1 - fx) + of.u[nb] * fx) * mesh.arx[:F] + (of.v[o] * (1 - fx) + of.v[nb] * fx) * mesh.ary[:F]
               + (of.w[o] * (1 - fx) + of.w[nb] * fx) * mesh.arz[:F])
    if flip:
        fl[::3] *= -1.0
    return fl


def make_state(mesh, f, seed=3, stale=True):
    """Oracle-side state of one calcuvw call (fields of module variables / sparse_matrix)."""
    rng = np.random.default_rng(seed)
    csr = oracle.create_csr(mesh)
    nt = mesh.numTotal
    of = oracle.Fields(mesh, csr.nnz)
    for k in ("u", "v", "w", "p", "den"):
        getattr(of, k)[:] = f[k]
    of.flmass[:] = linear_fluxes(mesh, of, flip=stale)
    fmi, flomas = cases.inlet_fluxes(mesh, f)
    of.fmi[:fmi.size] = fmi
    fs, sl = mesh.boundary_faces("outlet"), mesh.boundary_slots("outlet")
    of.fmo[:len(fs)] = f["den"][sl] * (f["u"][sl] * mesh.arx[fs] + f["v"][sl] * mesh.ary[fs] + f["w"][sl] * mesh.arz[fs])
    if stale:
        of.a[:] = rng.standard_normal(csr.nnz)
    x = oracle.UvwFields(mesh, of, 0.0)
    x.vis[:] = 0.01 * (1.0 + 0.3 * rng.random(nt))
    for k in ("uo", "vo", "wo", "uoo", "voo", "woo", "t"):
        getattr(x, k)[:] = rng.standard_normal(nt)
    return csr, of, x, flomas


def upload_state(ctx, mesh, of, x):
    for name, arr in (("U", of.u), ("V", of.v), ("W", of.w), ("P", of.p), ("DEN", of.den), ("FLMASS", of.flmass),
                      ("A", of.a), ("DPDXI", of.dPdxi), ("VIS", x.vis), ("UO", x.uo), ("VO", x.vo), ("WO", x.wo),
                      ("UOO", x.uoo), ("VOO", x.voo), ("WOO", x.woo), ("T", x.t)):
        ctx.upload(name, arr)
    if mesh.count("inlet"):
        ctx.upload("FMI", of.fmi[:mesh.count("inlet")])
    if mesh.count("outlet"):
        ctx.upload("FMO", of.fmo[:mesh.count("outlet")])


def both_opts(fc, **kw):
    return oracle.uvw_opts(**kw), fc.calcuvw_opts(**kw)


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("scheme", ["muscl-f", "linear-f", "central", "smart", "charm"])
def test_calcuvw_assembly_bit_exact(fc, name, scheme):
    mesh = MESHES[name]()
    csr, of, x, _ = make_state(mesh, cases.flow_fields(mesh))
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    upload_state(ctx, mesh, of, x)
    oo, go = both_opts(fc, scheme=scheme, urf=(0.7, 0.8, 0.6), bdf=True, btime=1.0, timestep=0.02)
    oracle.calcuvw_assemble(mesh, csr, of, x, oo)
    ctx.calcuvw_assemble(go)
    n = mesh.numCells
    for fld, ref in (("P", of.p), ("DPDXI", of.dPdxi.ravel()), ("DUDXI", of.dUdxi.ravel()), ("SU", of.su), ("SV", x.sv),
                     ("SW", x.sw), ("SPU", x.spu), ("SPV", x.spv), ("SP", x.sp), ("A", of.a)):
        got = ctx.download(fld)[:ref.size]
        assert np.array_equal(got, ref), (fld, float(np.abs(got - ref).max()))
    # per component: diagonal, under-relaxation, ap* (zero BiCGStab sweeps on both sides)
    oo0, go0 = both_opts(fc, scheme=scheme, urf=(0.7, 0.8, 0.6), bdf=True, btime=1.0, timestep=0.02, nsw=(0, 0, 0))
    for comp, apn in enumerate(("APU", "APV", "APW")):
        oracle.calcuvw_component(mesh, csr, of, x, oo0, comp)
        ctx.calcuvw_component(go0, comp)
        assert np.array_equal(ctx.download("A"), of.a), comp
        assert np.array_equal(ctx.download("SU"), of.su), comp
        assert np.array_equal(ctx.download(apn)[:n], (x.apu, x.apv, x.apw)[comp][:n]), comp
    ctx.close()


@pytest.mark.parametrize("kw", [dict(cn=True, bdf=True, btime=1.0, timestep=0.02), dict(const_mflux=True, gradPcmf=0.37),
                                dict(lbuoy=True, boussinesq=True, beta=0.3, tref=0.1, densit=1.1, grav=(0.1, -9.81, 0.2)),
                                dict(lbuoy=True, boussinesq=False, densit=1.1, grav=(0.1, -9.81, 0.2))])
def test_calcuvw_source_options_bit_exact(fc, kw):
    mesh = MESHES["skew"]()
    csr, of, x, _ = make_state(mesh, cases.flow_fields(mesh))
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    upload_state(ctx, mesh, of, x)
    oo, go = both_opts(fc, nsw=(0, 0, 0), **kw)
    oracle.calcuvw(mesh, csr, of, x, oo)
    ctx.calcuvw(go)
    n = mesh.numCells
    for fld, ref in (("SV", x.sv), ("SW", x.sw), ("SPU", x.spu), ("SPV", x.spv), ("SP", x.sp), ("A", of.a),
                     ("SU", of.su), ("APU", x.apu[:n]), ("APV", x.apv[:n]), ("APW", x.apw[:n])):
        got = ctx.download(fld)[:ref.size]
        assert np.array_equal(got, ref), (fld, float(np.abs(got - ref).max()))
    ctx.close()


@pytest.mark.parametrize("name,sor,ftol", [("skew", 1e-8, 1e-7), ("cavity", 1e-8, 1e-6), ("pitzDaily", 1e-6, 1e-4),
                                           ("poly", 1e-8, 1e-7)])
def test_calcuvw_full_parity(fc, name, sor, ftol):
    """`call calcuvw` end to end: BiCGStab iteration counts within +-1; u/v/w to the accuracy the stopping
    tolerance fixes them (a solve stopped at rsm < sor leaves an O(sor) difference when the counts agree and the
    last residuals differ in the final digits); ap* do not depend on the solves and are exact."""
    mesh = MESHES[name]()
    csr, of, x, _ = make_state(mesh, cases.channel_fields(mesh), stale=False)
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    upload_state(ctx, mesh, of, x)
    oo, go = both_opts(fc, scheme="muscl-f", sor=(sor,) * 3, nsw=(200,) * 3, bdf=True, timestep=0.05)
    rr = oracle.calcuvw(mesh, csr, of, x, oo)
    rg = ctx.calcuvw(go)
    n = mesh.numCells
    same = True
    for k in range(3):
        assert 0 < rr.rep[k].iters < 200
        assert abs(rg.rep[k].iters - rr.rep[k].iters) <= 1, (k, rg.rep[k].iters, rr.rep[k].iters)
        assert rg.rep[k].res0 == pytest.approx(rr.rep[k].res0, rel=1e-9)
        same = same and rg.rep[k].iters == rr.rep[k].iters
    assert np.array_equal(ctx.download("APU")[:n], x.apu[:n])
    if same:
        for fld, ref in (("U", of.u), ("V", of.v), ("W", of.w)):
            assert cases.rel_l2(ctx.download(fld)[:n], ref[:n]) < ftol, fld
    ctx.close()


def test_simple_iterations_device_resident_cavity(fc):
    """Config 1 as shipped (examples/cavity/input: muscl-f, gauss, urf 0.7/0.3, sor 1e-2, nsw 20/100, bdf with
    timestep 1e20, lid U = 1): ten SIMPLE iterations calcuvw -> calcp with every field staying on the GPU,
    against the oracle doing the same.  Iteration counts within +-1 at every step; fields compared while the
    counts agree (a solve stopped at rsm < 1e-2 fixes the fields only to that tolerance otherwise)."""
    mesh = MESHES["cavity"]()
    nt = mesh.numTotal
    f = dict(u=np.zeros(nt), v=np.zeros(nt), w=np.zeros(nt), p=np.zeros(nt), den=np.ones(nt))
    # the lid: first wall patch of the fixture = 'movingWall' (20 faces, U = (1,0,0)), examples/cavity/0/U
    f["u"][mesh.boundary_slots("wall")[:20]] = 1.0
    csr, of, x, _ = make_state(mesh, f, stale=False)
    x.vis[:] = 0.01
    for k in ("uo", "vo", "wo", "uoo", "voo", "woo", "t"):
        getattr(x, k)[:] = 0.0
    of.flmass[:] = 0.0
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    upload_state(ctx, mesh, of, x)
    kw_u = dict(scheme="muscl-f", urf=(0.7, 0.7, 0.7), sor=(1e-2,) * 3, nsw=(20,) * 3, bdf=True, btime=0.0,
                timestep=1e20, viscos=0.01)
    kw_p = dict(solver="dpcg", sor=1e-2, nsw=100, urf_p=0.3, pRefCell=1, const_mflux=True)
    oo, go = both_opts(fc, **kw_u)
    agree = True
    for it in range(10):
        ru = oracle.calcuvw(mesh, csr, of, x, oo)
        gu = ctx.calcuvw(go)
        rp = oracle.calcp(mesh, csr, of, oracle.calcp_opts(**kw_p))
        gp = ctx.calcp(fc.calcp_opts(**kw_p))
        for k in range(3):
            assert abs(gu.rep[k].iters - ru.rep[k].iters) <= 1, (it, k, gu.rep[k].iters, ru.rep[k].iters)
            agree = agree and gu.rep[k].iters == ru.rep[k].iters
        assert abs(gp.rep[0].iters - rp.rep[0].iters) <= 1, (it, gp.rep[0].iters, rp.rep[0].iters)
        agree = agree and gp.rep[0].iters == rp.rep[0].iters
        if agree:
            for fld, ref in (("U", of.u), ("V", of.v), ("P", of.p), ("FLMASS", of.flmass)):
                assert cases.rel_l2(ctx.download(fld)[:ref.size], ref) < 1e-9, (it, fld)
    assert np.abs(of.u[:mesh.numCells]).max() > 0.05   # the lid drives a flow
    ctx.close()


def test_calcuvw_host_form(fc):
    mesh = MESHES["skew"]()
    csr, of, x, _ = make_state(mesh, cases.channel_fields(mesh), stale=False)
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    upload_state(ctx, mesh, of, x)
    oo, go = both_opts(fc, sor=(1e-8,) * 3, nsw=(200,) * 3, bdf=True, timestep=0.05)
    u, v, w, p = of.u.copy(), of.v.copy(), of.w.copy(), of.p.copy()
    n = mesh.numCells
    apu, apv, apw = np.zeros(n), np.zeros(n), np.zeros(n)
    rg = ctx.calcuvw_host(go, u, v, w, p, x.vis, of.flmass, apu, apv, apw)
    rr = oracle.calcuvw(mesh, csr, of, x, oo)
    same = True
    for k in range(3):
        assert abs(rg.rep[k].iters - rr.rep[k].iters) <= 1
        same = same and rg.rep[k].iters == rr.rep[k].iters
    assert np.array_equal(p, of.p)
    assert np.array_equal(apu, x.apu[:n]) and np.array_equal(apw, x.apw[:n])
    if same:
        for got, ref in ((u, of.u), (v, of.v), (w, of.w)):
            assert cases.rel_l2(got[:n], ref[:n]) < 1e-7
    ctx.close()


def test_calcuvw_rejects_what_it_does_not_cover(fc):
    mesh = cases.hex_case(2, 1, 1)
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    with pytest.raises(fc.FcError) as e:
        ctx.calcuvw(fc.calcuvw_opts())
    assert e.value.code == fc.FC_ERR_UNSUPPORTED   # numCells < 3: the reference's df(ijp,3) leaves the array
    ctx.close()


def test_cavity_driver_simple_loop_matches_oracle():
    """host/cavity: the SIMPLE loop of src/main.f90 (call calcuvw; call calcp) as a compiled program against
    libfcapp_cuda.so with the shipped cavity settings; every solver report line it prints is compared with the
    oracle running the same loop (iteration counts within +-1, printed residuals to the printed digits while the
    counts agree)."""
    import re
    import subprocess
    from freecappuccino_b200 import mesh as M
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "host", "cavity")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(root, "host")])
    n, outer = 20, 12
    out = subprocess.run([exe, str(n), str(outer), "1e-30"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = re.compile(r"Solving for (\w+), Initial residual = +(\d\.\d{3}E[+-]\d\d), Final residual = +"
                      r"(\d\.\d{3}E[+-]\d\d), No Iterations (\d+)$", re.M)
    got = [(m.group(1), float(m.group(2)), int(m.group(4))) for m in line.finditer(out.stdout)]
    assert len(got) == 4 * outer, out.stdout[-2000:]
    its, source, umax = out.stdout.strip().splitlines()[-1].split()
    assert int(its) == outer

    mesh = M.hex_mesh(n, n, 1, (0.1, 0.1, 0.01), ("wall", "wall", "wall", "wall", "symmetry", "symmetry"))
    nt = mesh.numTotal
    f = dict(u=np.zeros(nt), v=np.zeros(nt), w=np.zeros(nt), p=np.zeros(nt), den=np.ones(nt))
    fs, sl = mesh.boundary_faces("wall"), mesh.boundary_slots("wall")
    f["u"][sl[mesh.ary[fs] > 0.0]] = 1.0
    csr, of, x, _ = make_state(mesh, f, stale=False)
    x.vis[:] = 0.01
    for k in ("uo", "vo", "wo", "uoo", "voo", "woo", "t"):
        getattr(x, k)[:] = 0.0
    of.flmass[:] = 0.0
    oo = oracle.uvw_opts(scheme="muscl-f", urf=(0.7,) * 3, sor=(1e-2,) * 3, nsw=(20,) * 3, bdf=True, btime=0.0,
                         timestep=1e20, viscos=0.01)
    po = oracle.calcp_opts(solver="iccg", sor=1e-2, nsw=100, urf_p=0.3, pRefCell=1, const_mflux=False, flomas=0.0)
    agree = True
    for it in range(outer):
        ru = oracle.calcuvw(mesh, csr, of, x, oo)
        rp = oracle.calcp(mesh, csr, of, po)
        ref = [("U", ru.rep[0]), ("V", ru.rep[1]), ("W", ru.rep[2]), ("p", rp.rep[0])]
        for k, (name, r) in enumerate(ref):
            g = got[4 * it + k]
            assert g[0] == name
            assert abs(g[2] - r.iters) <= 1, (it, name, g[2], r.iters)
            agree = agree and g[2] == r.iters
            if agree and r.res0 > 0:
                assert g[1] == pytest.approx(r.res0, rel=2e-3), (it, name)
    if agree:
        assert float(umax) == pytest.approx(float(np.abs(of.u[:mesh.numCells]).max()), rel=1e-3)


def test_cavity_driver_on_the_shipped_polymesh(tmp_path):
    """host/cavity fed with the polyMesh of examples/cavity (read by the C++ mesh_geometry of host/fcapp_mesh.cpp)
    prints the same solver report lines as with its generated 20 x 20 x 1 box: same mesh, two routes."""
    import re
    import subprocess
    from test_polymesh_reader import write_polymesh
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "host", "cavity")
    d = np.load(os.path.join(GOLD, "cavity.npz"))
    patches = [(str(k), int(nf), int(st)) for k, nf, st in zip(d["bkind"], d["bn"], d["bstart"])]
    pm = os.path.join(str(tmp_path), "polyMesh")
    write_polymesh(pm, d["points"], d["faces"], d["owner"], d["neighbour"], patches)
    line = re.compile(r"Solving for (\w+), Initial residual = +(\d\.\d{3}E[+-]\d\d), Final residual = +"
                      r"(\d\.\d{3}E[+-]\d\d), No Iterations (\d+)$", re.M)
    outs = []
    for arg in (pm, "20"):
        out = subprocess.run([exe, arg, "6", "1e-30"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        outs.append([(m.group(1), m.group(2), int(m.group(4))) for m in line.finditer(out.stdout)])
    assert len(outs[0]) == 24
    for a, b in zip(*outs):
        assert a[0] == b[0] and abs(a[2] - b[2]) <= 1
        if a[2] == b[2]:
            assert float(a[1]) == pytest.approx(float(b[1]), rel=5e-3, abs=1e-12)


@pytest.mark.parametrize("make", [lambda: cases.hex_case(3, 1, 1),
                                  lambda: cases.hex_case(20, 20, 1, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry")),
                                  lambda: cases.hex_case(4, 3, 2, kinds=("symmetry",) * 6)])
def test_calcuvw_edge_meshes_bit_exact(fc, make):
    """Smallest mesh the predictor accepts (three cells in a row), a one-cell-thick slab, a single boundary kind."""
    mesh = make()
    csr, of, x, _ = make_state(mesh, cases.flow_fields(mesh))
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    upload_state(ctx, mesh, of, x)
    oo, go = both_opts(fc, bdf=True, btime=1.0, timestep=0.05, nsw=(0, 0, 0))
    oracle.calcuvw(mesh, csr, of, x, oo)
    ctx.calcuvw(go)
    n = mesh.numCells
    for fld, ref in (("SV", x.sv), ("SW", x.sw), ("SPU", x.spu), ("SP", x.sp), ("A", of.a), ("SU", of.su),
                     ("APU", x.apu[:n]), ("APW", x.apw[:n])):
        got = ctx.download(fld)[:ref.size]
        assert np.array_equal(got, ref), (fld, float(np.abs(got - ref).max()))
    ctx.close()


@pytest.mark.parametrize("solver", ["dpcg", "iccg", "bicgstab"])
def test_named_solver_entry_points(fc, solver):
    """fc_dpcg / fc_iccg / fc_bicgstab (the reference's subroutine names) are fc_solve with the solver fixed."""
    mesh = cases.hex_case(8, 7, 6, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry"))
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    su = cases.poisson_rhs(mesh)
    ctx.upload("APU", -np.ones(mesh.numCells))
    out = []
    for call in (lambda o: ctx.solve(solver, "PP", o), lambda o: getattr(ctx, solver)("PP", o)):
        ctx.upload("SU", su)
        ctx.fill("PP", 0.0)
        ctx.laplacian("APU", "PP")
        rep = call(fc.solver_opts(1e-8, 500))
        out.append((rep.iters, rep.res0, rep.resl, ctx.download("PP")))
    assert out[0][0] == out[1][0] > 0 and out[0][1] == out[1][1] and out[0][2] == out[1][2]
    assert np.array_equal(out[0][3], out[1][3])
    ctx.close()


def test_gpu_simple_loop_reproduces_ghia_re100(fc):
    """The device-resident SIMPLE loop (fc_calcuvw + fc_calcp, no field leaves the GPU between iterations) run to
    convergence on the 20 x 20 lid-driven cavity at Re = 100, compared directly with the benchmark table of Ghia, Ghia &
    Shin (1982) -- an oracle-independent check of the CUDA path -- and with the oracle's converged field."""
    from test_oracle_ghia import cavity_mesh, centreline_error, lid_slots, simple_to_convergence
    n = 20
    mesh = cavity_mesh(n)
    nt = mesh.numTotal
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    u0 = np.zeros(nt)
    u0[lid_slots(mesh)] = 1.0
    ctx.upload("U", u0)
    ctx.upload("VIS", np.full(nt, 0.01))
    uo = fc.calcuvw_opts(scheme="muscl-f", urf=(0.7,) * 3, sor=(1e-2,) * 3, nsw=(20,) * 3, bdf=True, btime=0.0,
                         timestep=1e20, viscos=0.01)
    po = fc.calcp_opts(solver="iccg", sor=1e-2, nsw=100, urf_p=0.3, pRefCell=1)
    for it in range(1, 3001):
        ru = ctx.calcuvw(uo)
        rp = ctx.calcp(po)
        if max(ru.rep[0].res0, ru.rep[1].res0, rp.rep[0].res0) < 1e-7:
            break
    u = ctx.download("U")
    err, umin = centreline_error(mesh, n, u)
    assert it < 3000 and err < 0.02 and abs(umin + 0.2109) < 0.01
    _, of, it_o = simple_to_convergence(n)
    assert abs(it - it_o) <= max(10, it_o // 50)      # hundreds of loosely converged solves: a few iterations of slack
    assert cases.rel_l2(u[:mesh.numCells], of.u[:mesh.numCells]) < 1e-4   # both stopped at source < 1e-7
    ctx.close()
