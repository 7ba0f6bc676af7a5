"""GPU parity of the PISO / PIMPLE pressure equation (fc_piso) through the C ABI.

Written after this round's GPU budget was spent: the new kernels' bodies are checked bit for bit against the oracle
on the CPU (tests/test_momentum_bodies.py); everything else the driver launches is the already verified
pressure-correction path.  These tests are the first run on hardware; the file name sorts last so that the verified
suites run first.
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
    "cavity": lambda: cases.golden_mesh(os.path.join(GOLD, "cavity.npz")),
    "poly": lambda: cases.poly_case(5),
}


@pytest.fixture(scope="module")
def fc():
    from freecappuccino_b200 import lib
    return lib


def start_state(mesh, seed=5):
    """Oracle-side state right after a momentum predictor: momentum matrix in `a`, ap*, u, v, w, fluxes."""
    rng = np.random.default_rng(seed)
    f = cases.channel_fields(mesh)
    csr = oracle.create_csr(mesh)
    nt, n, F = mesh.numTotal, mesh.numCells, mesh.numInnerFaces
    of = oracle.Fields(mesh, csr.nnz)
    for k in ("u", "v", "w", "p", "den"):
        getattr(of, k)[:] = f[k]
    o, nb = mesh.owner[:F] - 1, mesh.neighbour - 1
    fx = mesh.facint
    of.flmass[:] = (of.den[o] * (1 - fx) + of.den[nb] * fx) * (
        (of.u[o] * (1 - fx) + of.u[nb] * fx) * mesh.arx[:F] + (of.v[o] * (1 - fx) + of.v[nb] * fx) * mesh.ary[:F]
        + (of.w[o] * (1 - fx) + of.w[nb] * fx) * mesh.arz[:F])
    fmi, flomas = cases.inlet_fluxes(mesh, f)
    of.fmi[:fmi.size] = fmi
    x = oracle.UvwFields(mesh, of, 0.01)
    x.uo[:], x.vo[:], x.wo[:] = of.u, of.v, of.w
    x.uoo[:] = of.u + 0.01 * rng.standard_normal(nt)
    x.voo[:] = of.v + 0.01 * rng.standard_normal(nt)
    x.woo[:] = of.w + 0.01 * rng.standard_normal(nt)
    x.t[:] = rng.standard_normal(nt)
    oo = oracle.uvw_opts(scheme="muscl-f", urf=(1.0,) * 3, sor=(1e-8,) * 3, nsw=(100,) * 3, bdf=True, timestep=0.01)
    oracle.calcuvw(mesh, csr, of, x, oo)
    of.pp[:] = of.p
    return csr, of, x, flomas


def upload_state(ctx, mesh, of, x):
    n = mesh.numCells
    for name, arr in (("U", of.u), ("V", of.v), ("W", of.w), ("P", of.p), ("PP", of.pp), ("DEN", of.den),
                      ("FLMASS", of.flmass), ("A", of.a), ("DPDXI", of.dPdxi), ("APU", of.apu), ("APV", of.apv),
                      ("APW", of.apw), ("UO", x.uo), ("VO", x.vo), ("WO", x.wo), ("UOO", x.uoo), ("VOO", x.voo),
                      ("WOO", x.woo), ("T", x.t)):
        ctx.upload(name, arr)
    if mesh.count("inlet"):
        ctx.upload("FMI", of.fmi[:mesh.count("inlet")])
    if mesh.count("outlet"):
        ctx.upload("FMO", of.fmo[:mesh.count("outlet")])


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("kw", [dict(pimple=False, bdf=True, btime=1.0, timestep=0.01),
                                dict(pimple=True, urf_p=0.6, bdf=True, btime=0.0, timestep=0.01, cn=True),
                                dict(pimple=False, bdf=False, lbuoy=True, boussinesq=False, densit=1.1,
                                     grav=(0.1, -9.81, 0.2))])
def test_piso_without_krylov_sweeps_is_bit_exact(fc, name, kw):
    """nsw(ip) = 0: get_rAU_x_UEqnH, the facefluxmass_piso assembly, the reference-row pinning, the flux /
    velocity / pressure updates and the boundary corrections run with no Krylov iteration in between, so every
    field must equal the oracle's bit for bit."""
    mesh = MESHES[name]()
    csr, of, x, flomas = start_state(mesh)
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    upload_state(ctx, mesh, of, x)
    pref = min(7, mesh.numCells)
    kw = dict(kw, ncorr=2, npcor=2, pRefCell=pref, flomas=flomas, nsw=0, sor=1e-2)
    rr = oracle.piso(mesh, csr, of, x, oracle.piso_opts(**kw))
    rg = ctx.piso(fc.piso_opts(**kw))
    assert rg.nsolves == rr.nsolves == 4
    n = mesh.numCells
    for fld, ref in (("SU", of.su), ("SV", x.sv), ("SW", x.sw), ("A", of.a), ("U", of.u), ("V", of.v), ("W", of.w),
                     ("P", of.p), ("PP", of.pp), ("FLMASS", of.flmass), ("DPDXI", of.dPdxi.ravel())):
        got = ctx.download(fld)[:ref.size]
        assert np.array_equal(got, ref), (fld, float(np.abs(got - ref).max()))
    assert rg.sumLocalContErr == pytest.approx(rr.sumLocalContErr, rel=1e-9, abs=1e-300)
    for k in range(4):
        assert rg.rep[k].res0 == pytest.approx(rr.rep[k].res0, rel=1e-11)
    ctx.close()


@pytest.mark.parametrize("name,pimple", [("skew", False), ("skew", True), ("hex_mixed_bc", False), ("cavity", False),
                                         ("poly", True)])
def test_piso_full_parity(fc, name, pimple):
    mesh = MESHES[name]()
    csr, of, x, flomas = start_state(mesh)
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    upload_state(ctx, mesh, of, x)
    kw = dict(ncorr=2, npcor=1, pRefCell=3, pimple=pimple, urf_p=0.7, flomas=flomas, bdf=True, timestep=0.01,
              sor=1e-8, nsw=500)
    rr = oracle.piso(mesh, csr, of, x, oracle.piso_opts(**kw))
    rg = ctx.piso(fc.piso_opts(**kw))
    assert rg.nsolves == rr.nsolves == 2
    same = True
    for k in range(2):
        assert 0 < rr.rep[k].iters < 500
        assert abs(rg.rep[k].iters - rr.rep[k].iters) <= 1, (k, rg.rep[k].iters, rr.rep[k].iters)
        same = same and rg.rep[k].iters == rr.rep[k].iters
    assert rg.rep[0].res0 == pytest.approx(rr.rep[0].res0, rel=1e-10)
    if same:
        n = mesh.numCells
        for fld, ref in (("U", of.u), ("V", of.v), ("W", of.w), ("P", of.p), ("FLMASS", of.flmass)):
            assert cases.rel_l2(ctx.download(fld)[:ref.size], ref) < 1e-6, fld
    ctx.close()


def test_transient_cavity_piso_device_resident(fc):
    """Lid-driven cavity from rest, eight time steps of calcuvw + PISO (two correctors) with every field resident
    on the GPU (old time levels shifted with device-to-device copies), in lock step with the oracle."""
    mesh = MESHES["cavity"]()
    nt, n = mesh.numTotal, mesh.numCells
    csr = oracle.create_csr(mesh)
    of = oracle.Fields(mesh, csr.nnz)
    of.u[mesh.boundary_slots("wall")[:20]] = 1.0
    x = oracle.UvwFields(mesh, of, 0.01)
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    ctx.upload("U", of.u)
    ctx.upload("VIS", x.vis)
    dt = 5e-3
    kw_u = dict(scheme="muscl-f", urf=(1.0,) * 3, sor=(1e-6,) * 3, nsw=(50,) * 3, bdf=True, btime=0.0, timestep=dt,
                viscos=0.01)
    kw_p = dict(ncorr=2, npcor=1, pRefCell=1, const_mflux=True, bdf=True, btime=0.0, timestep=dt, sor=1e-6, nsw=300)
    agree = True
    for step in range(8):
        x.uoo[:], x.voo[:], x.woo[:] = x.uo, x.vo, x.wo
        x.uo[:], x.vo[:], x.wo[:] = of.u, of.v, of.w
        for src, dst in (("UO", "UOO"), ("VO", "VOO"), ("WO", "WOO"), ("U", "UO"), ("V", "VO"), ("W", "WO")):
            ctx.copy(src, dst)
        ru = oracle.calcuvw(mesh, csr, of, x, oracle.uvw_opts(**kw_u))
        gu = ctx.calcuvw(fc.calcuvw_opts(**kw_u))
        rp = oracle.piso(mesh, csr, of, x, oracle.piso_opts(**kw_p))
        gp = ctx.piso(fc.piso_opts(**kw_p))
        for k in range(3):
            assert abs(gu.rep[k].iters - ru.rep[k].iters) <= 1, (step, k)
            agree = agree and gu.rep[k].iters == ru.rep[k].iters
        for k in range(2):
            assert abs(gp.rep[k].iters - rp.rep[k].iters) <= 1, (step, k, gp.rep[k].iters, rp.rep[k].iters)
            agree = agree and gp.rep[k].iters == rp.rep[k].iters
        if agree:
            for fld, ref in (("U", of.u), ("V", of.v), ("P", of.p), ("FLMASS", of.flmass)):
                assert cases.rel_l2(ctx.download(fld)[:ref.size], ref) < 1e-5, (step, fld)
    assert np.abs(of.u[:n]).max() > 0.05
    ctx.close()


def test_piso_rejects_bad_arguments(fc):
    mesh = MESHES["skew"]()
    ctx = fc.Context(0)
    ctx.set_mesh(mesh)
    ctx.create_csr(download=False)
    with pytest.raises(fc.FcError) as e:
        ctx.piso(fc.piso_opts(pRefCell=mesh.numCells + 1))
    assert e.value.code == fc.FC_ERR_ARG
    ctx.close()
