"""CPU checks of the oracle's momentum predictor (calcuvw).  The reference stores no
outputs of calcuvw, so the restatement is pinned by properties the discretisation must have."""
import numpy as np
import pytest

from freecappuccino_b200 import cases
from oracle import oracle


def setup(mesh, f, viscos=0.01):
    csr = oracle.create_csr(mesh)
    of = oracle.Fields(mesh, csr.nnz)
    for k in ("u", "v", "w", "p", "den"):
        getattr(of, k)[:] = f[k]
    x = oracle.UvwFields(mesh, of, viscos)
    return csr, of, x


def face_mass_fluxes(mesh, of):
    """flmass = den_f (U_f . S) with linear interpolation -- a consistent starting flux field."""
    F = mesh.numInnerFaces
    o, nb = mesh.owner[:F] - 1, mesh.neighbour - 1
    fx = mesh.facint
    uf = of.u[o] * (1 - fx) + of.u[nb] * fx
    vf = of.v[o] * (1 - fx) + of.v[nb] * fx
    wf = of.w[o] * (1 - fx) + of.w[nb] * fx
    df = of.den[o] * (1 - fx) + of.den[nb] * fx
    return df * (uf * mesh.arx[:F] + vf * mesh.ary[:F] + wf * mesh.arz[:F])


def boundary_fluxes(mesh, of, kind):
    fs, sl = mesh.boundary_faces(kind), mesh.boundary_slots(kind)
    return of.den[sl] * (of.u[sl] * mesh.arx[fs] + of.v[sl] * mesh.ary[fs] + of.w[sl] * mesh.arz[fs])


def test_uniform_flow_is_a_fixed_point():
    """Uniform velocity, uniform pressure, inflow on the low faces and outflow on the high ones: every explicit
    term vanishes and the assembled system is satisfied by the field itself (res0 ~ round-off), for every
    convection scheme."""
    mesh = cases.hex_case(7, 6, 5, kinds=("inlet", "outlet", "inlet", "outlet", "inlet", "outlet"))
    nt = mesh.numTotal
    f = dict(u=np.full(nt, 1.0), v=np.full(nt, 0.5), w=np.full(nt, 0.25), p=np.full(nt, 3.0), den=np.ones(nt))
    for scheme in oracle.SCHEMES:
        if scheme in ("smart", "avl-smart", "muscl", "umist", "koren", "charm", "ospre", "linear"):
            continue  # r = 0/0 on a uniform field (the reference divides by u_N - u_P unguarded)
        csr, of, x = setup(mesh, f)
        of.flmass[:] = face_mass_fluxes(mesh, of)
        of.fmi[:] = boundary_fluxes(mesh, of, "inlet")
        of.fmo[:] = boundary_fluxes(mesh, of, "outlet")
        rep = oracle.calcuvw(mesh, csr, of, x, oracle.uvw_opts(scheme=scheme, sor=(1e-12,) * 3, nsw=(50,) * 3))
        n = mesh.numCells
        scale = np.abs(of.a).max()
        for k in range(3):
            assert rep.rep[k].res0 < 1e-10 * scale * n, (scheme, k, rep.rep[k].res0)
        assert np.allclose(of.u[:n], 1.0, atol=1e-10) and np.allclose(of.v[:n], 0.5, atol=1e-10)
        assert np.allclose(of.w[:n], 0.25, atol=1e-10)


def test_matrix_is_diagonally_dominant_and_ap_is_its_inverse_diagonal():
    mesh = cases.skew_case()
    f = cases.flow_fields(mesh)
    csr, of, x = setup(mesh, f)
    of.flmass[:] = face_mass_fluxes(mesh, of)
    of.fmi[:mesh.count("inlet")] = boundary_fluxes(mesh, of, "inlet")
    of.fmo[:mesh.count("outlet")] = boundary_fluxes(mesh, of, "outlet")
    opts = oracle.uvw_opts(urf=(0.7, 0.8, 0.9), bdf=True, timestep=1e-2)
    oracle.calcuvw_assemble(mesh, csr, of, x, opts)
    n = mesh.numCells
    for comp, ap in enumerate((x.apu, x.apv, x.apw)):
        oracle.calcuvw_component(mesh, csr, of, x, opts, comp)
        d = of.a[csr.diag - 1]
        off = np.add.reduceat(np.abs(of.a), csr.ioffset[:-1] - 1) - np.abs(d)
        assert np.all(of.a[np.setdiff1d(np.arange(csr.nnz), csr.diag - 1)] <= 0.0)
        assert np.all(d > off)      # strict: under-relaxation + the unsteady term
        assert np.allclose(ap[:n], 1.0 / (d + oracle.SMALL), rtol=1e-15)


def test_second_order_schemes_agree_on_a_linear_field():
    """Linear velocity on an orthogonal mesh: the Gauss gradient is exact, so central-f, linear-f, muscl-f and
    cds-corrected reconstruct the same face value and give the same explicit source."""
    mesh = cases.hex_case(6, 5, 4, kinds=("inlet", "outlet", "wall", "wall", "symmetry", "symmetry"))
    nt, n = mesh.numTotal, mesh.numCells
    x_, y_, z_ = np.zeros(nt), np.zeros(nt), np.zeros(nt)
    x_[:n], y_[:n], z_[:n] = mesh.xc[:n], mesh.yc[:n], mesh.zc[:n]
    from freecappuccino_b200 import mesh as M
    for kind in M.KINDS:
        fs, sl = mesh.boundary_faces(kind), mesh.boundary_slots(kind)
        x_[sl], y_[sl], z_[sl] = mesh.xf[fs], mesh.yf[fs], mesh.zf[fs]
    f = dict(u=1.0 + 0.3 * x_ + 0.2 * y_ - 0.1 * z_, v=0.2 - 0.1 * x_ + 0.05 * y_ + 0.3 * z_,
             w=-0.1 + 0.2 * x_ + 0.1 * y_ - 0.35 * z_, p=1.0 - 0.5 * x_, den=np.ones(nt))
    su = {}
    for scheme in ("central-f", "linear-f", "muscl-f", "cds-corrected"):
        csr, of, x = setup(mesh, f)
        of.flmass[:] = face_mass_fluxes(mesh, of)
        of.fmi[:mesh.count("inlet")] = boundary_fluxes(mesh, of, "inlet")
        of.fmo[:mesh.count("outlet")] = boundary_fluxes(mesh, of, "outlet")
        oracle.calcuvw_assemble(mesh, csr, of, x, oracle.uvw_opts(scheme=scheme))
        su[scheme] = np.concatenate([of.su, x.sv, x.sw])
    ref = su["central-f"]
    for scheme, s in su.items():
        assert np.allclose(s, ref, rtol=0, atol=1e-12 * np.abs(ref).max()), scheme


def test_pressure_source_is_minus_the_surface_integral_of_p():
    """calcPressDiv on an orthogonal mesh (the skewness term vanishes): su = -sum_f p_f S_x with linear face
    interpolation, boundary faces taking the boundary slot's pressure."""
    mesh = cases.hex_case(5, 4, 3, kinds=("inlet", "outlet", "wall", "wall", "symmetry", "prOutlet"))
    f = cases.flow_fields(mesh)
    f["u"][:] = 0.0; f["v"][:] = 0.0; f["w"][:] = 0.0
    csr, of, x = setup(mesh, f, viscos=0.0)
    oracle.calcuvw_assemble(mesh, csr, of, x, oracle.uvw_opts(viscos=0.0))
    n, F = mesh.numCells, mesh.numInnerFaces
    p = of.p   # boundary slots were overwritten by bpres (fieldManipulation.f90:82-87)
    o, nb = mesh.owner[:F] - 1, mesh.neighbour - 1
    pf = p[o] * (1 - mesh.facint) + p[nb] * mesh.facint
    ex = np.zeros(n)
    np.add.at(ex, o, -pf * mesh.arx[:F])
    np.add.at(ex, nb, pf * mesh.arx[:F])
    from freecappuccino_b200 import mesh as M
    for kind in M.KINDS:
        fs, sl = mesh.boundary_faces(kind), mesh.boundary_slots(kind)
        np.add.at(ex, mesh.owner[fs] - 1, -p[sl] * mesh.arx[fs])
    assert np.allclose(of.su, ex, rtol=0, atol=1e-13 * np.abs(ex).max())


def test_press_div_keeps_the_reference_gradient_addressing():
    """fieldManipulation.f90:433-435 reads df(ijp,k) from a (3,numCells) array: flat elements ijp-1, ijp+2,
    ijp+5 (0-based).  On a skewed mesh the result differs from the 'intended' gradient; the oracle must
    reproduce the addressing the compiled reference uses."""
    mesh = cases.skew_case()
    f = cases.flow_fields(mesh)
    f["u"][:] = 0.0; f["v"][:] = 0.0; f["w"][:] = 0.0
    csr, of, x = setup(mesh, f, viscos=0.0)
    oracle.calcuvw_assemble(mesh, csr, of, x, oracle.uvw_opts(viscos=0.0))
    n, F = mesh.numCells, mesh.numInnerFaces
    p, flat = of.p, of.dPdxi.ravel()
    o, nb = mesh.owner[:F] - 1, mesh.neighbour - 1
    fxn = mesh.facint; fxp = 1.0 - fxn
    xi = mesh.xc[o] * fxp + mesh.xc[nb] * fxn
    yi = mesh.yc[o] * fxp + mesh.yc[nb] * fxn
    zi = mesh.zc[o] * fxp + mesh.zc[nb] * fxn
    d = [flat[o + 3 * k] * fxp + flat[nb + 3 * k] * fxn for k in range(3)]
    fie = p[o] * fxp + p[nb] * fxn + d[0] * (mesh.xf[:F] - xi) + d[1] * (mesh.yf[:F] - yi) + d[2] * (mesh.zf[:F] - zi)
    ex = np.zeros(n)
    np.add.at(ex, o, -fie * mesh.arx[:F])
    np.add.at(ex, nb, fie * mesh.arx[:F])
    from freecappuccino_b200 import mesh as M
    for kind in M.KINDS:
        fs, sl = mesh.boundary_faces(kind), mesh.boundary_slots(kind)
        np.add.at(ex, mesh.owner[fs] - 1, -p[sl] * mesh.arx[fs])
    assert np.allclose(of.su, ex, rtol=0, atol=1e-12 * np.abs(ex).max())
    # and it is NOT what the intended (k, ijp) addressing would give
    g = of.dPdxi
    d2 = [g[o, k] * fxp + g[nb, k] * fxn for k in range(3)]
    fie2 = p[o] * fxp + p[nb] * fxn + d2[0] * (mesh.xf[:F] - xi) + d2[1] * (mesh.yf[:F] - yi) + d2[2] * (mesh.zf[:F] - zi)
    assert np.abs(fie2 - fie).max() > 1e-6


@pytest.mark.parametrize("scheme", list(oracle.SCHEMES))
def test_predictor_reduces_the_momentum_residual(scheme):
    mesh = cases.skew_case(7, 6, 5)
    f = cases.channel_fields(mesh)
    csr, of, x = setup(mesh, f)
    of.flmass[:] = face_mass_fluxes(mesh, of)
    of.fmi[:mesh.count("inlet")] = boundary_fluxes(mesh, of, "inlet")
    of.fmo[:mesh.count("outlet")] = boundary_fluxes(mesh, of, "outlet")
    rep = oracle.calcuvw(mesh, csr, of, x, oracle.uvw_opts(scheme=scheme, sor=(1e-6,) * 3, nsw=(100,) * 3,
                                                          bdf=True, timestep=0.05))
    for k in range(3):
        assert 0 < rep.rep[k].iters < 100
        assert rep.rep[k].resl < 1e-6 * rep.rep[k].res0 * 1.0001
        assert np.isfinite(rep.rep[k].resl)


# ---- src-parallel semantics (fc_oracle_par_uvw.c) ----
def par_setup(mesh, f, nranks, flmass, x_ser):
    from freecappuccino_b200 import mesh as M
    from oracle import oracle_par as OP
    parts = M.partition(mesh, M.slab_ranks(mesh.numCells, nranks), nranks)
    pc = OP.ParCase(parts)
    xs = pc.uvw_fields(0.0)
    for r, part in enumerate(parts):
        fr = pc.fields[r]
        for k in ("u", "v", "w", "p", "den"):
            getattr(fr, k)[:] = M.scatter_total(mesh, part, f[k])
        gf = part.face_global
        F = part.numInnerFaces
        fr.flmass[:] = flmass[gf[:F]]
        if part.npro:
            pf = gf[part.iProcFacesStart:part.iProcFacesStart + part.npro]
            dot = (part.arx[part.iProcFacesStart:] * mesh.arx[pf] + part.ary[part.iProcFacesStart:] * mesh.ary[pf]
                   + part.arz[part.iProcFacesStart:] * mesh.arz[pf])
            pc.fmpro[r][:part.npro] = flmass[pf] * np.sign(dot)
        for kind, dst, src in (("inlet", fr.fmi, None), ("outlet", fr.fmo, None)):
            c = part.count(kind)
            if c:
                g0 = gf[part.faces_start(kind):part.faces_start(kind) + c] - mesh.faces_start(kind)
                dst[:c] = (x_ser["fmi"] if kind == "inlet" else x_ser["fmo"])[g0]
        for k in ("vis", "uo", "vo", "wo", "uoo", "voo", "woo", "t"):
            getattr(xs[r], k)[:] = M.scatter_total(mesh, part, x_ser[k])
    return parts, pc, xs


@pytest.mark.parametrize("nranks", [1, 2, 3])
def test_parallel_predictor_agrees_with_the_serial_one(nranks):
    """src-parallel/calcuvw.f90 on 1-3 ranks (lock-step oracle) against the serial routine: the processor faces use
    the same facefluxuvw, so the global system is the serial one up to rounding; with tight BiCGStab solves the
    fields agree to the solver tolerance.  The pressure is uniform here: on a skewed mesh the reference's
    df(ijp,k) addressing in presFaceDivInner reads other cells' gradient components by LOCAL flat index, so with a
    non-trivial pressure the parallel and the serial build legitimately differ."""
    from freecappuccino_b200 import mesh as M
    mesh = cases.skew_case(6, 5, 9)
    f = cases.channel_fields(mesh)
    f["p"][:] = 2.5
    rng = np.random.default_rng(4)
    csr, of, x = setup(mesh, f)
    nt = mesh.numTotal
    of.flmass[:] = face_mass_fluxes(mesh, of)
    of.fmi[:mesh.count("inlet")] = boundary_fluxes(mesh, of, "inlet")
    of.fmo[:mesh.count("outlet")] = boundary_fluxes(mesh, of, "outlet")
    x.vis[:] = 0.01 * (1.0 + 0.3 * rng.random(nt))
    for k in ("uo", "vo", "wo", "uoo", "voo", "woo", "t"):
        getattr(x, k)[:] = 0.1 * rng.standard_normal(nt)
    xser = dict(vis=x.vis.copy(), uo=x.uo.copy(), vo=x.vo.copy(), wo=x.wo.copy(), uoo=x.uoo.copy(), voo=x.voo.copy(),
                woo=x.woo.copy(), t=x.t.copy(), fmi=of.fmi.copy(), fmo=of.fmo.copy())
    # no Crank-Nicolson here: src-parallel halves apr AFTER the processor faces' full `can` went into sp*
    # (calcuvw.f90:240-244 vs :427), so its cn diagonal differs from the serial one by construction
    kw = dict(scheme="muscl-f", urf=(0.7, 0.8, 0.6), sor=(1e-11,) * 3, nsw=(300,) * 3, bdf=True, btime=1.0, timestep=0.05)
    parts, pc, xs = par_setup(mesh, f, nranks, of.flmass.copy(), xser)
    rs = oracle.calcuvw(mesh, csr, of, x, oracle.uvw_opts(**kw))
    rp = pc.calcuvw(oracle.uvw_opts(**kw))
    n = mesh.numCells
    for k in range(3):
        assert rp.rep[k].iters < 300 and rs.rep[k].iters < 300
        assert rp.rep[k].res0 == pytest.approx(rs.rep[k].res0, rel=1e-9)
    for name, ser in (("u", of.u), ("v", of.v), ("w", of.w), ("apu", x.apu)):
        got = M.gather_cells(mesh, parts, [getattr(fr, name) for fr in pc.fields])
        assert np.allclose(got, ser[:n], rtol=1e-8, atol=1e-10), name
    # halo of u and apu is current after the final exchanges
    for r, part in enumerate(parts):
        if part.npro:
            assert np.array_equal(pc.fields[r].u[part.numCells:part.numCells + part.npro],
                                  M.gather_cells(mesh, parts, [fr.u for fr in pc.fields])[part.halo_global])
            assert np.array_equal(pc.fields[r].apu[part.numCells:], M.gather_cells(mesh, parts, [fr.apu for fr in pc.fields])[part.halo_global])
