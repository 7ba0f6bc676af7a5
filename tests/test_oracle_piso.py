"""CPU checks of the oracle's PISO / PIMPLE pressure equation.  The reference stores no
outputs of these routines; the restatement is pinned by properties."""
import numpy as np
import pytest

from freecappuccino_b200 import cases
from oracle import oracle


def momentum_state(mesh, f, timestep, scheme="muscl-f", seed=5):
    """Fields after one calcuvw (tight solves): the state PISO starts from (a = W momentum matrix, ap*, u, v, w)."""
    rng = np.random.default_rng(seed)
    csr = oracle.create_csr(mesh)
    of = oracle.Fields(mesh, csr.nnz)
    for k in ("u", "v", "w", "p", "den"):
        getattr(of, k)[:] = f[k]
    F = mesh.numInnerFaces
    o, nb = mesh.owner[:F] - 1, mesh.neighbour - 1
    fx = mesh.facint
    of.flmass[:] = (of.den[o] * (1 - fx) + of.den[nb] * fx) * (
        (of.u[o] * (1 - fx) + of.u[nb] * fx) * mesh.arx[:F] + (of.v[o] * (1 - fx) + of.v[nb] * fx) * mesh.ary[:F]
        + (of.w[o] * (1 - fx) + of.w[nb] * fx) * mesh.arz[:F])
    fmi, flomas = cases.inlet_fluxes(mesh, f)
    of.fmi[:fmi.size] = fmi
    x = oracle.UvwFields(mesh, of, 0.01)
    x.uo[:], x.vo[:], x.wo[:] = of.u, of.v, of.w
    x.uoo[:], x.voo[:], x.woo[:] = of.u, of.v, of.w
    x.t[:] = rng.standard_normal(mesh.numTotal)
    oo = oracle.uvw_opts(scheme=scheme, urf=(1.0, 1.0, 1.0), sor=(1e-10,) * 3, nsw=(100,) * 3, bdf=True,
                         timestep=timestep)
    oracle.calcuvw(mesh, csr, of, x, oo)
    return csr, of, x, flomas


def true_divergence(mesh, of):
    n, F = mesh.numCells, mesh.numInnerFaces
    o, nb = mesh.owner[:F] - 1, mesh.neighbour - 1
    div = np.zeros(n)
    np.add.at(div, o, of.flmass)
    np.add.at(div, nb, -of.flmass)
    for kind, fm in (("inlet", of.fmi), ("outlet", of.fmo)):
        fs = mesh.boundary_faces(kind)
        np.add.at(div, mesh.owner[fs] - 1, fm[:len(fs)])
    return div


def test_hbya_matches_an_independent_numpy_evaluation():
    mesh = cases.skew_case(6, 5, 4)
    csr, of, x, _ = momentum_state(mesh, cases.channel_fields(mesh), timestep=0.01)
    n = mesh.numCells
    h = of.a.copy()
    u0, v0 = of.u.copy(), of.v.copy()
    po = oracle.piso_opts(bdf=True, btime=1.0, timestep=0.01, lbuoy=True, beta=0.2, tref=0.1, densit=1.05,
                          grav=(0.0, -9.81, 0.3))
    oracle.get_rAU_x_UEqnH(mesh, csr, of, x, po, h)
    # numpy: H(u) = sources - sum_nb h(P,nb) u_nb, off-diagonals only
    import scipy.sparse as sp
    H = sp.csr_matrix((h, csr.ja - 1, csr.ioffset - 1), shape=(n, n))
    H = H - sp.diags(H.diagonal())
    apo = of.den[:n] * mesh.vol[:n] / 0.01
    heat = 0.2 * 1.05 * (x.t[:n] - 0.1) * mesh.vol[:n]
    su = apo * (2.0 * x.uo[:n]) - apo * (0.5 * x.uoo[:n]) - 0.0 * heat - H @ u0[:n]
    sv = apo * (2.0 * x.vo[:n]) - apo * (0.5 * x.voo[:n]) + 9.81 * heat - H @ v0[:n]
    assert np.allclose(of.u[:n], x.apu[:n] * su, rtol=1e-11, atol=1e-13)
    assert np.allclose(of.v[:n], x.apv[:n] * sv, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("pimple", [False, True])
def test_corrected_fluxes_satisfy_continuity_away_from_the_reference_cell(pimple):
    """After the pressure solve (tight tolerance) the corrected face fluxes are divergence free, except in the
    cells that touch pRefCell's row: the reference clears that row (PISO :188-189), so its own faces get no
    correction (:262 reads a(icell_jcell) of the cleared row)."""
    mesh = cases.hex_case(8, 7, 6, kinds=("inlet", "outlet", "wall", "wall", "symmetry", "symmetry"))
    csr, of, x, flomas = momentum_state(mesh, cases.channel_fields(mesh), timestep=0.01)
    pref = 40
    of.pp[:] = of.p     # PIMPLE starts pp from the current pressure
    po = oracle.piso_opts(ncorr=1, npcor=1, pRefCell=pref, pimple=pimple, urf_p=0.8, flomas=flomas, bdf=True,
                          timestep=0.01, sor=1e-12, nsw=500)
    rep = oracle.piso(mesh, csr, of, x, po)
    assert rep.nsolves == 1 and 0 < rep.rep[0].iters < 500
    div = true_divergence(mesh, of)
    row = csr.ja[csr.ioffset[pref - 1] - 1:csr.ioffset[pref] - 1] - 1     # pRefCell and its neighbours
    mask = np.ones(mesh.numCells, bool)
    mask[row] = False
    # outlet cells: correctBoundaryConditionsVelocity re-extrapolates and rescales fmo afterwards (global balance only)
    mask[mesh.owner[mesh.boundary_faces("outlet")] - 1] = False
    scale = np.abs(of.flmass).max()
    assert np.abs(div[mask]).max() < 1e-9 * scale
    quirk = np.zeros(mesh.numCells, bool)
    quirk[row] = True
    assert np.abs(div[quirk]).max() > 1e-6 * scale      # the quirk is really there
    assert of.pp[pref - 1] == pytest.approx(of.p[pref - 1] if not pimple else of.pp[pref - 1])


def test_transient_cavity_with_piso_develops_and_stays_bounded():
    """Lid-driven cavity started from rest, 15 time steps of calcuvw + PISO (2 correctors): the vortex spins up,
    the pressure solver converges every time and nothing blows up."""
    mesh = cases.golden_mesh(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "cavity.npz"))
    nt, n = mesh.numTotal, mesh.numCells
    csr = oracle.create_csr(mesh)
    of = oracle.Fields(mesh, csr.nnz)
    of.u[mesh.boundary_slots("wall")[:20]] = 1.0
    x = oracle.UvwFields(mesh, of, 0.01)
    dt = 5e-3
    uo = oracle.uvw_opts(scheme="muscl-f", urf=(1.0,) * 3, sor=(1e-6,) * 3, nsw=(50,) * 3, bdf=True, btime=0.0,
                         timestep=dt, viscos=0.01)
    po = oracle.piso_opts(ncorr=2, npcor=1, pRefCell=1, const_mflux=True, bdf=True, btime=0.0, timestep=dt, sor=1e-6,
                          nsw=300)
    ke = []
    for step in range(15):
        x.uoo[:], x.voo[:], x.woo[:] = x.uo, x.vo, x.wo
        x.uo[:], x.vo[:], x.wo[:] = of.u, of.v, of.w
        oracle.calcuvw(mesh, csr, of, x, uo)
        rep = oracle.piso(mesh, csr, of, x, po)
        assert rep.nsolves == 2
        assert all(rep.rep[k].iters < 300 for k in range(2))
        ke.append(float(np.sum(mesh.vol[:n] * (of.u[:n] ** 2 + of.v[:n] ** 2))))
        assert np.isfinite(ke[-1])
    assert ke[-1] > ke[0] > 0 and np.abs(of.u[:n]).max() < 1.5
    assert all(b > a for a, b in zip(ke, ke[1:]))      # monotone spin-up from rest


# ---- several ranks: src-parallel/PISO_multiple_correction.f90 in lock step (fc_oracle_par_piso.c) ----
def _par_state(mesh, nranks, timestep=0.01):
    """R-rank state after one lock-step calcuvw: the matrices / apr / ap* PISO starts from."""
    from freecappuccino_b200 import mesh as M
    from oracle import oracle_par
    f = cases.channel_fields(mesh)
    rng = np.random.default_rng(5)
    t = rng.standard_normal(mesh.numTotal)
    F = mesh.numInnerFaces
    o, nb = mesh.owner[:F] - 1, mesh.neighbour - 1
    fx = mesh.facint
    fl_g = (f["den"][o] * (1 - fx) + f["den"][nb] * fx) * (
        (f["u"][o] * (1 - fx) + f["u"][nb] * fx) * mesh.arx[:F] + (f["v"][o] * (1 - fx) + f["v"][nb] * fx) * mesh.ary[:F]
        + (f["w"][o] * (1 - fx) + f["w"][nb] * fx) * mesh.arz[:F])
    fmi_g, flomas = cases.inlet_fluxes(mesh, f)
    parts = M.partition(mesh, M.slab_ranks(mesh.numCells, nranks), nranks) if nranks > 1 else [mesh]
    pc = oracle_par.ParCase(parts)
    xs = pc.uvw_fields(0.01)
    for r, part in enumerate(parts):
        fr = pc.fields[r]
        if nranks > 1:
            for k in ("u", "v", "w", "p", "den"):
                getattr(fr, k)[:] = M.scatter_total(mesh, part, f[k])
            xs[r].t[:] = M.scatter_total(mesh, part, t)
            gf = part.face_global
            fr.flmass[:] = fl_g[gf[:part.numInnerFaces]]
            pf = gf[part.iProcFacesStart:part.iProcFacesStart + part.npro]
            sgn = np.sign(part.arx[part.iProcFacesStart:] * mesh.arx[pf] + part.ary[part.iProcFacesStart:] * mesh.ary[pf]
                          + part.arz[part.iProcFacesStart:] * mesh.arz[pf])
            pc.fmpro[r][:part.npro] = fl_g[pf] * sgn
            c = part.count("inlet")
            gl = gf[part.faces_start("inlet"):part.faces_start("inlet") + c] - mesh.faces_start("inlet")
            fr.fmi[:c] = fmi_g[gl]
        else:
            for k in ("u", "v", "w", "p", "den"):
                getattr(fr, k)[:] = f[k]
            xs[r].t[:] = t
            fr.flmass[:] = fl_g
            fr.fmi[:fmi_g.size] = fmi_g
        for a, b in (("uo", "u"), ("vo", "v"), ("wo", "w"), ("uoo", "u"), ("voo", "v"), ("woo", "w")):
            getattr(xs[r], a)[:] = getattr(fr, b)
    oo = oracle.uvw_opts(scheme="muscl-f", urf=(1.0, 1.0, 1.0), sor=(1e-10,) * 3, nsw=(100,) * 3, bdf=True, timestep=timestep)
    pc.calcuvw(oo)
    return pc, parts, flomas


def test_one_rank_of_the_parallel_piso_is_the_serial_piso_up_to_the_parallel_small():
    """With one rank there are no processor faces: the lock-step routine differs from the serial one only by the
    `+small` of the parallel preconditioner (and by reporting the continuity errors once), so the fields agree to
    round-off."""
    mesh = cases.hex_case(8, 7, 6, kinds=("inlet", "outlet", "wall", "wall", "symmetry", "symmetry"))
    csr, of, x, flomas = momentum_state(mesh, cases.channel_fields(mesh), timestep=0.01)
    pc, parts, flomas_p = _par_state(mesh, 1)
    assert flomas_p == flomas
    for k in ("u", "v", "w", "a"):
        assert np.allclose(getattr(pc.fields[0], k), getattr(of, k), rtol=1e-9, atol=1e-12), k
    po = oracle.piso_opts(ncorr=2, npcor=1, pRefCell=40, flomas=flomas, bdf=True, timestep=0.01, sor=1e-12, nsw=500)
    rep_s = oracle.piso(mesh, csr, of, x, po)
    rep_p = pc.piso(po)
    assert rep_p.nsolves == rep_s.nsolves == 2
    for k in ("u", "v", "w", "p", "flmass"):
        assert np.allclose(getattr(pc.fields[0], k), getattr(of, k), rtol=1e-7, atol=1e-10), k


@pytest.mark.parametrize("nranks", [2, 3])
@pytest.mark.parametrize("pimple", [False, True])
def test_parallel_piso_fluxes_satisfy_continuity(nranks, pimple):
    """Several ranks: after a tight pressure solve the corrected fluxes (inner faces + processor faces) are divergence
    free on every rank away from pRefCell's row (rank 0) and the outlet cells."""
    mesh = cases.hex_case(8, 7, 6 * nranks // 2 * 2, kinds=("inlet", "outlet", "wall", "wall", "symmetry", "symmetry"))
    pc, parts, flomas = _par_state(mesh, nranks)
    pref = 40
    for fr in pc.fields:
        fr.pp[:] = fr.p
    po = oracle.piso_opts(ncorr=1, npcor=1, pRefCell=pref, pimple=pimple, urf_p=0.8, flomas=flomas, bdf=True, timestep=0.01,
                          sor=1e-12, nsw=800)
    rep = pc.piso(po)
    assert rep.nsolves == 1 and 0 < rep.rep[0].iters < 800
    for r, part in enumerate(parts):
        fr = pc.fields[r]
        n, F = part.numCells, part.numInnerFaces
        o, nb = part.owner[:F] - 1, part.neighbour - 1
        div = np.zeros(n)
        np.add.at(div, o, fr.flmass)
        np.add.at(div, nb, -fr.flmass)
        po_cells = part.owner[part.iProcFacesStart:part.iProcFacesStart + part.npro] - 1
        np.add.at(div, po_cells, pc.fmpro[r][:part.npro])
        for kind, fm in (("inlet", fr.fmi), ("outlet", fr.fmo)):
            fs = part.boundary_faces(kind)
            np.add.at(div, part.owner[fs] - 1, fm[:len(fs)])
        mask = np.ones(n, bool)
        if r == 0:
            csr = pc.csr[0]
            mask[csr.ja[csr.ioffset[pref - 1] - 1:csr.ioffset[pref] - 1] - 1] = False
        mask[part.owner[part.boundary_faces("outlet")] - 1] = False
        scale = np.abs(fr.flmass).max()
        assert np.abs(div[mask]).max() < 1e-8 * scale, (r, np.abs(div[mask]).max() / scale)
    # the two sides of a processor face carry opposite fluxes
    if nranks == 2:
        assert np.allclose(pc.fmpro[0][:parts[0].npro], -pc.fmpro[1][:parts[1].npro], rtol=1e-9, atol=1e-14)
