"""Seeded test cases shared by the CPU and GPU suites (meshes + fields of the pressure path)."""
from __future__ import annotations

import numpy as np

from freecappuccino_b200 import mesh as M


def hex_case(nx=12, ny=10, nz=8, kinds=("wall",) * 6, lengths=(1.0, 1.0, 1.0)):
    return M.hex_mesh(nx, ny, nz, lengths, kinds)


def skew_case(nx=10, ny=9, nz=8, kinds=("inlet", "outlet", "wall", "wall", "symmetry", "symmetry"), jitter=0.2,
              seed=12345, lengths=(1.0, 0.8, 0.6)):
    """Non-orthogonal hex mesh (jittered points) with every boundary kind of the path."""
    pts, faces, owner, neigh, counts, starts = M.hex_polymesh_arrays(nx, ny, nz, lengths, kinds, jitter, seed)
    return M.geometry_from_polymesh(pts, faces, owner, neigh, counts, starts)


def poly_case(N=6, jitter=0.15):
    """Non-orthogonal polyhedral mesh of BASELINE config 5 at test size (2 N^3 truncated octahedra)."""
    return M.bcc_poly_mesh(N, jitter)


def flow_fields(mesh, seed=7):
    """Deterministic smooth velocity / pressure + momentum-diagonal fields (config 4 of SURVEY 8d,
    plus a small seeded perturbation so that no two values coincide)."""
    rng = np.random.default_rng(seed)
    n, nt = mesh.numCells, mesh.numTotal
    two_pi = 2.0 * np.pi
    x = np.zeros(nt); y = np.zeros(nt); z = np.zeros(nt)
    x[:n], y[:n], z[:n] = mesh.xc[:n], mesh.yc[:n], mesh.zc[:n]
    # boundary slots take the face-centre coordinates
    for kind in M.KINDS:
        fs, sl = mesh.boundary_faces(kind), mesh.boundary_slots(kind)
        x[sl], y[sl], z[sl] = mesh.xf[fs], mesh.yf[fs], mesh.zf[fs]
    u = np.sin(two_pi * x) * np.cos(two_pi * y) * np.cos(two_pi * z) + 1e-3 * rng.standard_normal(nt)
    v = -np.cos(two_pi * x) * np.sin(two_pi * y) * np.cos(two_pi * z) + 1e-3 * rng.standard_normal(nt)
    w = 0.1 * np.sin(two_pi * x) * np.sin(two_pi * y) * np.sin(two_pi * z) + 1e-3 * rng.standard_normal(nt)
    p = np.cos(two_pi * x) * np.cos(two_pi * y) + 1e-3 * rng.standard_normal(nt)
    for kind in ("wall",):
        sl = mesh.boundary_slots(kind)
        u[sl] = v[sl] = w[sl] = 0.0
    den = 1.0 + 0.05 * np.cos(two_pi * (x + y))
    h = float(np.mean(mesh.vol[:n])) ** (1.0 / 3.0)
    npc = mesh.numCells + mesh.npro
    xa, ya, za = mesh.xc[:npc], mesh.yc[:npc], mesh.zc[:npc]
    base = 1.0 / (6.0 * h * (1.0 + 0.1 * np.sin(two_pi * (xa + ya + za))))
    apu, apv, apw = base.copy(), base * 1.03, base * 0.97
    return dict(u=u, v=v, w=w, p=p, den=den, apu=apu, apv=apv, apw=apw)


def inlet_fluxes(mesh, fields):
    """fmi = den * (U.S) at inlet faces (bcin.f90), flomas = -sum(fmi) (inflow is negative)."""
    fs, sl = mesh.boundary_faces("inlet"), mesh.boundary_slots("inlet")
    fmi = fields["den"][sl] * (fields["u"][sl] * mesh.arx[fs] + fields["v"][sl] * mesh.ary[fs]
                               + fields["w"][sl] * mesh.arz[fs])
    flomas = float(-np.sum(fmi)) if fmi.size else 0.0
    return fmi, flomas


def poisson_rhs(mesh):
    """su = 8 pi^2 sin(2 pi x) sin(2 pi y) vol (poisson.f90:63)."""
    n = mesh.numCells
    return 8.0 * np.pi ** 2 * np.sin(2 * np.pi * mesh.xc[:n]) * np.sin(2 * np.pi * mesh.yc[:n]) * mesh.vol[:n]


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def config4_fields(mesh):
    """Synthetic pressure-correction case of SURVEY 8(d) config 4 (deterministic, no RNG):
    u = sin cos cos, v = -cos sin cos, w = 0.1 sin sin sin (not divergence free, so su is non-trivial),
    p = cos cos, den = 1, ap* = 1/(6 h (1 + 0.1 sin 2pi(x+y+z))), no-slip walls.  Works on a
    partitioned mesh too: halo / boundary slots are filled from their own coordinates."""
    n, nt, npc = mesh.numCells, mesh.numTotal, mesh.numCells + mesh.npro
    two_pi = 2.0 * np.pi
    x = np.zeros(nt); y = np.zeros(nt); z = np.zeros(nt)
    x[:npc], y[:npc], z[:npc] = mesh.xc[:npc], mesh.yc[:npc], mesh.zc[:npc]
    u = np.sin(two_pi * x) * np.cos(two_pi * y) * np.cos(two_pi * z)
    v = -np.cos(two_pi * x) * np.sin(two_pi * y) * np.cos(two_pi * z)
    w = 0.1 * np.sin(two_pi * x) * np.sin(two_pi * y) * np.sin(two_pi * z)
    p = np.cos(two_pi * x) * np.cos(two_pi * y)
    u[npc:] = v[npc:] = w[npc:] = 0.0
    p[npc:] = 0.0
    den = np.ones(nt)
    h = 1.0 / round((mesh.gloCells or n) ** (1.0 / 3.0))
    ap = 1.0 / (6.0 * h * (1.0 + 0.1 * np.sin(two_pi * (x[:npc] + y[:npc] + z[:npc]))))
    return dict(u=u, v=v, w=w, p=p, den=den, apu=ap.copy(), apv=ap.copy(), apw=ap.copy())


def golden_mesh(npz_path, prefix=""):
    """Mesh of a stored polyMesh fixture (tests/golden/*.npz, written by make_fixtures.py)."""
    d = np.load(npz_path)
    g = lambda k: d[prefix + k]
    counts, starts = {}, {}
    for kind, nf, st in zip(g("bkind"), g("bn"), g("bstart")):
        kind = {"wallIsoth": "wall", "wallAdiab": "wall", "wallQFlux": "wall"}.get(str(kind), str(kind))
        if kind not in counts:
            counts[kind] = 0
            starts[kind] = int(st)
        counts[kind] += int(nf)
    return M.geometry_from_polymesh(g("points"), g("faces").astype(np.int64), g("owner").astype(np.int32) + 1,
                                    g("neighbour").astype(np.int32) + 1, counts, starts)


def channel_fields(mesh, seed=11):
    """Through-flow in +x (inlet left, outlet right, as in pitzDaily) with smooth perturbations: the outlet
    mass-flow scaling of adjustMassFlow / correctBoundaryConditionsVelocity is then well conditioned."""
    f = flow_fields(mesh, seed)
    n, nt = mesh.numCells, mesh.numTotal
    x = np.zeros(nt); y = np.zeros(nt)
    x[:n], y[:n] = mesh.xc[:n], mesh.yc[:n]
    for kind in M.KINDS:
        fs, sl = mesh.boundary_faces(kind), mesh.boundary_slots(kind)
        x[sl], y[sl] = mesh.xf[fs], mesh.yf[fs]
    lx = max(float(np.ptp(mesh.xf)), 1e-30)
    f["u"] = 1.0 + 0.1 * np.sin(2 * np.pi * x / lx) * np.cos(2 * np.pi * y / lx) + 0.05 * f["u"]
    f["v"] = 0.05 * f["v"]
    f["w"] = 0.05 * f["w"]
    sl = mesh.boundary_slots("wall")
    f["u"][sl] = f["v"][sl] = f["w"][sl] = 0.0
    return f
