"""Mesh arrays in the layout of the reference's ``module geometry``.

This is the *input contract* of the pressure-correction path: the arrays the
Fortran host owns after ``mesh_geometry`` has run
(``src/mesh_geometry_and_topology.f90:13-98``; parallel additions
``src-parallel/mesh_geometry_and_topology.f90:497-515, 637-660, 879-881``).
Index arrays are 1-based int32, reals are float64, exactly as Fortran keeps
them, so they can be handed to the C-ABI unchanged.

Contents
  * :class:`Mesh`            -- the array bundle
  * :func:`hex_mesh`         -- synthetic structured-hex generator (configs 3/4)
  * :func:`bcc_poly_mesh`    -- synthetic non-orthogonal polyhedral generator (config 5)
  * :func:`geometry_from_polymesh` / :func:`read_polymesh` -- OpenFOAM polyMesh
    reader restating ``mesh_geometry`` (:859-1081) for the shipped cases
  * :func:`partition`        -- cell partitioner emitting per-rank meshes with
    processor faces + halo slots like an OpenFOAM ``decomposePar`` +
    the reference's ``process`` file
"""
from __future__ import annotations

import dataclasses
import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# boundary kinds in the order their slots follow the cells in every field array
# (mesh_geometry_and_topology.f90:597-609)
KINDS = ("inlet", "outlet", "symmetry", "wall", "prOutlet")


@dataclasses.dataclass
class Mesh:
    numCells: int
    numInnerFaces: int
    numFaces: int
    owner: np.ndarray       # int32 [numFaces], 1-based
    neighbour: np.ndarray   # int32 [numInnerFaces], 1-based
    xc: np.ndarray
    yc: np.ndarray
    zc: np.ndarray
    vol: np.ndarray         # [numCells (+npro in a partitioned mesh)]
    arx: np.ndarray
    ary: np.ndarray
    arz: np.ndarray
    xf: np.ndarray
    yf: np.ndarray
    zf: np.ndarray          # [numFaces]
    facint: np.ndarray      # [numInnerFaces]
    # boundary kinds: count and 0-based "FacesStart" (face = start + i, i = 1..count)
    counts: Dict[str, int] = dataclasses.field(default_factory=dict)
    starts: Dict[str, int] = dataclasses.field(default_factory=dict)
    # processor boundary (partitioned meshes only)
    npro: int = 0
    iProcFacesStart: int = 0
    fpro: Optional[np.ndarray] = None
    neighbProcNo: Optional[np.ndarray] = None      # [numConnections], 0-based ranks
    neighbProcOffset: Optional[np.ndarray] = None  # [numConnections+1], 1-based like the reference
    gloCells: int = 0
    cell_global: Optional[np.ndarray] = None       # local cell -> global cell (0-based), partitioned only
    halo_global: Optional[np.ndarray] = None       # processor face i -> global id of the remote cell
    face_global: Optional[np.ndarray] = None       # local face -> global face (0-based)
    # O-C cuts (none of the BASELINE configs has them; carried for the oracle)
    noc: int = 0
    iOCFacesStart: int = 0
    ijl: Optional[np.ndarray] = None
    ijr: Optional[np.ndarray] = None
    ijlFace: Optional[np.ndarray] = None
    foc: Optional[np.ndarray] = None

    # ---- derived sizes (mesh_geometry_and_topology.f90:580-609) ----
    @property
    def numBoundaryFaces(self) -> int:
        return self.numFaces - self.numInnerFaces

    @property
    def numTotal(self) -> int:
        return self.numCells + self.numBoundaryFaces

    @property
    def numPCells(self) -> int:
        return self.numCells + self.npro

    @property
    def nnz(self) -> int:
        return self.numCells + 2 * self.numInnerFaces

    def count(self, kind: str) -> int:
        return int(self.counts.get(kind, 0))

    def faces_start(self, kind: str) -> int:
        return int(self.starts.get(kind, 0))

    def slot_start(self, kind: str) -> int:
        """0-based start of the boundary slots of ``kind`` (``iWallStart`` etc.)."""
        s = self.numCells + self.npro
        for k in KINDS:
            if k == kind:
                return s
            s += self.count(k)
        if kind == "oc":
            return s
        raise KeyError(kind)

    def boundary_slots(self, kind: str) -> np.ndarray:
        """0-based indices into a numTotal-sized field for the slots of ``kind``."""
        s = self.slot_start(kind)
        return np.arange(s, s + self.count(kind))

    def boundary_faces(self, kind: str) -> np.ndarray:
        """0-based face indices of ``kind``."""
        s = self.faces_start(kind)
        return np.arange(s, s + self.count(kind))


# --------------------------------------------------------------------------
# synthetic structured hex mesh
# --------------------------------------------------------------------------
def _group_patches(patch_kinds: Sequence[str]) -> List[int]:
    """Order patches so faces of one kind are contiguous (the reference only
    records the first ``startFace`` of each kind, :414-470)."""
    seen: List[str] = []
    for k in patch_kinds:
        if k not in seen:
            seen.append(k)
    order: List[int] = []
    for k in seen:
        order += [i for i, pk in enumerate(patch_kinds) if pk == k]
    return order


def hex_mesh(nx: int, ny: int, nz: int, lengths: Tuple[float, float, float] = (1.0, 1.0, 1.0),
             kinds: Sequence[str] = ("wall",) * 6) -> Mesh:
    """Uniform hex mesh of a box, OpenFOAM (blockMesh) numbering.

    Cell id ``i + nx*(j + ny*k)`` (+1).  Inner faces are ordered by owner then
    neighbour ascending (upper-triangular order); boundary patches follow in the
    order x-, x+, y-, y+, z-, z+ (regrouped so equal kinds are contiguous),
    each ordered by owner.  ``kinds[p]`` in {inlet, outlet, symmetry, wall, prOutlet}.
    """
    lx, ly, lz = lengths
    dx, dy, dz = lx / nx, ly / ny, lz / nz
    n = nx * ny * nz
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    # flatten in cell-id order: i fastest
    order = np.argsort((ii + nx * (jj + ny * kk)).ravel(), kind="stable")
    ci = ii.ravel()[order].astype(np.int64)
    cj = jj.ravel()[order].astype(np.int64)
    ck = kk.ravel()[order].astype(np.int64)
    cid = np.arange(n, dtype=np.int64)
    xc = (ci + 0.5) * dx
    yc = (cj + 0.5) * dy
    zc = (ck + 0.5) * dz
    vol = np.full(n, dx * dy * dz)

    # inner faces: per cell, candidates (+x, +y, +z) in that order
    valid = np.stack([ci < nx - 1, cj < ny - 1, ck < nz - 1], axis=1)            # [n,3]
    nb = np.stack([cid + 1, cid + nx, cid + nx * ny], axis=1)
    own = np.repeat(cid[:, None], 3, axis=1)
    dirn = np.tile(np.arange(3), (n, 1))
    m = valid.ravel()
    f_own = own.ravel()[m]
    f_nb = nb.ravel()[m]
    f_dir = dirn.ravel()[m]
    nin = f_own.size
    area = np.array([dy * dz, dx * dz, dx * dy])
    half = np.array([0.5 * dx, 0.5 * dy, 0.5 * dz])
    arx = np.where(f_dir == 0, area[0], 0.0)
    ary = np.where(f_dir == 1, area[1], 0.0)
    arz = np.where(f_dir == 2, area[2], 0.0)
    xf = xc[f_own] + np.where(f_dir == 0, half[0], 0.0)
    yf = yc[f_own] + np.where(f_dir == 1, half[1], 0.0)
    zf = zc[f_own] + np.where(f_dir == 2, half[2], 0.0)
    # facint = |x_j' - x_P| / |x_N - x_P| (:1040-1062)
    dpn = np.sqrt((xc[f_nb] - xc[f_own]) ** 2 + (yc[f_nb] - yc[f_own]) ** 2 + (zc[f_nb] - zc[f_own]) ** 2)
    djn = np.sqrt((xf - xc[f_own]) ** 2 + (yf - yc[f_own]) ** 2 + (zf - zc[f_own]) ** 2)
    facint = djn / dpn

    # boundary patches
    patch_cells = [cid[ci == 0], cid[ci == nx - 1], cid[cj == 0], cid[cj == ny - 1], cid[ck == 0], cid[ck == nz - 1]]
    patch_axis = [0, 0, 1, 1, 2, 2]
    patch_sign = [-1.0, 1.0, -1.0, 1.0, -1.0, 1.0]
    porder = _group_patches(list(kinds))
    b_own, b_ar, b_xf = [], [], []
    counts: Dict[str, int] = {}
    starts: Dict[str, int] = {}
    pos = nin
    for p in porder:
        cells = patch_cells[p]
        ax, sg = patch_axis[p], patch_sign[p]
        ar = np.zeros((cells.size, 3))
        ar[:, ax] = sg * area[ax]
        cf = np.stack([xc[cells], yc[cells], zc[cells]], axis=1)
        cf[:, ax] += sg * half[ax]
        b_own.append(cells)
        b_ar.append(ar)
        b_xf.append(cf)
        k = kinds[p]
        if k not in counts:
            counts[k] = 0
            starts[k] = pos
        counts[k] += cells.size
        pos += cells.size
    b_own_a = np.concatenate(b_own)
    b_ar_a = np.concatenate(b_ar)
    b_xf_a = np.concatenate(b_xf)
    owner = np.concatenate([f_own, b_own_a]) + 1
    return Mesh(
        numCells=n, numInnerFaces=nin, numFaces=owner.size,
        owner=owner.astype(np.int32), neighbour=(f_nb + 1).astype(np.int32),
        xc=xc, yc=yc, zc=zc, vol=vol,
        arx=np.concatenate([arx, b_ar_a[:, 0]]), ary=np.concatenate([ary, b_ar_a[:, 1]]),
        arz=np.concatenate([arz, b_ar_a[:, 2]]),
        xf=np.concatenate([xf, b_xf_a[:, 0]]), yf=np.concatenate([yf, b_xf_a[:, 1]]),
        zf=np.concatenate([zf, b_xf_a[:, 2]]),
        facint=facint, counts=counts, starts=starts, gloCells=n)


def hex_polymesh_arrays(nx: int, ny: int, nz: int, lengths=(1.0, 1.0, 1.0), kinds=("wall",) * 6,
                        jitter: float = 0.0, seed: int = 12345):
    """Points + quad faces of the same hex box (optionally with jittered interior
    points => non-orthogonal, skewed cells).  Returns the inputs of
    :func:`geometry_from_polymesh`."""
    lx, ly, lz = lengths
    px, py, pz = nx + 1, ny + 1, nz + 1
    gi, gj, gk = np.meshgrid(np.arange(px), np.arange(py), np.arange(pz), indexing="ij")
    pid = gi + px * (gj + py * gk)
    pts = np.zeros((px * py * pz, 3))
    pts[pid.ravel(), 0] = (gi * (lx / nx)).ravel()
    pts[pid.ravel(), 1] = (gj * (ly / ny)).ravel()
    pts[pid.ravel(), 2] = (gk * (lz / nz)).ravel()
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        interior = ((gi > 0) & (gi < nx) & (gj > 0) & (gj < ny) & (gk > 0) & (gk < nz)).ravel()
        d = rng.uniform(-1.0, 1.0, size=(px * py * pz, 3)) * jitter * np.array([lx / nx, ly / ny, lz / nz])
        pts[pid.ravel()[interior]] += d[pid.ravel()[interior]]

    def P(i, j, k):
        return i + px * (j + py * k)

    m = hex_mesh(nx, ny, nz, lengths, kinds)
    own0 = m.owner.astype(np.int64) - 1
    ci, cj, ck = own0 % nx, (own0 // nx) % ny, own0 // (nx * ny)
    # direction / side of every face from its area vector
    axis = np.argmax(np.abs(np.stack([m.arx, m.ary, m.arz], axis=1)), axis=1)
    sign = np.sign(np.stack([m.arx, m.ary, m.arz], axis=1)[np.arange(m.numFaces), axis])
    faces = np.zeros((m.numFaces, 4), dtype=np.int64)
    for ax in range(3):
        for sg in (-1.0, 1.0):
            sel = (axis == ax) & (sign == sg)
            i, j, k = ci[sel], cj[sel], ck[sel]
            if ax == 0:
                i0 = i + (1 if sg > 0 else 0)
                quad = [P(i0, j, k), P(i0, j + 1, k), P(i0, j + 1, k + 1), P(i0, j, k + 1)]
            elif ax == 1:
                j0 = j + (1 if sg > 0 else 0)
                quad = [P(i, j0, k), P(i, j0, k + 1), P(i + 1, j0, k + 1), P(i + 1, j0, k)]
            else:
                k0 = k + (1 if sg > 0 else 0)
                quad = [P(i, j, k0), P(i + 1, j, k0), P(i + 1, j + 1, k0), P(i, j + 1, k0)]
            q = np.stack(quad, axis=1)
            if sg < 0:
                q = q[:, ::-1]  # outward normal for the low-side patches
            faces[sel] = q
    return pts, faces, m.owner.copy(), m.neighbour.copy(), dict(m.counts), dict(m.starts)


# --------------------------------------------------------------------------
# geometry from points/faces: restates mesh_geometry (:859-1081)
# --------------------------------------------------------------------------
def geometry_from_polymesh(points: np.ndarray, faces, owner: np.ndarray, neighbour: np.ndarray,
                           counts: Dict[str, int], starts: Dict[str, int]) -> Mesh:
    """Face area vectors by fan triangulation about node 1, face centre = node
    average, cell volume / centroid by the pyramid sums of the reference, and
    ``facint`` from the intersection of the P-N line with the plane of the first
    three face nodes.  ``faces`` is an [nFaces, k] array or a list of node arrays
    (0-based nodes); ``owner``/``neighbour`` are 1-based."""
    nF = len(owner)
    nI = len(neighbour)
    own = owner.astype(np.int64) - 1
    nb = neighbour.astype(np.int64) - 1
    n = int(own.max()) + 1
    if isinstance(faces, np.ndarray) and faces.ndim == 2:
        groups = [(np.arange(nF), faces)]
    else:
        lens = np.array([len(f) for f in faces])
        groups = []
        for k in np.unique(lens):
            idx = np.nonzero(lens == k)[0]
            groups.append((idx, np.array([faces[i] for i in idx], dtype=np.int64)))
    ar = np.zeros((nF, 3))
    cf = np.zeros((nF, 3))
    vol = np.zeros(n)
    cen = np.zeros((n, 3))
    first3 = np.zeros((nF, 3), dtype=np.int64)
    tri_store = []
    for idx, fa in groups:
        k = fa.shape[1]
        p1 = points[fa[:, 0]]
        first3[idx] = fa[:, :3]
        cf[idx] = points[fa].sum(axis=1) / float(k)
        for t in range(k - 2):
            a = points[fa[:, t + 1]] - p1
            b = points[fa[:, t + 2]] - p1
            nrm = np.cross(a, b)                      # triangular_face_area_components_polymesh
            ar[idx] += 0.5 * nrm
            c = (points[fa[:, t + 2]] + points[fa[:, t + 1]] + p1) / 3.0
            dv = (c * nrm).sum(axis=1) / 6.0          # cell_volume_part_polymesh
            np.add.at(vol, own[idx], dv)
            inner = idx < nI
            np.add.at(vol, nb[idx[inner]], -dv[inner])
            tri_store.append((idx, fa[:, 0], fa[:, t + 1], fa[:, t + 2], nrm))
    for idx, n1, n2, n3, nrm in tri_store:
        # centroid_component_part_polymesh: 1/(2*vol) * 1/24 * n_c * ((a+b)^2+(b+c)^2+(c+a)^2)
        pa, pb, pc = points[n1], points[n2], points[n3]
        s = (pa + pb) ** 2 + (pb + pc) ** 2 + (pc + pa) ** 2
        contrib = nrm * s / 48.0
        np.add.at(cen, own[idx], contrib / vol[own[idx], None])
        inner = idx < nI
        np.add.at(cen, nb[idx[inner]], -contrib[inner] / vol[nb[idx[inner]], None])
    # interpolation factor (:1040-1062)
    io, inn = own[:nI], nb[:nI]
    a0 = points[first3[:nI, 0]]
    pn = np.cross(points[first3[:nI, 1]] - a0, points[first3[:nI, 2]] - a0)
    d = cen[inn] - cen[io]
    t = ((a0 - cen[io]) * pn).sum(axis=1) / (d * pn).sum(axis=1)
    xj = cen[io] + t[:, None] * d
    djn = np.sqrt(((xj - cen[io]) ** 2).sum(axis=1))
    dpn = np.sqrt((d ** 2).sum(axis=1))
    facint = djn / dpn
    return Mesh(numCells=n, numInnerFaces=nI, numFaces=nF, owner=owner.astype(np.int32),
                neighbour=neighbour.astype(np.int32),
                xc=cen[:, 0].copy(), yc=cen[:, 1].copy(), zc=cen[:, 2].copy(), vol=vol,
                arx=ar[:, 0].copy(), ary=ar[:, 1].copy(), arz=ar[:, 2].copy(),
                xf=cf[:, 0].copy(), yf=cf[:, 1].copy(), zf=cf[:, 2].copy(),
                facint=facint, counts=dict(counts), starts=dict(starts), gloCells=n)


def _foam_body(text: str) -> str:
    """Strip the FoamFile header; return from the entry count on."""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//.*", "", text)
    i = text.find("}")
    return text[i + 1:] if "FoamFile" in text[:i + 1] else text


def read_polymesh(polymesh_dir: str) -> Mesh:
    """Read OpenFOAM ASCII ``points/faces/owner/neighbour`` plus the reference's
    simplified ``boundary`` file (``#type nFaces startFace``, :414-470)."""
    def body(name):
        with open(os.path.join(polymesh_dir, name)) as fh:
            return _foam_body(fh.read())
    b = body("points")
    npts = int(re.search(r"(\d+)\s*\(", b).group(1))
    nums = re.findall(r"\(\s*([-+0-9.eE]+)\s+([-+0-9.eE]+)\s+([-+0-9.eE]+)\s*\)", b)
    points = np.array(nums, dtype=np.float64)
    assert points.shape[0] == npts
    b = body("faces")
    faces = [np.array(m.split(), dtype=np.int64) for m in re.findall(r"\d+\(([\d\s]+)\)", b)]
    def labels(name):
        b = body(name)
        m = re.search(r"(\d+)\s*\(([\d\s]*)\)", b)
        arr = np.array(m.group(2).split(), dtype=np.int64)
        assert arr.size == int(m.group(1))
        return arr
    owner = labels("owner") + 1
    neighbour = labels("neighbour") + 1
    counts: Dict[str, int] = {}
    starts: Dict[str, int] = {}
    with open(os.path.join(polymesh_dir, "boundary")) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            kind, nf, st = line.split()[:3]
            kind = {"wallIsoth": "wall", "wallAdiab": "wall", "wallQFlux": "wall"}.get(kind, kind)
            if kind not in KINDS:
                raise ValueError(f"boundary type {kind!r} is outside the supported path")
            if kind not in counts:
                counts[kind] = 0
                starts[kind] = int(st)
            counts[kind] += int(nf)
    if len({len(f) for f in faces}) == 1:
        faces = np.array(faces)
    return geometry_from_polymesh(points, faces, owner, neighbour, counts, starts)


# --------------------------------------------------------------------------
# synthetic non-orthogonal polyhedral mesh
# --------------------------------------------------------------------------
def _splitmix64(x: np.ndarray) -> np.ndarray:
    """Counter-based RNG keyed by cell id (values do not depend on how the mesh is partitioned)."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def bcc_poly_mesh(N: int, jitter: float = 0.15, seed: int = 2024, length: float = 1.0,
                  layers: Optional[Tuple[int, int]] = None) -> Mesh:
    """Voronoi tessellation of a body-centred cubic lattice over N^3 cubic unit cells: 2 N^3 truncated
    octahedra, each with 8 hexagonal faces towards the (+-1/2,+-1/2,+-1/2) neighbours and 6 square faces
    towards the (+-1,0,0) neighbours = 14 faces per cell (nnz/row = 15).  The geometry arrays are emitted
    directly (no points / faces files): face area vectors and face centres are those of the regular
    tessellation, cell centres are displaced by up to ``jitter`` x the nearest-neighbour spacing
    (splitmix64 keyed by the GLOBAL cell id), which makes every face non-orthogonal; ``facint`` follows from the
    displaced centres (intersection of the P-N line with the face plane, mesh_geometry...:1040-1062).
    A cell on the hull of the lattice is closed by ONE wall face carrying the sum of its missing faces'
    area vectors, so that sum(S) = 0 holds for every cell.  Cells are numbered unit cell by unit cell
    (corner site, then body centre; i fastest), faces in OpenFOAM's upper-triangular order.

    ``layers = (ka, kb)``: only the unit-cell layers ka <= k < kb are generated (cells keep their relative order,
    local id = global id - 2 N^2 ka; faces towards cells outside the range are left out, every per-cell quantity --
    centre, volume, wall closure -- is the global mesh's).  ``poly_slab_part`` builds one rank's partition from such
    a sub-lattice, so that a 20 M-cell case never needs the global mesh on every rank."""
    a = length / N
    ka, kb = (0, N) if layers is None else (max(0, int(layers[0])), min(N, int(layers[1])))
    nk = kb - ka
    n = 2 * N * N * nk
    id0 = 2 * N * N * ka
    # unit cells in numbering order (i fastest, then j, then k)
    ck, cj, ci = (v.ravel().astype(np.int64) for v in np.meshgrid(np.arange(ka, kb), np.arange(N), np.arange(N),
                                                                   indexing="ij"))
    # site positions in units of a/2: corner sites (type 0) at even, body centres (type 1) at odd coordinates
    P = np.empty((n, 3), dtype=np.int64)
    P[0::2] = np.stack([2 * ci, 2 * cj, 2 * ck], axis=1)
    P[1::2] = P[0::2] + 1
    del ci, cj, ck
    cid = np.arange(n, dtype=np.int64)          # local ids; global id = cid + id0

    def site_id(q):
        """GLOBAL cell id of the site at half-lattice coordinates q [m,3], or -1 outside the lattice"""
        t = q[:, 0] & 1
        same = ((q[:, 1] & 1) == t) & ((q[:, 2] & 1) == t)
        u = (q - t[:, None]) >> 1
        inside = same & np.all((u >= 0) & (u < N), axis=1)
        out = np.full(q.shape[0], -1, dtype=np.int64)
        uu = u[inside]
        out[inside] = 2 * (uu[:, 0] + N * (uu[:, 1] + N * uu[:, 2])) + t[inside]
        return out

    hexd = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.int64)
    sqd = np.array([[2, 0, 0], [-2, 0, 0], [0, 2, 0], [0, -2, 0], [0, 0, 2], [0, 0, -2]], dtype=np.int64)
    dirs = np.concatenate([hexd, sqd])                               # in units of a/2
    a_hex, a_sq = 3.0 * a * a / 16.0, a * a / 8.0                    # |S_hex| / sqrt(3) per component, |S_sq|
    Sdir = np.concatenate([hexd * a_hex, (sqd // 2) * a_sq]).astype(np.float64)

    f_own, f_nb, f_dir = [], [], []
    bS = np.zeros((n, 3)); bC = np.zeros((n, 3)); bW = np.zeros(n)
    vsum = np.zeros(n)                                               # sum over the cell's faces of (c . S_outward) / 3
    for di, (d, S) in enumerate(zip(dirs, Sdir)):
        nb = site_id(P + d) 
        ctr = (P + 0.5 * d) * (0.5 * a)                              # face centre = midpoint of the two sites
        vsum += (ctr * S).sum(axis=1) / 3.0 * (nb >= 0)
        miss = nb < 0
        w = float(np.sqrt((S * S).sum()))
        bS[miss] += S; bC[miss] += w * ctr[miss]; bW[miss] += w
        nbl = nb - id0
        keep = (nbl > cid) & (nbl < n)                               # every inner face once, owner < neighbour
        f_own.append(cid[keep]); f_nb.append(nbl[keep])
        f_dir.append(np.full(int(keep.sum()), di, dtype=np.int8))
        del nb, nbl, ctr, keep, miss
    f_own = np.concatenate(f_own); f_nb = np.concatenate(f_nb); f_dir = np.concatenate(f_dir)
    o = np.lexsort((f_nb, f_own))
    f_own, f_nb, f_dir = f_own[o], f_nb[o], f_dir[o]
    del o
    nin = f_own.size
    f_S = Sdir[f_dir]
    f_c = (P[f_own] + 0.5 * dirs[f_dir]) * (0.5 * a)
    del f_dir
    # displaced cell centres
    spacing = a * np.sqrt(3.0) / 2.0
    cen = P * (0.5 * a)
    if jitter > 0.0:
        gid = cid + id0
        u = np.stack([_splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + (gid * 3 + c).astype(np.uint64))
                      for c in range(3)], axis=1)
        u = (u >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)     # [0,1)
        cen = cen + (2.0 * u - 1.0) * (jitter * spacing / np.sqrt(3.0))
        del u, gid
    del P
    dPN = cen[f_nb] - cen[f_own]
    facint = ((f_c - cen[f_own]) * f_S).sum(axis=1) / (dPN * f_S).sum(axis=1)
    del dPN
    bcell = np.nonzero(bW > 0)[0]
    b_S = bS[bcell]
    b_c = bC[bcell] / bW[bcell, None]
    owner = np.concatenate([f_own, bcell]) + 1
    # volumes: a^3/2 inside; a hull cell is the polyhedron its own faces close (divergence theorem over its existing
    # faces and its closing wall face)
    vol = np.full(n, 0.5 * a ** 3)
    vol[bcell] = vsum[bcell] + (b_c * b_S).sum(axis=1) / 3.0
    del bS, bC, bW, vsum
    arx = np.concatenate([f_S[:, 0], b_S[:, 0]]); ary = np.concatenate([f_S[:, 1], b_S[:, 1]])
    arz = np.concatenate([f_S[:, 2], b_S[:, 2]])
    del f_S
    xf = np.concatenate([f_c[:, 0], b_c[:, 0]]); yf = np.concatenate([f_c[:, 1], b_c[:, 1]])
    zf = np.concatenate([f_c[:, 2], b_c[:, 2]])
    del f_c
    return Mesh(
        numCells=n, numInnerFaces=nin, numFaces=owner.size,
        owner=owner.astype(np.int32), neighbour=(f_nb + 1).astype(np.int32),
        xc=cen[:, 0].copy(), yc=cen[:, 1].copy(), zc=cen[:, 2].copy(), vol=vol,
        arx=arx, ary=ary, arz=arz, xf=xf, yf=yf, zf=zf,
        facint=facint, counts={"wall": int(bcell.size)}, starts={"wall": int(nin)}, gloCells=2 * N ** 3)


def poly_slab_part(N: int, rank: int, nranks: int, jitter: float = 0.15, seed: int = 2024) -> Mesh:
    """Rank ``rank``'s mesh of ``partition(bcc_poly_mesh(N), slab_ranks(2 N^3, nranks), nranks)`` without the global
    mesh: the rank's unit-cell layers plus one ghost layer on either side are generated (``layers`` of
    ``bcc_poly_mesh``) and cut with the same routine, so every array equals the global partition's
    (tests/test_poly_mesh.py); ``cell_global`` etc. are shifted back to global numbers.  N must be a multiple of
    ``nranks`` (cell-id blocks = whole layers)."""
    if N % nranks:
        raise ValueError("poly_slab_part: N must be a multiple of the number of ranks")
    per = N // nranks
    k0, k1 = rank * per, (rank + 1) * per
    ka, kb = max(0, k0 - 1), min(N, k1 + 1)
    sub = bcc_poly_mesh(N, jitter, seed, layers=(ka, kb))
    id0 = 2 * N * N * ka
    ntot = 2 * N ** 3
    cell_rank = ((np.arange(sub.numCells, dtype=np.int64) + id0) * nranks) // ntot
    part = partition(sub, cell_rank, nranks, only=rank)[0]
    part.gloCells = ntot
    part.cell_global = part.cell_global + id0
    part.halo_global = part.halo_global + id0
    part.face_global = None            # numbers of the sub-lattice's face list: meaningless outside this function
    return part


def rcb_ranks(g: Mesh, nranks: int) -> np.ndarray:
    """Recursive coordinate bisection of the cell centres into ``nranks`` parts of (nearly) equal size:
    the partitioner for unstructured meshes (the reference relies on OpenFOAM's decomposePar)."""
    rank = np.zeros(g.numCells, dtype=np.int64)
    xyz = np.stack([g.xc[:g.numCells], g.yc[:g.numCells], g.zc[:g.numCells]], axis=1)

    def split(idx, r0, k):
        if k == 1:
            rank[idx] = r0
            return
        kl = k // 2
        ax = int(np.argmax(np.ptp(xyz[idx], axis=0)))
        o = idx[np.argsort(xyz[idx, ax], kind="stable")]
        cut = (o.size * kl) // k
        split(o[:cut], r0, kl)
        split(o[cut:], r0 + kl, k - kl)

    split(np.arange(g.numCells), 0, nranks)
    return rank


# --------------------------------------------------------------------------
# cell partitioner: per-rank meshes in the src-parallel layout
# --------------------------------------------------------------------------
def slab_ranks(mesh_n: int, nranks: int) -> np.ndarray:
    """Contiguous blocks of cell ids (z-slabs for the hex numbering)."""
    return (np.arange(mesh_n, dtype=np.int64) * nranks) // mesh_n


def partition(g: Mesh, cell_rank: np.ndarray, nranks: int, only: Optional[int] = None) -> List[Mesh]:
    """Split ``g`` by ``cell_rank`` into per-rank meshes laid out like an
    OpenFOAM decomposition read by src-parallel: faces = [inner | boundary
    patches in the global order | processor faces grouped by neighbour rank
    ascending], slots = [cells | npro halo | inlet | outlet | symmetry | wall |
    prOutlet] (:637-660).  Processor faces keep the global face order on both
    sides, so buffer position i of a connection pairs with position i of the
    mirrored connection (exchange.f90:49-90).  ``fpro`` as in :1183-1250."""
    own = g.owner.astype(np.int64) - 1
    nb = g.neighbour.astype(np.int64) - 1
    nI = g.numInnerFaces
    r_own = cell_rank[own]
    r_nb = cell_rank[nb]
    out: List[Mesh] = []
    for r in (range(nranks) if only is None else [only]):   # `only`: build just that rank's mesh
        cells = np.nonzero(cell_rank == r)[0]
        g2l = np.full(g.numCells, -1, dtype=np.int64)
        g2l[cells] = np.arange(cells.size)
        fin = np.nonzero((r_own[:nI] == r) & (r_nb == r))[0]
        cut_o = np.nonzero((r_own[:nI] == r) & (r_nb != r))[0]   # we own, other side remote
        cut_n = np.nonzero((r_own[:nI] != r) & (r_nb == r))[0]   # we are neighbour: flip
        # boundary faces of this rank, kept in global face order inside each kind
        bfaces, counts, starts = [], {}, {}
        pos = fin.size
        # kinds must be traversed in global *face* order of their first patch
        for kind in sorted(g.counts, key=lambda k: g.starts[k]):
            fr = np.arange(g.starts[kind], g.starts[kind] + g.counts[kind])
            fr = fr[r_own[fr] == r]
            counts[kind] = fr.size
            starts[kind] = pos
            pos += fr.size
            bfaces.append(fr)
        bfaces = np.concatenate(bfaces) if bfaces else np.zeros(0, dtype=np.int64)
        # processor faces grouped by neighbour rank
        cut = np.concatenate([cut_o, cut_n])
        flip = np.concatenate([np.zeros(cut_o.size, bool), np.ones(cut_n.size, bool)])
        other = np.where(flip, r_own[cut], r_nb[cut])
        o = np.lexsort((cut, other))
        cut, flip, other = cut[o], flip[o], other[o]
        nbr_ranks = np.unique(other)
        offs = [1]
        for q in nbr_ranks:
            offs.append(offs[-1] + int((other == q).sum()))
        npro = cut.size
        sgn = np.where(flip, -1.0, 1.0)
        loc_cell = np.where(flip, nb[cut], own[cut])
        rem_cell = np.where(flip, own[cut], nb[cut])
        allf = np.concatenate([fin, bfaces, cut])
        fsign = np.concatenate([np.ones(fin.size + bfaces.size), sgn])
        owner_l = np.concatenate([g2l[own[fin]], g2l[own[bfaces]], g2l[loc_cell]]) + 1
        neigh_l = g2l[nb[fin]] + 1
        nloc = cells.size
        # geometry; halo copies of xc,yc,zc,vol sit in [nloc, nloc+npro) (:1112-1115)
        xc = np.concatenate([g.xc[cells], g.xc[rem_cell]])
        yc = np.concatenate([g.yc[cells], g.yc[rem_cell]])
        zc = np.concatenate([g.zc[cells], g.zc[rem_cell]])
        vol = np.concatenate([g.vol[cells], g.vol[rem_cell]])
        # fpro: distance-weighted factor towards the remote cell; for a face we own it
        # equals the global facint, for a flipped face 1 - facint
        fpro = np.where(flip, 1.0 - g.facint[cut], g.facint[cut])
        out.append(Mesh(
            numCells=nloc, numInnerFaces=fin.size, numFaces=allf.size,
            owner=owner_l.astype(np.int32), neighbour=neigh_l.astype(np.int32),
            xc=xc, yc=yc, zc=zc, vol=vol,
            arx=g.arx[allf] * fsign, ary=g.ary[allf] * fsign, arz=g.arz[allf] * fsign,
            xf=g.xf[allf].copy(), yf=g.yf[allf].copy(), zf=g.zf[allf].copy(),
            facint=g.facint[fin].copy(), counts=counts, starts=starts,
            npro=npro, iProcFacesStart=fin.size + bfaces.size, fpro=fpro,
            neighbProcNo=nbr_ranks.astype(np.int32), neighbProcOffset=np.array(offs, dtype=np.int32),
            gloCells=g.numCells, cell_global=cells, halo_global=rem_cell, face_global=allf))
    return out


def scatter_total(g: Mesh, part: Mesh, arr: np.ndarray) -> np.ndarray:
    """Global numTotal-sized field -> the rank's numTotal-sized field: own cells, halo slots
    (value of the remote cell) and boundary slots."""
    out = np.zeros(part.numTotal)
    n = part.numCells
    out[:n] = arr[part.cell_global]
    out[n:n + part.npro] = arr[part.halo_global]
    for kind in KINDS:
        c = part.count(kind)
        if c == 0:
            continue
        gf = part.face_global[part.faces_start(kind):part.faces_start(kind) + c]
        out[part.slot_start(kind):part.slot_start(kind) + c] = arr[g.slot_start(kind) + (gf - g.faces_start(kind))]
    return out


def scatter_cells(g: Mesh, part: Mesh, arr: np.ndarray, width: int = 1) -> np.ndarray:
    """Global per-cell field (numCells[, width]) -> the rank's numCells+npro field incl. halo copies."""
    idx = np.concatenate([part.cell_global, part.halo_global]) if part.npro else part.cell_global
    return np.ascontiguousarray(arr[idx])


def gather_cells(g: Mesh, parts: List[Mesh], arrs: List[np.ndarray]) -> np.ndarray:
    """Per-rank cell fields -> global cell field."""
    first = np.asarray(arrs[0])
    out = np.zeros((g.numCells,) + first.shape[1:])
    for m, a in zip(parts, arrs):
        out[m.cell_global] = np.asarray(a)[:m.numCells]
    return out


def scatter_faces(part: Mesh, arr: np.ndarray) -> np.ndarray:
    """Global per-boundary-kind face list helper: values of a global per-face array at the rank's faces."""
    return np.ascontiguousarray(arr[part.face_global])


# --------------------------------------------------------------------------
# polyMesh / decomposition writers
# --------------------------------------------------------------------------
_FOAM_HEADER = """/*--------------------------------*- C++ -*----------------------------------*\\
| =========                 |                                                 |
| \\\\      /  F ield         | freeCappuccino polyMesh (OpenFOAM ASCII layout)  |
\\*---------------------------------------------------------------------------*/
FoamFile
{
    version     2.0;
    format      ascii;
    class       %s;
    location    "constant/polyMesh";
    object      %s;
}
// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //

"""


def write_polymesh(polymesh_dir: str, points: np.ndarray, faces, owner0: np.ndarray, neighbour0: np.ndarray,
                   patches: Sequence[Tuple[str, int, int]]) -> None:
    """Write ``points / faces / owner / neighbour`` in OpenFOAM ASCII and the reference's simplified ``boundary``
    table (``#type nFaces startFace``): what ``read_polymesh`` here, ``mesh_geometry`` of the reference
    (src/mesh_geometry_and_topology.f90:395-470, 540-600) and host/fcapp_mesh.cpp read.  ``owner0`` /
    ``neighbour0`` are 0-based like the files."""
    os.makedirs(polymesh_dir, exist_ok=True)
    with open(os.path.join(polymesh_dir, "points"), "w") as fh:
        fh.write(_FOAM_HEADER % ("vectorField", "points") + f"\n{len(points)}\n(\n")
        fh.writelines("(%s %s %s)\n" % (repr(float(p[0])), repr(float(p[1])), repr(float(p[2]))) for p in points)
        fh.write(")\n")
    with open(os.path.join(polymesh_dir, "faces"), "w") as fh:
        fh.write(_FOAM_HEADER % ("faceList", "faces") + f"\n{len(faces)}\n(\n")
        fh.writelines("%d(%s)\n" % (len(f), " ".join(str(int(v)) for v in f)) for f in faces)
        fh.write(")\n")
    for name, arr in (("owner", owner0), ("neighbour", neighbour0)):
        with open(os.path.join(polymesh_dir, name), "w") as fh:
            fh.write(_FOAM_HEADER % ("labelList", name) + f"\n{len(arr)}\n(\n")
            fh.write("\n".join(str(int(v)) for v in arr) + ("\n" if len(arr) else "") + ")\n")
    with open(os.path.join(polymesh_dir, "boundary"), "w") as fh:
        fh.write("#type nFaces startFace\n")
        fh.writelines(f"{kind} {int(nf)} {int(st)}\n" for kind, nf, st in patches)


def _write_labels(path: str, name: str, arr: np.ndarray) -> None:
    with open(path, "w") as fh:
        fh.write(_FOAM_HEADER % ("labelList", name) + f"\n{len(arr)}\n(\n")
        fh.write("\n".join(str(int(v)) for v in arr) + ("\n" if len(arr) else "") + ")\n")


def write_decomposition(case_dir: str, points: np.ndarray, faces, g: Mesh, cell_rank: np.ndarray, nranks: int,
                        patches: Sequence[Tuple[str, int, int]]) -> List[Mesh]:
    """The decomposition ``src-parallel`` reads (examples/*/…-setup-parallel.tar.gz): for every rank
    ``processor<r>/constant/polyMesh/`` with points, faces, owner, neighbour, the ``boundary`` table of the rank's
    share of every global patch (empty patches keep their row, like decomposePar), the ``process`` file
    (``neighbProcNo nfaces startFace`` per connection, src-parallel/mesh_geometry_and_topology.f90:497-515) and
    OpenFOAM's ``cellProcAddressing`` / ``faceProcAddressing`` (1-based global face, negative when the rank sees
    the face flipped).  ``patches`` is the global boundary table; returns the per-rank meshes of ``partition``."""
    parts = partition(g, cell_rank, nranks)
    nI = g.numInnerFaces
    for r, p in enumerate(parts):
        d = os.path.join(case_dir, f"processor{r}", "constant", "polyMesh")
        gf = p.face_global
        flip = np.zeros(gf.size, bool)
        if p.npro:   # processor faces whose global owner is remote are seen reversed
            pf = gf[p.iProcFacesStart:]
            flip[p.iProcFacesStart:] = cell_rank[g.owner[pf].astype(np.int64) - 1] != r
        used = np.unique(np.concatenate([np.asarray(faces[f], dtype=np.int64) for f in gf]))
        g2l = np.full(len(points), -1, dtype=np.int64)
        g2l[used] = np.arange(used.size)
        lf = []
        for f, fl in zip(gf, flip):
            nodes = g2l[np.asarray(faces[f], dtype=np.int64)]
            # OpenFOAM's reverseFace: keep the first node, reverse the rest -- the fan triangulation about node 1
            # then consists of the same triangles, so the area vector and the volume sums are exactly negated
            lf.append(np.concatenate([nodes[:1], nodes[:0:-1]]) if fl else nodes)
        # boundary rows: every global patch keeps a row; rows of equal kind are contiguous in the rank's numbering
        rows, pos = [], p.numInnerFaces
        for kind, nf, st in sorted(patches, key=lambda t: t[2]):
            mine = int(np.count_nonzero((gf[p.numInnerFaces:p.iProcFacesStart] >= st)
                                        & (gf[p.numInnerFaces:p.iProcFacesStart] < st + nf)))
            rows.append((kind, mine, pos))
            pos += mine
        write_polymesh(d, points[used], lf, p.owner - 1, p.neighbour - 1, rows)
        with open(os.path.join(d, "process"), "w") as fh:
            fh.write("# neighbProcNo nfaces startFace\n%d\n" % len(p.neighbProcNo))
            for c, q in enumerate(p.neighbProcNo):
                s, e = int(p.neighbProcOffset[c]), int(p.neighbProcOffset[c + 1])
                fh.write("%d %d %d\n" % (int(q), e - s, p.iProcFacesStart + s - 1))
        _write_labels(os.path.join(d, "cellProcAddressing"), "cellProcAddressing", p.cell_global)
        _write_labels(os.path.join(d, "faceProcAddressing"), "faceProcAddressing", np.where(flip, -(gf + 1), gf + 1))
    return parts
