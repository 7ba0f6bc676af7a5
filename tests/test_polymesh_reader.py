"""host/fcapp_mesh.cpp -- the C++ mirror of the reference's polyMesh reader + mesh_geometry --
against the Python mesh module on the shipped example meshes and a skewed synthetic one.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

from freecappuccino_b200 import cases
from freecappuccino_b200 import mesh as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "meshdump")
GOLD = os.path.join(os.path.dirname(__file__), "golden")

HEADER = """/*--------------------------------*- C++ -*----------------------------------*\\
| =========                 |                                                 |
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


def write_polymesh(d, points, faces, owner0, neighbour0, patches):
    """OpenFOAM ASCII points / faces / owner / neighbour + the reference's simplified boundary table."""
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "points"), "w") as fh:
        fh.write(HEADER % ("vectorField", "points") + f"\n{len(points)}\n(\n")
        for p in points:
            fh.write("(%s %s %s)\n" % tuple(repr(float(v)) for v in p))
        fh.write(")\n\n// ************************************************************************* //\n")
    with open(os.path.join(d, "faces"), "w") as fh:
        fh.write(HEADER % ("faceList", "faces") + f"\n{len(faces)}\n(\n")
        for f in faces:
            fh.write("%d(%s)\n" % (len(f), " ".join(str(int(v)) for v in f)))
        fh.write(")\n")
    for name, arr in (("owner", owner0), ("neighbour", neighbour0)):
        with open(os.path.join(d, name), "w") as fh:
            fh.write(HEADER % ("labelList", name) + f"\n{len(arr)}\n(\n")
            fh.write("\n".join(str(int(v)) for v in arr) + "\n)\n")
    with open(os.path.join(d, "boundary"), "w") as fh:
        fh.write("#type nFaces startFace\n")
        for kind, nf, st in patches:
            fh.write(f"{kind} {nf} {st}\n")


def run_meshdump(d, tmp_path):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host"), "meshdump"])
    out = os.path.join(str(tmp_path), "mesh.bin")
    r = subprocess.run([EXE, d, out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    hdr = np.frombuffer(raw, np.int32, 13)
    n, F, NF = int(hdr[0]), int(hdr[1]), int(hdr[2])
    off = 13 * 4
    owner = np.frombuffer(raw, np.int32, NF, off); off += 4 * NF
    neigh = np.frombuffer(raw, np.int32, F, off); off += 4 * F
    arrs = {}
    for name, cnt in (("xc", n), ("yc", n), ("zc", n), ("vol", n), ("arx", NF), ("ary", NF), ("arz", NF), ("xf", NF),
                      ("yf", NF), ("zf", NF), ("facint", F)):
        arrs[name] = np.frombuffer(raw, np.float64, cnt, off); off += 8 * cnt
    assert off == len(raw)
    return hdr, owner, neigh, arrs


def compare(hdr, owner, neigh, arrs, ref):
    assert (int(hdr[0]), int(hdr[1]), int(hdr[2])) == (ref.numCells, ref.numInnerFaces, ref.numFaces)
    for k, kind in enumerate(M.KINDS):
        assert int(hdr[3 + k]) == ref.count(kind), kind
        if ref.count(kind):
            assert int(hdr[8 + k]) == ref.faces_start(kind), kind
    assert np.array_equal(owner, ref.owner) and np.array_equal(neigh, ref.neighbour)
    n = ref.numCells
    scale = float(np.abs(ref.xf).max())
    for name, a in arrs.items():
        b = getattr(ref, name)[:a.size]
        # same formulas, different accumulation order (face-major here and in the reference, triangle-major in
        # the vectorised Python module): agreement to round-off
        assert np.allclose(a, b, rtol=1e-11, atol=1e-13 * max(scale, np.abs(b).max())), name


@pytest.mark.parametrize("case", ["cavity", "pitzDaily"])
def test_reader_on_the_shipped_example_meshes(case, tmp_path):
    d = np.load(os.path.join(GOLD, case + ".npz"))
    patches = [(str(k), int(nf), int(st)) for k, nf, st in zip(d["bkind"], d["bn"], d["bstart"])]
    pm = os.path.join(str(tmp_path), "polyMesh")
    write_polymesh(pm, d["points"], d["faces"], d["owner"], d["neighbour"], patches)
    ref = cases.golden_mesh(os.path.join(GOLD, case + ".npz"))
    compare(*run_meshdump(pm, tmp_path), ref)
    # and the Python reader of the same directory agrees with itself
    ref2 = M.read_polymesh(pm)
    assert np.array_equal(ref2.vol, ref.vol)


def test_reader_on_a_skewed_mesh_with_every_boundary_kind(tmp_path):
    kinds = ("inlet", "outlet", "wall", "prOutlet", "symmetry", "symmetry")
    pts, faces, owner, neigh, counts, starts = M.hex_polymesh_arrays(7, 6, 5, (1.0, 0.8, 0.6), kinds, 0.2, 99)
    ref = M.geometry_from_polymesh(pts, faces, owner, neigh, counts, starts)
    patches = sorted(((k, counts[k], starts[k]) for k in counts), key=lambda t: t[2])
    pm = os.path.join(str(tmp_path), "polyMesh")
    write_polymesh(pm, pts, faces, np.asarray(owner) - 1, np.asarray(neigh) - 1, patches)
    compare(*run_meshdump(pm, tmp_path), ref)


def test_reader_rejects_what_the_path_does_not_cover(tmp_path):
    d = np.load(os.path.join(GOLD, "cavity.npz"))
    patches = [("cyclic", int(nf), int(st)) for nf, st in zip(d["bn"], d["bstart"])]
    pm = os.path.join(str(tmp_path), "polyMesh")
    write_polymesh(pm, d["points"], d["faces"], d["owner"], d["neighbour"], patches)
    r = subprocess.run([EXE, pm], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "cyclic" in r.stderr
