"""The decomposition writer (freecappuccino_b200.mesh.write_decomposition) against the
reference's own shipped 2-rank decomposition of the cavity case (tests/golden/cavity_par.npz =
examples/cavity/cavity-setup-parallel.tar.gz), and read back through both polyMesh readers.  CPU only."""
import os
import re
import subprocess

import numpy as np

from freecappuccino_b200 import cases
from freecappuccino_b200 import mesh as M

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def labels(path):
    b = M._foam_body(open(path).read())
    m = re.search(r"(\d+)\s*\(([-\d\s]*)\)", b)
    a = np.array(m.group(2).split(), dtype=np.int64)
    assert a.size == int(m.group(1))
    return a


def test_writer_reproduces_the_shipped_decomposition(tmp_path):
    d = np.load(os.path.join(GOLD, "cavity.npz"))
    par = np.load(os.path.join(GOLD, "cavity_par.npz"))
    g = cases.golden_mesh(os.path.join(GOLD, "cavity.npz"))
    cell_rank = np.zeros(g.numCells, dtype=np.int64)
    cell_rank[par["p1_cellProc"]] = 1
    patches = [(str(k), int(nf), int(st)) for k, nf, st in zip(d["bkind"], d["bn"], d["bstart"])]
    M.write_decomposition(str(tmp_path), d["points"], d["faces"], g, cell_rank, 2, patches)
    for r in range(2):
        pm = os.path.join(str(tmp_path), f"processor{r}", "constant", "polyMesh")
        assert np.array_equal(labels(os.path.join(pm, "owner")), par[f"p{r}_owner"])
        assert np.array_equal(labels(os.path.join(pm, "neighbour")), par[f"p{r}_neighbour"])
        assert np.array_equal(labels(os.path.join(pm, "cellProcAddressing")), par[f"p{r}_cellProc"])
        assert np.array_equal(labels(os.path.join(pm, "faceProcAddressing")), par[f"p{r}_faceProc"])
        rows = [l.split() for l in open(os.path.join(pm, "boundary")) if not l.startswith("#") and l.strip()]
        assert [(k, int(n), int(s)) for k, n, s in rows] == [
            (str(k), int(n), int(s)) for k, n, s in zip(par[f"p{r}_bkind"], par[f"p{r}_bn"], par[f"p{r}_bstart"])]
        proc = [l.split() for l in open(os.path.join(pm, "process")) if not l.startswith("#") and l.strip()]
        assert int(proc[0][0]) == 1 and [int(v) for v in proc[1]] == [int(v) for v in par[f"p{r}_process"][0]]
        # faces: the rank's own point numbering may differ from decomposePar's; the node COORDINATES in face order do not
        b = M._foam_body(open(os.path.join(pm, "points")).read())
        pts = np.array(re.findall(r"\(\s*([-+0-9.eE]+)\s+([-+0-9.eE]+)\s+([-+0-9.eE]+)\s*\)", b), dtype=np.float64)
        b = M._foam_body(open(os.path.join(pm, "faces")).read())
        fc = np.array([m.split() for m in re.findall(r"\d+\(([\d\s]+)\)", b)], dtype=np.int64)
        ref_xyz = par[f"p{r}_points"][par[f"p{r}_faces"]]
        got_xyz = pts[fc]
        # node by node, flipped processor faces included (decomposePar reverses them with face::reverseFace:
        # first node kept, the others in reverse order)
        assert np.array_equal(got_xyz, ref_xyz)
        own = M.read_polymesh(pm)
        ref = cases.golden_mesh(os.path.join(GOLD, "cavity_par.npz"), prefix=f"p{r}_")
        assert np.allclose(own.arx, ref.arx, atol=1e-16) and np.allclose(own.ary, ref.ary, atol=1e-16)
        assert np.allclose(own.vol, ref.vol, rtol=1e-12) and np.allclose(own.xc, ref.xc, atol=1e-14)


def test_written_ranks_are_readable_by_the_cpp_reader(tmp_path):
    kinds = ("inlet", "outlet", "wall", "prOutlet", "symmetry", "symmetry")
    pts, faces, owner, neigh, counts, starts = M.hex_polymesh_arrays(6, 5, 4, (1.0, 0.8, 0.6), kinds, 0.15, 7)
    g = M.geometry_from_polymesh(pts, faces, owner, neigh, counts, starts)
    patches = sorted(((k, counts[k], starts[k]) for k in counts), key=lambda t: t[2])
    parts = M.write_decomposition(str(tmp_path), pts, faces, g, M.slab_ranks(g.numCells, 3), 3, patches)
    exe = os.path.join(ROOT, "host", "meshdump")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host"), "meshdump"])
    for r, p in enumerate(parts):
        pm = os.path.join(str(tmp_path), f"processor{r}", "constant", "polyMesh")
        # the rank's processor faces are boundary faces without a patch row: both readers need them declared, so the
        # serial readers are given the rank mesh with its processor faces appended as a wall patch
        rows = [l for l in open(os.path.join(pm, "boundary"))]
        if p.npro:
            rows.append(f"wall {p.npro} {p.iProcFacesStart}\n")
        open(os.path.join(pm, "boundary"), "w").writelines(rows)
        out = subprocess.run([exe, pm], capture_output=True, text=True, timeout=60)
        assert out.returncode == 0, out.stderr
        assert f"cells {p.numCells} innerFaces {p.numInnerFaces} faces {p.numFaces}" in out.stdout
        own = M.read_polymesh(pm)
        assert np.allclose(own.vol, p.vol[:p.numCells], rtol=1e-11)
        assert np.allclose(own.arx, p.arx, atol=1e-15) and np.allclose(own.xf, p.xf, atol=1e-14)


def test_decompose_cli_on_the_shipped_cavity_mesh(tmp_path):
    import sys
    d = np.load(os.path.join(GOLD, "cavity.npz"))
    patches = [(str(k), int(nf), int(st)) for k, nf, st in zip(d["bkind"], d["bn"], d["bstart"])]
    M.write_polymesh(os.path.join(str(tmp_path), "polyMesh"), d["points"], d["faces"], d["owner"], d["neighbour"], patches)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "decompose.py"), str(tmp_path), "4"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    cells = 0
    for r in range(4):
        cellproc = labels(os.path.join(str(tmp_path), f"processor{r}", "constant", "polyMesh", "cellProcAddressing"))
        cells += cellproc.size
    assert cells == 400 and len(out.stdout.strip().splitlines()) == 4
