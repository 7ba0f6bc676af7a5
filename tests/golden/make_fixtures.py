#!/usr/bin/env python
"""Regenerates the mesh fixtures in this directory from the reference's shipped example cases.

Run in the build container (needs /root/reference; the GPU box never runs this):
    python tests/golden/make_fixtures.py

  cavity.npz      examples/cavity/cavity-setup.tar.gz      (400 cells)
  pitzDaily.npz   examples/pitzDaily/pitzDaily-setup.tar.gz (12 225 cells, config 2)
  cavity_par.npz  examples/cavity/cavity-setup-parallel.tar.gz: the reference's own 2-rank
                  decomposition (per-rank owner / neighbour / boundary / process files and
                  cell/faceProcAddressing) -- pins the partitioner's array layout.
  pitzDaily_par_cells.npz  examples/pitzDaily/pitzDaily-setup-parallel.tar.gz: which rank owns every cell of the shipped
                  2-rank decomposition (cellProcAddressing of both ranks -> cell_rank[12225]); freecappuccino_b200.mesh.
                  partition(pitzDaily, cell_rank) then rebuilds that decomposition (as tests/test_partition.py proves for
                  the cavity) -- the multi-rank lstsq_qr / limiter / PISO checks run on it.
Only mesh INPUT data is stored (points, faces, owner, neighbour, boundary table).
"""
import glob
import os
import re
import sys
import tarfile
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from freecappuccino_b200 import mesh as M  # noqa: E402

REF = "/root/reference/examples"


def labels(path):
    b = M._foam_body(open(path).read())
    m = re.search(r"(\d+)\s*\(([-\d\s]*)\)", b)
    a = np.array(m.group(2).split(), dtype=np.int32)
    assert a.size == int(m.group(1))
    return a


def raw_polymesh(d):
    b = M._foam_body(open(os.path.join(d, "points")).read())
    pts = np.array(re.findall(r"\(\s*([-+0-9.eE]+)\s+([-+0-9.eE]+)\s+([-+0-9.eE]+)\s*\)", b), dtype=np.float64)
    b = M._foam_body(open(os.path.join(d, "faces")).read())
    faces = np.array([m.split() for m in re.findall(r"\d+\(([\d\s]+)\)", b)], dtype=np.int32)
    kinds, nf, st = [], [], []
    text = open(os.path.join(d, "boundary")).read()
    if "FoamFile" in text:   # processor1 of the shipped decomposition still has the OpenFOAM-format file
        for m in re.finditer(r"type\s+(\w+);.*?nFaces\s+(\d+);\s*startFace\s+(\d+);", text, flags=re.S):
            if m.group(1) == "processor":
                continue
            kinds.append({"empty": "symmetry"}.get(m.group(1), m.group(1)))
            nf.append(int(m.group(2))); st.append(int(m.group(3)))
    else:                    # the reference's simplified table: "#type nFaces startFace"
        for line in text.splitlines():
            t = line.split()
            if line.startswith("#") or len(t) < 3:
                continue
            kinds.append(t[0]); nf.append(int(t[1])); st.append(int(t[2]))
    return dict(points=pts, faces=faces, owner=labels(os.path.join(d, "owner")),
                neighbour=labels(os.path.join(d, "neighbour")), bkind=np.array(kinds), bn=np.array(nf, np.int32),
                bstart=np.array(st, np.int32))


def main():
    for case in ("cavity", "pitzDaily"):
        with tempfile.TemporaryDirectory() as t:
            tarfile.open(os.path.join(REF, case, f"{case}-setup.tar.gz")).extractall(t)
            d = glob.glob(os.path.join(t, "**", "polyMesh"), recursive=True)[0]
            np.savez_compressed(os.path.join(HERE, f"{case}.npz"), **raw_polymesh(d))
    with tempfile.TemporaryDirectory() as t:
        tarfile.open(os.path.join(REF, "cavity", "cavity-setup-parallel.tar.gz")).extractall(t)
        out = {}
        for r in (0, 1):
            d = os.path.join(t, f"processor{r}", "constant", "polyMesh")
            raw = raw_polymesh(d)
            for k, v in raw.items():
                out[f"p{r}_{k}"] = v
            out[f"p{r}_cellProc"] = labels(os.path.join(d, "cellProcAddressing"))
            out[f"p{r}_faceProc"] = labels(os.path.join(d, "faceProcAddressing"))
            proc = [l.split() for l in open(os.path.join(d, "process")) if not l.startswith("#") and l.strip()]
            out[f"p{r}_process"] = np.array([[int(x) for x in row] for row in proc[1:]], dtype=np.int32)
        np.savez_compressed(os.path.join(HERE, "cavity_par.npz"), **out)
    with tempfile.TemporaryDirectory() as t:
        tarfile.open(os.path.join(REF, "pitzDaily", "pitzDaily-setup-parallel.tar.gz")).extractall(t)
        cells = [labels(os.path.join(t, f"processor{r}", "constant", "polyMesh", "cellProcAddressing")) for r in (0, 1)]
        rank = np.full(sum(c.size for c in cells), -1, dtype=np.int8)
        for r, c in enumerate(cells):
            rank[c] = r
        assert (rank >= 0).all()
        np.savez_compressed(os.path.join(HERE, "pitzDaily_par_cells.npz"), cell_rank=rank)
    for f in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        print(os.path.basename(f), os.path.getsize(f), "bytes")


if __name__ == "__main__":
    main()
