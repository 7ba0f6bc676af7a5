#!/usr/bin/env python
"""Decompose a polyMesh case for the multi-rank build (what OpenFOAM's decomposePar + the reference's hand-written
`process` file provide for src-parallel):

    python tools/decompose.py <case dir with polyMesh/> <nranks> [--method rcb|slab]

reads `<case>/polyMesh/{points,faces,owner,neighbour,boundary}` and writes `<case>/processor<r>/constant/polyMesh/`
(points, faces, owner, neighbour, boundary, process, cellProcAddressing, faceProcAddressing) for every rank.
`rcb` = recursive coordinate bisection of the cell centres (default), `slab` = contiguous blocks of cell ids."""
import argparse
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freecappuccino_b200 import mesh as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("nranks", type=int)
    ap.add_argument("--method", default="rcb", choices=["rcb", "slab"])
    args = ap.parse_args()
    pm = os.path.join(args.case, "polyMesh")
    g = M.read_polymesh(pm)
    b = M._foam_body(open(os.path.join(pm, "points")).read())
    points = np.array(re.findall(r"\(\s*([-+0-9.eE]+)\s+([-+0-9.eE]+)\s+([-+0-9.eE]+)\s*\)", b), dtype=np.float64)
    b = M._foam_body(open(os.path.join(pm, "faces")).read())
    faces = [np.array(m.split(), dtype=np.int64) for m in re.findall(r"\d+\(([\d\s]+)\)", b)]
    patches = []
    for line in open(os.path.join(pm, "boundary")):
        t = line.split()
        if line.startswith("#") or len(t) < 3:
            continue
        patches.append((t[0], int(t[1]), int(t[2])))
    rank = M.rcb_ranks(g, args.nranks) if args.method == "rcb" else M.slab_ranks(g.numCells, args.nranks)
    parts = M.write_decomposition(args.case, points, faces, g, rank, args.nranks, patches)
    for r, p in enumerate(parts):
        print(f"processor{r}: {p.numCells} cells, {p.numInnerFaces} inner faces, {p.npro} processor faces to ranks "
              f"{list(map(int, p.neighbProcNo))}")


if __name__ == "__main__":
    main()
