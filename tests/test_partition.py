"""Cell partition (src-parallel layout) and the multi-rank logic, on CPU: the partitioner is pinned to
the reference's own shipped 2-rank decomposition; the lock-step multi-rank oracle reduces to the serial
one for a single rank; a world_size-2 gloo job runs the distributed DPCG over torch.distributed halo
exchange / all-reduce and matches the lock-step oracle."""
import os
import sys

import numpy as np
import pytest

from freecappuccino_b200 import cases, mesh as M
from oracle import oracle as O, oracle_par as OP

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitioner_reproduces_the_reference_decomposition():
    """examples/cavity/cavity-setup-parallel.tar.gz (OpenFOAM decomposePar + the reference's `process` file)."""
    g = cases.golden_mesh(os.path.join(GOLD, "cavity.npz"))
    d = np.load(os.path.join(GOLD, "cavity_par.npz"))
    cell_rank = np.zeros(g.numCells, dtype=np.int64)
    cell_rank[d["p1_cellProc"]] = 1
    parts = M.partition(g, cell_rank, 2)
    for r, p in enumerate(parts):
        assert np.array_equal(p.owner - 1, d[f"p{r}_owner"])
        assert np.array_equal(p.neighbour - 1, d[f"p{r}_neighbour"])
        assert np.array_equal(p.cell_global, d[f"p{r}_cellProc"])
        assert np.array_equal(p.face_global, np.abs(d[f"p{r}_faceProc"]) - 1)
        # `process`: neighbProcNo nfaces startFace ; boundary table: wall 10+30 @370, symmetry 400 @410
        other, nfaces, start = d[f"p{r}_process"][0]
        assert list(p.neighbProcNo) == [other] and p.npro == nfaces and p.iProcFacesStart == start
        assert p.counts == {"wall": 40, "symmetry": 400} and p.starts == {"wall": 370, "symmetry": 410}
        # geometry of the rank mesh recomputed from ITS OWN points/faces equals the partitioned global geometry
        own = cases.golden_mesh(os.path.join(GOLD, "cavity_par.npz"), prefix=f"p{r}_")
        assert np.allclose(own.vol, p.vol[:p.numCells], rtol=1e-12)
        assert np.allclose(own.xc, p.xc[:p.numCells], atol=1e-14)
        nb = p.iProcFacesStart
        assert np.allclose(own.arx[:nb], p.arx[:nb], atol=1e-16) and np.allclose(own.xf[:nb], p.xf[:nb], atol=1e-14)


@pytest.mark.parametrize("nr", [2, 3, 4])
def test_partition_invariants(nr):
    g = cases.skew_case(8, 7, 6)
    parts = M.partition(g, M.slab_ranks(g.numCells, nr), nr)
    assert sum(p.numCells for p in parts) == g.numCells
    assert sum(p.numInnerFaces for p in parts) + sum(p.npro for p in parts) // 2 == g.numInnerFaces
    for r, p in enumerate(parts):
        assert p.numTotal == p.numCells + p.npro + sum(p.counts.values())
        assert p.neighbProcOffset[0] == 1 and p.neighbProcOffset[-1] == p.npro + 1
        for c, q in enumerate(p.neighbProcNo):        # mirrored connection has the same length
            k = list(parts[q].neighbProcNo).index(r)
            assert (p.neighbProcOffset[c + 1] - p.neighbProcOffset[c]
                    == parts[q].neighbProcOffset[k + 1] - parts[q].neighbProcOffset[k])
    # software exchange delivers the remote cell's value into every halo slot
    pc = OP.ParCase(parts)
    phis = []
    for p in parts:
        a = np.full(p.numTotal, -1.0)
        a[:p.numCells] = p.cell_global
        phis.append(a)
    pc.exchange(phis)
    for p, a in zip(parts, phis):
        assert np.array_equal(a[p.numCells:p.numCells + p.npro], p.halo_global.astype(float))


def load_case(g, parts, pc, f, fmi):
    gp = O.grad_gauss(g, f["p"], 1)
    for m, fl in zip(parts, pc.fields):
        for k in ("u", "v", "w", "p", "den"):
            getattr(fl, k)[:] = M.scatter_total(g, m, f[k])
        for k in ("apu", "apv", "apw"):
            getattr(fl, k)[:] = M.scatter_cells(g, m, f[k])
        fl.dPdxi[:] = M.scatter_cells(g, m, gp)
        c = m.count("inlet")
        if c:
            gf = m.face_global[m.faces_start("inlet"):m.faces_start("inlet") + c]
            fl.fmi[:c] = fmi[gf - g.faces_start("inlet")]


@pytest.mark.parametrize("solver", ["dpcg", "iccg", "bicgstab"])
def test_single_rank_parallel_oracle_equals_serial(solver):
    g = cases.skew_case(8, 7, 6)
    f = cases.flow_fields(g)
    fmi, flomas = cases.inlet_fluxes(g, f)
    csr = O.create_csr(g)
    of = O.Fields(g, csr.nnz)
    for k in ("u", "v", "w", "p", "den", "apu", "apv", "apw"):
        getattr(of, k)[:] = f[k]
    of.fmi[:fmi.size] = fmi
    of.dPdxi[:] = O.grad_gauss(g, of.p, 1)
    kw = dict(solver=solver, flomas=flomas, sor=1e-8, nsw=500, npcor=2, lsq_flag=True)
    oo = O.calcp_opts(**kw)
    oo.sol.parallel = 1          # same arithmetic (+small in the preconditioners) on both sides
    # serial ppref = pp(pRefCell); parallel ppref = mean(pp): compare quantities that do not depend on it
    rep_s = O.calcp(g, csr, of, oo)
    parts = M.partition(g, np.zeros(g.numCells, dtype=np.int64), 1)
    pc = OP.ParCase(parts)
    load_case(g, parts, pc, f, fmi)
    rep_p = pc.calcp(oo)
    assert [rep_p.rep[k].iters for k in range(2)] == [rep_s.rep[k].iters for k in range(2)]
    fl = pc.fields[0]
    assert np.array_equal(fl.a, of.a) and np.array_equal(fl.flmass, of.flmass)
    assert np.array_equal(fl.u, of.u) and np.array_equal(fl.pp, of.pp)
    n = g.numCells   # p differs by the constant urf*(ppref_serial - ppref_parallel) on the cells
    assert np.allclose(fl.p[:n] - fl.p[:n].mean(), of.p[:n] - of.p[:n].mean(), atol=1e-12)


def test_two_rank_dpcg_is_partition_independent():
    """Jacobi-PCG does not depend on the partition (only on the reduction order): same iteration count
    and solution as the serial solver on the same Poisson system (SURVEY 8e parity caveat 1)."""
    g = cases.hex_case(10, 8, 12, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry"))
    csr = O.create_csr(g)
    su = cases.poisson_rhs(g)
    su_s = su.copy()
    a = O.laplacian(g, csr, -np.ones(g.numCells), np.zeros(g.numTotal), su_s)
    x = np.zeros(g.numTotal)
    _, _, iters, _ = O.solve("dpcg", csr, a, su_s, x, sor=1e-8, nsw=1000, parallel=True)
    parts = M.partition(g, M.slab_ranks(g.numCells, 2), 2)
    pc = OP.ParCase(parts)
    mus, phis = [], []
    for m, fl in zip(parts, pc.fields):
        fl.su[:] = su[m.cell_global]
        mus.append(-np.ones(m.numCells + m.npro))
        phis.append(np.zeros(m.numTotal))
    pc.laplacian(mus, phis)
    xs = [np.zeros(m.numTotal) for m in parts]
    rep = pc.solve("dpcg", xs, 1e-8, 1000)
    assert abs(rep.iters - iters) <= 1
    xg = M.gather_cells(g, parts, xs)
    assert cases.rel_l2(xg, x[:g.numCells]) < 1e-7


def test_gloo_two_rank_distributed_dpcg():
    """world_size 2 over gloo: the host-side exchange / global_sum (freecappuccino_b200.parallel) on the
    partition arrays the NCCL path consumes, driving a distributed DPCG; matches the lock-step oracle."""
    import subprocess
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "gloo_dpcg_check.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "GLOO DPCG OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_noise_bars_of_the_multi_rank_check():
    """tests/mgpu_check.py holds BiCGStab to a bar measured with the oracle itself (one-ulp noise on the input
    velocities): on 2 ranks the reference algorithm is stable (bar = the north star's 1e-10), on the 8-rank partition
    of the same kind of mesh one-ulp noise moves its result by more than 1e-9 -- with unchanged iteration counts."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mgpu_check_helpers", os.path.join(ROOT, "tests", "mgpu_check.py"))
    src = open(spec.origin).read().split("def main")[0]
    ns = {"__file__": spec.origin, "__name__": "mgpu_check_helpers"}
    exec(compile(src, spec.origin, "exec"), ns)
    from oracle import oracle as O, oracle_par as OP
    for world, stable in ((2, True), (8, False)):
        g = cases.hex_case(12, 9, 4 * world, kinds=("inlet", "outlet", "wall", "symmetry", "wall", "wall"))
        f = cases.flow_fields(g)
        fmi, flomas = cases.inlet_fluxes(g, f)
        gp = O.grad_gauss(g, f["p"], 1)
        parts = M.partition(g, M.slab_ranks(g.numCells, world), world)
        oo = O.calcp_opts(solver="bicgstab", flomas=flomas, npcor=1, lsq_flag=False, nigrad=1, sor=1e-7, nsw=2000)
        oo.sol.parallel = 1
        pc = OP.ParCase(parts)
        for m, fl in zip(parts, pc.fields):
            sc = ns["scatter_case"](g, m, f, fmi, gp)
            for k in ("u", "v", "w", "p", "den", "apu", "apv", "apw", "dPdxi"):
                getattr(fl, k)[:] = sc[k]
            fl.fmi[:sc["fmi"].size] = sc["fmi"]
        rep = pc.calcp(oo)
        assert rep.rep[0].iters < 200
        bars = ns["noise_bars"](parts, g, f, fmi, gp, oo, pc, seeds=3)
        assert set(bars) == {"u", "v", "w", "p", "pp", "flmass"} and min(bars.values()) >= 1e-10
        if stable:
            assert max(bars.values()) == 1e-10, bars
        else:
            assert max(bars.values()) > 1e-9, bars
