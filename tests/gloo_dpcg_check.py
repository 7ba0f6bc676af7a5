"""world_size-2 gloo job (launched by tests/test_partition.py): every rank owns one slab of a Poisson
problem and runs src-parallel's DPCG (src-parallel/dpcg.f90) with the halo exchange and global sums of
freecappuccino_b200.parallel; the result must equal the lock-step multi-rank oracle bit for bit (the
all-reduce of two ranks is a single commutative addition, so even the reduction order agrees)."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freecappuccino_b200 import cases, mesh as M, parallel  # noqa: E402
from oracle import oracle as O, oracle_par as OP  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = cases.hex_case(9, 7, 10, kinds=("wall", "wall", "wall", "wall", "symmetry", "symmetry"))
    parts = M.partition(g, M.slab_ranks(g.numCells, world), world)
    su_g = cases.poisson_rhs(g)
    # lock-step oracle on every rank (cheap): assembles a / apr per rank and gives the expected answer
    pc = OP.ParCase(parts)
    mus, phis = [], []
    for m, fl in zip(parts, pc.fields):
        fl.su[:] = su_g[m.cell_global]
        mus.append(-np.ones(m.numCells + m.npro))
        phis.append(np.zeros(m.numTotal))
    pc.laplacian(mus, phis)
    a, apr, su = pc.fields[rank].a.copy(), pc.apr[rank].copy(), pc.fields[rank].su.copy()
    xs = [np.zeros(m.numTotal) for m in parts]
    rep, hist = pc.solve("dpcg", xs, 1e-8, 500, history=True)

    # distributed DPCG of this rank: local pieces via the single-rank oracle SpMV, comms via torch.distributed
    m, csr = parts[rank], pc.csr[rank]
    n = m.numCells
    own = parallel.bufind(m)
    small = O.SMALL
    fi = np.zeros(m.numTotal)
    pk = np.zeros(n + m.npro)
    adiag = a[csr.diag - 1]

    def matvec(x):
        y = O.spmv(csr, a, x)
        for i in range(m.npro):                     # processor strip, ascending i (src-parallel/dpcg.f90:132-136)
            y[own[i]] = y[own[i]] + apr[i] * x[n + i]
        return y

    res = su - matvec(fi)
    # sequential sums like the Fortran intrinsics the oracle restates
    seq = lambda v: float(np.add.reduce(v)) if False else float(sum_seq(v))
    res0 = parallel.global_sum_host(sum_seq(np.abs(res)))
    s0, it, mine = float(np.float32(1e20)), 0, []
    for l in range(500):
        zk = res / (adiag + small)
        sk = parallel.global_sum_host(sum_seq(res * zk))
        bet = sk / s0
        pk[:n] = zk + bet * pk[:n]
        parallel.exchange_host(m, pk)
        zk = matvec(pk)
        pkapk = parallel.global_sum_host(sum_seq(pk[:n] * zk))
        alf = sk / pkapk
        fi[:n] = fi[:n] + alf * pk[:n]
        res = res - alf * zk
        resl = parallel.global_sum_host(sum_seq(np.abs(res)))
        s0 = sk
        it += 1
        mine.append(resl)
        if resl / (res0 + small) < 1e-8:
            break
    parallel.exchange_host(m, fi)
    ok = (it == rep.iters and np.array_equal(np.array(mine), hist) and np.array_equal(fi[:n + m.npro], xs[rank][:n + m.npro]))
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if rank == 0:
        print("GLOO DPCG OK" if all(flags) else f"GLOO DPCG MISMATCH it={it} oracle={rep.iters}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


def sum_seq(v):
    s = 0.0
    for x in v.tolist():
        s = s + x
    return s


if __name__ == "__main__":
    main()
