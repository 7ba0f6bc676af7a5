"""Host-side plumbing of the cell-partitioned (multi-GPU) build.

One process per GPU, exactly like one MPI rank of ``src-parallel``.  ``torch.distributed`` is
only the bootstrap (it carries the 128-byte NCCL id, where the Fortran MPI build would use
``MPI_Bcast``); the data path -- halo exchange and scalar all-reduces -- is NCCL inside
``libfcapp_cuda.so`` (``csrc/fc_comm.cu``).

``exchange_host`` / ``global_sum_host`` restate ``exchange.f90`` / ``global_sum_mpi.f90`` on
``torch.distributed`` point-to-point calls; they drive the CPU (gloo, world_size 2) tests of the
partition arrays that the NCCL path consumes unchanged.
"""
from __future__ import annotations

import numpy as np


def init_comm(ctx):
    """Create the context's NCCL communicator over the torch.distributed world."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        ctx.comm_init(0, 1, b"\0" * 128)
        return 0, 1
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(rank, world, box[0])
    return rank, world


def enable_p2p(ctx) -> bool:
    """Switch the Krylov loop's halo exchange and reductions to direct NVLink stores between the ranks'
    kernels (csrc/fc_p2p.cu).  Call after ``create_csr``.  The blobs (CUDA IPC handle + connection table)
    are all-gathered in rank order -- MPI_Allgather in a Fortran MPI host.  Set FC_NO_P2P=1 to stay on NCCL."""
    import os
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1 or os.environ.get("FC_NO_P2P") == "1":
        return False
    blobs = [None] * dist.get_world_size()
    dist.all_gather_object(blobs, ctx.p2p_blob())
    ctx.p2p_open(blobs)
    return True


def bufind(mesh) -> np.ndarray:
    """0-based owner cell of every processor face (``bufind(i) = owner(iProcFacesStart+i)``,
    src-parallel/mesh_geometry_and_topology.f90:879-881)."""
    s = mesh.iProcFacesStart
    return mesh.owner[s:s + mesh.npro].astype(np.int64) - 1


def exchange_host(mesh, phi: np.ndarray, group=None) -> None:
    """``call exchange(phi)`` (src-parallel/exchange.f90:3-92) with torch.distributed send/recv:
    pack ``buffer(i) = phi(bufind(i))``, swap the per-neighbour slices, unpack into
    ``phi(iProcStart+i)``."""
    import torch
    import torch.distributed as dist
    if mesh.npro == 0:
        return
    buf = torch.from_numpy(np.ascontiguousarray(phi[bufind(mesh)]))
    rcv = torch.empty_like(buf)
    ops = []
    off = mesh.neighbProcOffset.astype(np.int64) - 1
    for c, q in enumerate(mesh.neighbProcNo):
        s, e = int(off[c]), int(off[c + 1])
        ops.append(dist.P2POp(dist.isend, buf[s:e], int(q), group))
        ops.append(dist.P2POp(dist.irecv, rcv[s:e], int(q), group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    phi[mesh.numCells:mesh.numCells + mesh.npro] = rcv.numpy()


def global_sum_host(x: float, group=None) -> float:
    """``call global_sum(x)`` (src-parallel/global_sum_mpi.f90:4-37)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())
