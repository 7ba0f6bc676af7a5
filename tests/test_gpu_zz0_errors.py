"""Error behaviour of the C ABI (return codes + fc_last_error, never abort, never a device fault):
a mesh whose index arrays would send the kernels out of bounds is refused by fc_set_mesh with FC_ERR_ARG, and the
context stays usable."""
import copy

import numpy as np
import pytest

from freecappuccino_b200 import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fc():
    from freecappuccino_b200 import lib
    return lib


def corrupt(m, what):
    m = copy.copy(m)
    if what == "owner-high":
        m.owner = m.owner.copy(); m.owner[m.numFaces - 1] = m.numCells + 1
    elif what == "owner-zero":
        m.owner = m.owner.copy(); m.owner[3] = 0
    elif what == "neighbour-high":
        m.neighbour = m.neighbour.copy(); m.neighbour[5] = m.numCells + 7
    elif what == "neighbour-is-owner":
        m.neighbour = m.neighbour.copy(); m.neighbour[2] = m.owner[2]
    return m


@pytest.mark.parametrize("what", ["owner-high", "owner-zero", "neighbour-high", "neighbour-is-owner"])
def test_set_mesh_refuses_out_of_range_indices(fc, what):
    good = cases.hex_case(5, 4, 3)
    ctx = fc.Context(0)
    with pytest.raises(fc.FcError) as e:
        ctx.set_mesh(corrupt(good, what))
    assert e.value.code == fc.FC_ERR_ARG and "fc_set_mesh" in str(e.value)
    # the context is still usable
    ctx.set_mesh(good)
    ioffset, ja, diag, icj, jci = ctx.create_csr()
    assert ioffset[-1] == good.nnz + 1
    ctx.close()


def tridiag(n):
    """1-based CSR of the 1-D Laplacian."""
    rows = []
    for i in range(n):
        rows.append([j for j in (i - 1, i, i + 1) if 0 <= j < n])
    ioffset = np.cumsum([1] + [len(r) for r in rows]).astype(np.int32)
    ja = np.array([j + 1 for r in rows for j in r], np.int32)
    diag = np.array([ioffset[i] + r.index(i) for i, r in enumerate(rows)], np.int32)
    a = np.array([2.0 if j == i else -1.0 for i, r in enumerate(rows) for j in r])
    return ioffset, ja, diag, a


@pytest.mark.parametrize("what", ["ja-high", "ja-zero", "diag-off", "ioffset-end", "ioffset-decreasing"])
def test_solve_csr_refuses_a_broken_pattern(fc, what):
    n = 50
    ioffset, ja, diag, a = tridiag(n)
    b, x = np.ones(n), np.zeros(n)
    opts = fc.solver_opts(1e-10, 500)
    ctx = fc.Context(0)
    bad = [ioffset.copy(), ja.copy(), diag.copy()]
    if what == "ja-high":
        bad[1][7] = n + 1
    elif what == "ja-zero":
        bad[1][0] = 0
    elif what == "diag-off":
        bad[2][10] += 1
    elif what == "ioffset-end":
        bad[0][-1] += 1
    elif what == "ioffset-decreasing":
        bad[0][5], bad[0][6] = bad[0][6], bad[0][5]
    with pytest.raises(fc.FcError) as e:
        ctx.solve_csr("iccg", bad[0], bad[1], bad[2], a, b, x, opts)
    assert e.value.code == fc.FC_ERR_ARG and "fc_solve_csr" in str(e.value)
    rep = ctx.solve_csr("iccg", ioffset, ja, diag, a, b, x, opts)     # the context still works
    assert rep.resl / rep.res0 < 1e-10
    assert np.abs(x[1:-1] * 2 - x[:-2] - x[2:] - 1).max() < 1e-8
    ctx.close()
