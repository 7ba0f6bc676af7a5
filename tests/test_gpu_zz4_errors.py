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
