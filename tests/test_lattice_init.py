"""CPU: hybird_b200.lattice_init (the box-domain restatement of LB::latticeBolzmannInit used by
bench / smoke / GPU tests) against the initial state the unmodified reference produced (golden)."""
import os
import sys

import numpy as np
import pytest

import common
import golden_util as gu
from hybird_b200 import lattice_init as li

sys.path.insert(0, os.path.join(common.ROOT, "oracle"))
import cases  # noqa: E402

# multi-sphere elements need the DEM's generateParticles; the DRUM geometry (cylinder) is set up by the reference's host code
NAMES = [n for n in gu.names() if not n.startswith("drum")]


@pytest.mark.parametrize("name", NAMES)
def test_init_state_matches_reference(name):
    g = gu.Golden(name)
    case = cases.catalogue()[name]
    prm = li.params_from_case(case)
    tr = common.KinematicTrace(case, prm)
    st = li.build_state(case, tr.initial_particles() if len(tr.parts) else None)
    for k in ("size", "boundary"):
        assert list(st.params[k]) == list(g.params[k])
    for k in ("lbF",):
        assert np.array_equal(np.array(st.params[k]), np.array(g.params[k])), k
    for k in ("initDynVisc", "plasticVisc", "yieldStress", "turbConst", "slipCoefficient"):
        assert st.params[k] == g.params[k], k
    assert st.params["nWalls"] == g.params["nWalls"]
    tf, si, n, u, mass, visc = g.init_arrays()
    assert np.array_equal(st.type_flags, tf), "type/p/node flags"
    sel = ((tf & 0x10) != 0) | ((tf & 0x0F) >= 5)
    assert np.array_equal(st.solidIndex[sel], si[sel])
    node = (tf & 0x20) != 0
    for a, b, nm in ((st.n, n, "n"), (st.u, u, "u"), (st.mass, mass, "mass"), (st.visc, visc, "visc")):
        assert np.array_equal(a[node], b[node]), nm


def test_neighbor_table_periodic_wrap():
    """per-axis wrap, diagonals double-wrap, shell cells self-link, d[0]=0."""
    size, bnd = (24, 20, 16), [4, 4, 4, 4, 7, 7]
    nb = li.neighbor_table(size, bnd)
    X, Y, Z = size
    idx = lambda x, y, z: x + X * (y + Y * z)
    c = idx(1, 1, 1)
    assert nb[2][c] == idx(22, 1, 1)
    assert nb[8][c] == idx(22, 18, 1)
    assert nb[16][c] == idx(22, 1, 0)
    assert nb[0][c] == 0
    s = idx(0, 5, 5)
    assert all(nb[j][s] == s for j in range(19))


@pytest.mark.parametrize("bnd", [[4, 4, 4, 4, 7, 7], [4, 4, 4, 4, 4, 4], [7, 7, 4, 4, 8, 7], [7, 5, 7, 7, 4, 4]])
@pytest.mark.parametrize("window", [None, (0, 6), (3, 9), (5, 11)])
def test_stencil_views_equal_the_neighbor_table(bnd, window):
    """The shifted-view stencil used by build_state == gathers / scatters through the explicit neighbour table."""
    size = (9, 7, 11)
    planes = None if window is None else li.window_planes(size[2], bnd[4] == 4, window[0], window[1])
    nb = li.Neighbors(size, bnd, planes)
    st3 = li.Stencil(nb)
    rng = np.random.default_rng(7)
    n = nb.idx.size
    a = rng.integers(0, 1000, n)
    mask = rng.random(n) < 0.3
    for j in range(19):
        if j:
            assert np.array_equal(st3.gather(a, j), a[nb[j]]), j
        want = np.zeros(n, dtype=bool); want[nb[j][mask]] = True
        got = np.zeros(n, dtype=bool); st3.scatter_or(got, mask, j)
        if window is None:
            assert np.array_equal(got, want), j
        else:  # links leaving the stored planes are only defined on the (discarded) margin: compare the window's cells
            keep = (nb.z >= window[0]) & (nb.z < window[1])
            assert np.array_equal(got[keep], want[keep]), j
