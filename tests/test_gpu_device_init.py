"""GPU: lbGpuInitBox (lattice initialisation on the device) against the state the UNMODIFIED
reference produced (golden init arrays) and against the engine started from those arrays: identical types, wall
indices, densities, velocities, masses, viscosities and populations at step 0, and bit-identical states after stepping."""
import os
import sys

import numpy as np
import pytest

import common
import golden_util as gu
from hybird_b200 import lattice_init as li

sys.path.insert(0, os.path.join(common.ROOT, "oracle"))
import cases  # noqa: E402

pytestmark = pytest.mark.gpu
# multi-sphere elements need the DEM's generateParticles; the DRUM geometry is set up by the reference's host code
NAMES = [n for n in gu.names() if not n.startswith("drum")]


def _pair(name, n_slabs=1):
    from hybird_b200 import LB
    g = gu.Golden(name)
    case = cases.catalogue()[name]
    prm = li.params_from_case(case)
    prm["nWalls"] = g.params["nWalls"]
    if n_slabs > 1:
        prm["nSlabs"] = n_slabs; prm["nLocalSlabs"] = n_slabs
    tr = common.KinematicTrace(case, prm)
    host = LB(dict(g.params, **({"nSlabs": n_slabs, "nLocalSlabs": n_slabs} if n_slabs > 1 else {})))
    host.latticeBolzmannInit(*g.init_arrays())
    dev = LB(prm)
    dev.latticeBolzmannInitBox(case, tr.initial_particles() if len(tr.parts) else None)
    return g, host, dev


@pytest.mark.parametrize("name", NAMES)
def test_device_init_equals_reference_init(name):
    g, host, dev = _pair(name)
    tf, si, n, u, mass, visc = g.init_arrays()
    a, b = dev.fetch(), host.fetch()
    assert np.array_equal(a["type_flags"], tf), "type / p / node flags differ from the reference's initial state"
    sel = ((tf & 0x10) != 0) | ((tf & 0x0F) >= 5)
    assert np.array_equal(a["solidIndex"][sel], si[sel])
    node = (tf & 0x20) != 0
    for k, ref in (("n", n), ("u", u), ("mass", mass), ("visc", visc)):
        assert np.array_equal(a[k][node], ref[node]), k
    assert np.array_equal(a["f"], b["f"])
    # and the two engines stay bit-identical when stepped with the recorded particle inputs
    it = gu.replay(g, host, None)
    for s, F, M, V, W in gu.replay(g, dev, None):
        _, F0, M0, V0, W0 = next(it)
        assert np.array_equal(F, F0) and np.array_equal(W, W0)
        if s == 12:
            break
    a, b = dev.fetch(), host.fetch()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    host.close(); dev.close()


@pytest.mark.parametrize("name", ["cfg4_mini", "cfg5_mini", "couette_dyn"])
def test_device_init_with_slabs(name):
    g, host, dev = _pair(name, n_slabs=3)
    a, b = dev.fetch(), host.fetch()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    host.close(); dev.close()
