"""GPU: checkpoint / restart through the C ABI (lbGpuSaveState / lbGpuLoadState).  The reference has no fluid
restart, so the property checked is the one that matters: a run continued from a saved state is
BIT-IDENTICAL to the uninterrupted run, for every kind of lattice (pure fluid, viscosity state, moving walls,
free surface, particles, curved walls)."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def _engine(g):
    from hybird_b200 import LB
    lb = LB(dict(g.params))
    lb.latticeBolzmannInit(*g.init_arrays())
    return g.configure(lb)


def _advance(g, lb, first, last):
    out = None
    for s in range(first, last + 1):
        parts, elmts, comps, flag = g.trace[s - 1]
        if g.params["freeSurface"]:
            lb.latticeBoltzmannFreeSurfaceStep()
        lb.latticeBoltzmannCouplingStep(flag, elmts, parts, comps)
        out = lb.latticeBolzmannStep(elmts, parts)
    return out


@pytest.mark.parametrize("name", ["cfg2_mini", "bingham_smago", "slip_dyn", "cfg4_mini", "cfg5_mini", "cluster_dem", "drum_bingham"])
def test_restart_is_bit_identical(name):
    g = gu.Golden(name)
    half, end = g.steps // 3, min(g.steps, 2 * (g.steps // 3) + 5)
    a = _engine(g)
    _advance(g, a, 1, half)
    blob = a.save_state()
    Fa = _advance(g, a, half + 1, end)
    sa = a.fetch()
    a.close()
    # a fresh handle of the same lattice; its initial fields are overwritten by the state
    b = _engine(g)
    b.load_state(blob)
    Fb = _advance(g, b, half + 1, end)
    sb = b.fetch()
    b.close()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    for x, y in zip(Fa, Fb):
        assert np.array_equal(x, y)


def test_state_of_another_lattice_is_refused():
    from hybird_b200 import LbGpuError
    a, b = _engine(gu.Golden("cfg2_mini")), _engine(gu.Golden("periodic_all"))
    with pytest.raises(LbGpuError):
        b.load_state(a.save_state())
    with pytest.raises(LbGpuError):
        b.load_state(np.zeros(64, dtype=np.uint8))
    a.close(); b.close()
