"""GPU (B200): north_star's success criterion on the configurations themselves, against hash-only fixtures produced
by the UNMODIFIED reference (tests/golden/make_golden_full.py):

  <cfg>_full   cfg1 100x150x30, cfg2 256^3, cfg3 128x128x256, cfg4 512x128x256 at FULL size for 1000 steps: the
               cell-type map after every one of the first 100 steps and after step 1000 (sha256), every field at steps
               1, 100, 1000 -- sha256 over the active cells (bit-exact, required without a free surface) and a regular
               subsample of ~25 000 cells within 1e-12 @ 1 / 1e-9 @ 1000, element and wall forces of every step.
  <mini>_long  the free-surface minis for 1000 steps, complete final arrays.

The engine is driven through the C ABI with the reference's recorded particle inputs."""
import numpy as np
import pytest

import common
import golden_full_util as gfu

pytestmark = pytest.mark.gpu


def _gpu(g):
    from hybird_b200 import LB
    lb = LB(dict(g.params))
    lb.latticeBolzmannInit(*g.init_arrays())
    return g.configure(lb)


def _types(lb):
    return lb.fetch(("type_flags",))["type_flags"]


@pytest.mark.parametrize("name", gfu.names("long"))
def test_gpu_free_surface_minis_for_1000_steps(name):
    g = gfu.GoldenFull(name)
    lb = _gpu(g)
    rep = gfu.check_run(g, lb, _types, common.gpu_state)
    lb.close()
    assert g.steps == 1000 and 1000 in rep


@pytest.mark.parametrize("name", gfu.names("full"))
def test_gpu_full_size_configuration_for_1000_steps(name):
    g = gfu.GoldenFull(name)
    lb = _gpu(g)
    rep = gfu.check_run(g, lb, _types, common.gpu_state)
    lb.close()
    assert g.steps == 1000 and 1000 in rep
