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


def _record(g, rep):
    """Worst errors per dump step, appended to gpurun_out/parity_report.jsonl (copied to profiles/ by hand)."""
    import json
    import os
    out = os.environ.get("LBGPU_PARITY_REPORT", os.path.join(common.ROOT, "gpurun_out", "parity_report.jsonl"))
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "a") as fh:
            fh.write(json.dumps(dict(fixture=g.name, case=g.case_name, lattice=g.params["size"], steps=g.steps,
                                     type_maps_checked=sorted(g.type_sha), report={str(k): v for k, v in rep.items()})) + "\n")
    except OSError:
        pass


@pytest.mark.parametrize("name", gfu.names("long"))
def test_gpu_free_surface_minis_for_1000_steps(name):
    g = gfu.GoldenFull(name)
    lb = _gpu(g)
    rep = gfu.check_run(g, lb, _types, common.gpu_state)
    lb.close()
    _record(g, rep)
    assert g.steps == 1000 and 1000 in rep


@pytest.mark.parametrize("name", gfu.names("full"))
def test_gpu_full_size_configuration_for_1000_steps(name):
    g = gfu.GoldenFull(name)
    lb = _gpu(g)
    rep = gfu.check_run(g, lb, _types, common.gpu_state)
    lb.close()
    _record(g, rep)
    assert g.steps == 1000 and 1000 in rep
