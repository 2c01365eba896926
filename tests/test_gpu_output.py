"""GPU: the output path.  lbGpuFluidSummary -- IO's screen-export reductions (IO.cpp:835-895,
969-999) on the device -- and lbGpuWriteVti -- IO::exportParaviewFluidOld's file (IO.cpp:698-831) with raw appended
data -- against the same quantities computed on the host from a full lbGpuFetchFields, as the reference's IO does."""
import numpy as np
import pytest

import golden_util as gu
from test_gpu_shim import read_vti

pytestmark = pytest.mark.gpu


def _run(name, steps):
    from hybird_b200 import LB
    g = gu.Golden(name)
    lb = LB(dict(g.params))
    lb.latticeBolzmannInit(*g.init_arrays())
    g.configure(lb)
    for s, *_ in gu.replay(g, lb, None):
        if s == steps:
            break
    return g, lb


@pytest.mark.parametrize("name", ["cfg2_mini", "cfg4_mini", "cfg5_mini", "bingham_smago", "cfg1_mini", "periodic_all"])
def test_summary_equals_the_host_walk_over_all_cells(name):
    g, lb = _run(name, 12)
    s = lb.summary()
    d = lb.fetch(("type_flags", "u", "mass", "visc"))
    t = d["type_flags"]
    act = np.isin(t & 0x0F, (0, 3))
    assert s["active"] == int(act.sum())
    u2 = (d["u"][act] ** 2)
    norm2 = u2[:, 0] + u2[:, 1] + u2[:, 2]              # tinyVector::norm2 (vector.cpp:143-145)
    assert s["max_speed"] == float(np.sqrt(norm2.max()))  # a maximum: exact
    mass = float(d["mass"][act & ((t & 0x10) == 0)].sum())
    assert abs(s["mass"] - mass) <= 1e-12 * abs(mass)     # a sum in another order
    max_visc = (1.8 - 0.5) / 3 / 1.0
    plastic = int(np.count_nonzero(d["visc"][act] > 0.95 * max_visc))
    assert abs(s["plastic_pct"] - 100.0 * plastic / int(act.sum())) <= 1e-12
    lb.close()


@pytest.mark.parametrize("name", ["cfg4_mini", "cfg5_mini", "cfg2_mini"])
def test_vti_file_holds_what_the_reference_writer_would_print(name, tmp_path):
    g, lb = _run(name, 9)
    path = tmp_path / "fluid.vti"
    lb.write_vti(path, dem_solve=True)
    meta, order, arrs = read_vti(path)
    prm = g.params
    X, Y, Z = prm["size"]
    assert meta["extent"] == "0 %d 0 %d 0 %d" % (X - 1, Y - 1, Z - 1)
    want = [("type", "Int8", 1), ("v", "Float64", 3), ("pressure", "Float64", 1)]
    if prm["nonNewtonian"]:
        want.append(("dynVisc", "Float64", 1))
    if prm["freeSurface"]:
        want.append(("AAAmass", "Float64", 1))
    want.append(("solidIndex", "Int16", 1))
    assert order == want
    d = lb.fetch(("type_flags", "solidIndex", "n", "u", "mass", "visc"))
    t = d["type_flags"]
    L, T, D = prm["unitLength"], prm["unitTime"], prm["unitDensity"]
    assert np.array_equal(arrs["type"], np.where(t & 0x10, 1, t & 0x0F))
    assert np.array_equal(arrs["v"], (d["u"] * (L / T)).reshape(-1))
    node = (t & 0x20) != 0
    pres = np.where(node & (d["n"] != 0.0), 0.3333333 * (d["n"] - 1.0) * (D * L * L / T / T), d["n"])
    assert np.array_equal(arrs["pressure"], pres)
    if prm["nonNewtonian"]:
        assert np.array_equal(arrs["dynVisc"], d["visc"] * (D * L * L / T))
    if prm["freeSurface"]:
        assert np.array_equal(arrs["AAAmass"], d["mass"] * D)
    assert np.array_equal(arrs["solidIndex"], d["solidIndex"].astype(np.int16).astype(np.float64))
    lb.close()
