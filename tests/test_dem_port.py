"""CPU: the Python restatement of DEM::discreteElementStep (oracle/dem_port.py, the checker of the device-side DEM) against
the particle traces the unmodified reference recorded (tests/golden/{spheres_dem,spheres_hertz,bed_dem}.npz): fed with the
reference's own hydrodynamic forces it must follow the reference's elements through contacts, wall contacts and rebuilds of
the neighbour table to rounding."""
import numpy as np
import pytest

import golden_util as gu

DEM_CASES = ("spheres_dem", "spheres_hertz", "bed_dem")


@pytest.mark.parametrize("name", DEM_CASES)
def test_dem_port_follows_reference_trace(name):
    import dem_port
    g = gu.Golden(name)
    dem = g.dem()
    assert dem is not None and dem["counts"] == dict(pbcs=0, cylinders=0, objects=0, ghosts=0)
    P = dem_port.DemPort(dem)
    n = len(dem["elmts"])
    F = np.zeros((n, 3)); M = np.zeros((n, 3))
    worst = 0.0
    for s in range(g.steps):
        parts, elmts, comps, flag = g.trace[s]
        x0, x1, w = P.step(F, M)
        assert not flag  # dem.newNeighborList is only raised with periodic DEM boundaries (DEM.cpp:1414)
        worst = max(worst, np.abs(x0 - parts["x0"]).max(), np.abs(x1 - elmts["x1"]).max(), np.abs(w - elmts["wGlobal"]).max())
        F, M = g.forces[s][0], g.forces[s][1]
    assert worst <= 1e-12, worst
    if name == "bed_dem":
        assert P.rebuilds >= 4  # the table is rebuilt several times and pairs enter / leave it
