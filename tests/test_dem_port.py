"""CPU: the Python restatement of DEM::discreteElementStep (oracle/dem_port.py, the checker of the device-side DEM) against
the particle traces the unmodified reference recorded (tests/golden/{spheres_dem,spheres_hertz,bed_dem}.npz): fed with the
reference's own hydrodynamic forces it must follow the reference's elements through contacts, wall contacts and rebuilds of
the neighbour table to rounding."""
import numpy as np
import pytest

import golden_util as gu

DEM_CASES = ("spheres_dem", "spheres_hertz", "bed_dem", "cfg3_mini")
CLUSTER_CASES = ("cluster_dem", "clusters_hit")
PBC_CASES = ("spheres_pbc_dem", "cfg5_mini_dem")


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


@pytest.mark.parametrize("name", DEM_CASES + CLUSTER_CASES + PBC_CASES)
def test_host_mirror_of_the_dem_initialisation(name):
    """hybird_b200.dem_init restates what DEM::discreteElementGet / discreteElementInit derive (material constants, sub-step,
    neighbour-table range, masses, inertias, walls): number by number what the unmodified reference held after its init."""
    import cases
    from hybird_b200 import dem_init
    g = gu.Golden(name)
    ref = g.dem()
    mine = dem_init.dem_from_case(cases.catalogue()[name])
    assert set(mine["params"]) == set(ref["params"])
    for k, v in ref["params"].items():
        assert mine["params"][k] == v, k
    assert len(mine["elmts"]) == len(ref["elmts"]) and len(mine["walls"]) == len(ref["walls"])
    for a, b in zip(mine["elmts"], ref["elmts"]):
        for k in ("size", "radius", "m", "I", "x0", "x1", "w0"):
            assert a[k] == b[k], k
    for a, b in zip(mine["walls"], ref["walls"]):
        for k in ("n", "p", "vel", "omega", "rotCenter", "moving"):
            assert a[k] == b[k], k
    assert mine["pbcs"] == ref["pbcs"]


def test_dem_init_refuses_what_the_device_does_not_cover():
    import cases
    from hybird_b200 import dem_init
    case = dict(cases.catalogue()["cluster_dem"], boundary0=4, boundary1=4)  # clusters with periodic boundaries
    with pytest.raises(ValueError):
        dem_init.dem_from_case(case)


def test_lb_mirror_refuses_unknown_shapes_and_periodic_dem_before_touching_the_device():
    """LB.demInit checks what the device-side DEM covers on the host side (no CUDA call is made for a refused set-up)."""
    from hybird_b200 import LB
    g = gu.Golden("cluster_dem")
    lb = LB(dict(g.params))
    dem = gu.Golden("spheres_dem").dem()
    dem["elmts"][0]["size"] = 5  # DEM::compositeProperties knows elements of one to four spheres
    with pytest.raises(ValueError, match="clusters of 2-4"):
        lb.demInit(dem)
    dem = gu.Golden("spheres_pbc_dem").dem()
    assert len(dem["pbcs"]) == 2
    dem["elmts"][1]["size"] = 2  # periodic DEM boundaries are covered for single spheres only
    with pytest.raises(ValueError, match="periodic DEM boundaries"):
        lb.demInit(dem)


@pytest.mark.parametrize("name", ("cluster_dem", "clusters_hit"))
def test_cluster_port_follows_reference_trace(name):
    """The general restatement (elements of 2-4 spheres: quaternions, lever arms, per-particle tables) against the reference's
    recorded particles -- positions and lever arms of every sphere, velocity and spin of every element, every LB step."""
    import dem_port
    g = gu.Golden(name)
    dem = g.dem()
    P = dem_port.DemPortClusters(dem)
    n = len(dem["elmts"])
    F = np.zeros((n, 3)); M = np.zeros((n, 3))
    worst = 0.0
    for s in range(g.steps):
        parts, elmts, comps, flag = g.trace[s]
        px0, prv, x1, wg = P.step(F, M)
        assert len(px0) == len(parts)
        worst = max(worst, np.abs(px0 - parts["x0"]).max(), np.abs(prv - parts["radiusVec"]).max(), np.abs(x1 - elmts["x1"]).max(),
                    np.abs(wg - elmts["wGlobal"]).max())
        F, M = g.forces[s][0], g.forces[s][1]
    assert worst <= 1e-12, worst


@pytest.mark.parametrize("name", ["spheres_pbc_dem", "cfg5_mini_dem"])
def test_periodic_port_follows_reference_trace(name):
    """Periodic DEM boundaries (pbcShift, ghost particles incl. the corner ghost, table over particles and ghosts, contacts
    across a periodic face): the particle list the LB side receives -- positions of particles and ghosts, their elements, the
    elements' component lists, the rescan flag -- against the reference's recording, every LB step."""
    import dem_port
    g = gu.Golden(name)
    dem = g.dem()
    P = dem_port.DemPortPbc(dem)
    n = len(dem["elmts"])
    F = np.zeros((n, 3)); M = np.zeros((n, 3))
    worst = 0.0
    counts = set()
    for s in range(g.steps):
        parts, elmts, comps, flag = g.trace[s]
        x0, ci, cl, x1, wg, new_list = P.step(F, M)
        assert len(x0) == len(parts) and np.array_equal(ci, parts["clusterIndex"])
        assert list(comps) == [p for c in cl for p in c] and bool(flag) == new_list
        worst = max(worst, np.abs(x0 - parts["x0"]).max(), np.abs(x1 - elmts["x1"]).max(), np.abs(wg - elmts["wGlobal"]).max())
        counts.add(len(parts))
        F, M = g.forces[s][0], g.forces[s][1]
    assert worst <= 1e-12, worst
    if name == "spheres_pbc_dem":
        assert P.rebuilds >= 4 and len(counts) >= 2, (P.rebuilds, counts)
