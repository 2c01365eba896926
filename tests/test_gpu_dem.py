"""GPU (B200): DEM::discreteElementStep on the device (lbGpuDem*) against the unmodified reference.

(1) the DEM sub-steps alone, fed with the reference's recorded hydrodynamic forces: positions, velocities and spins of every
    element after every LB step within 1e-12 of the reference's trace (and of the Python restatement);
(2) the whole coupled cycle on the device (DEM step -> coupling step -> LB step, no host round trip for the particles):
    identical cell-type / particle-flag maps at every step, forces within 1e-9, trajectories within 1e-9;
(3) lbGpuRunDem(count) == count x lbGpuRunDem(1); checkpoint / restart carries the elements."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
DEM_CASES = ("spheres_dem", "spheres_hertz", "bed_dem", "cfg3_mini", "cluster_dem", "clusters_hit")  # the last two: elements of 2-4 spheres
TOL_DEM = 1e-12
TOL_COUPLED = 1e-9


def _gpu(g):
    from hybird_b200 import LB
    lb = LB(dict(g.params))
    lb.latticeBolzmannInit(*g.init_arrays())
    return lb.demInit(g.dem())


def _worst(lb, parts, elmts):
    """Largest deviation of the device's particles / elements from the reference's recorded ones."""
    st, pt = lb.demState(), lb.demParticles()
    assert len(pt["x0"]) == len(parts) and np.array_equal(pt["clusterIndex"], parts["clusterIndex"])
    return max(np.abs(pt["x0"] - parts["x0"]).max(), np.abs(pt["radiusVec"] - parts["radiusVec"]).max(),
               np.abs(st["x1"] - elmts["x1"]).max(), np.abs(st["w0"] - elmts["wGlobal"]).max())


@pytest.mark.parametrize("name", DEM_CASES)
def test_device_dem_follows_reference_trace(name):
    import dem_port
    g = gu.Golden(name)
    lb = _gpu(g)
    clusters = any(e["size"] > 1 for e in g.dem()["elmts"])
    P = (dem_port.DemPortClusters if clusters else dem_port.DemPort)(g.dem())
    n = len(g.dem()["elmts"])
    hydro = np.zeros((n, 7))
    worst = worst_port = 0.0
    for s in range(g.steps):
        parts, elmts, comps, flag = g.trace[s]
        lb.demStep(hydro)
        st = lb.demState()
        out = P.step(hydro[:, 0:3], hydro[:, 3:6])
        worst = max(worst, _worst(lb, parts, elmts))
        px1, pw = (out[2], out[3]) if clusters else (out[1], out[2])
        worst_port = max(worst_port, np.abs(st["x1"] - px1).max(), np.abs(st["w0"] - pw).max())
        hydro[:, 0:3], hydro[:, 3:6] = g.forces[s][0], g.forces[s][1]
    assert st["rebuilds"] == P.rebuilds
    assert worst <= TOL_DEM and worst_port <= TOL_DEM, (worst, worst_port)
    lb.close()


@pytest.mark.parametrize("name", DEM_CASES)
def test_coupled_cycle_on_device_matches_reference(name):
    g = gu.Golden(name)
    lb = _gpu(g)
    worst_x = worst_f = 0.0
    for s in range(1, g.steps + 1):
        parts, elmts, comps, flag = g.trace[s - 1]
        lb.runDem(1)
        worst_x = max(worst_x, _worst(lb, parts, elmts))
        t = lb.fetch(("type_flags",))["type_flags"]
        assert np.array_equal(t & 0x1F, g.types[s]), "type / particle-flag map differs from the reference after step %d" % s
        F, M, V, W = lb.forces()
        rF, rM, rV, rW = g.forces[s - 1]
        arm = float(parts["r"].max()); fmax = float(np.abs(rF).max())
        for a, b, floor in ((F, rF, 0.0), (M, rM, fmax * arm), (V, rV, 0.0)):
            worst_f = max(worst_f, np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-300))
    assert worst_x <= TOL_COUPLED and worst_f <= TOL_COUPLED, (worst_x, worst_f)
    # the fields at the end, against the reference's hashes where the path is order-deterministic (no free surface here)
    mine = lb.fetch()
    h = gu.state_hashes(mine)
    ref = g.hashes(g.steps)
    print("%s: trajectories %.2e, forces %.2e, bit-identical fields: %s" % (name, worst_x, worst_f, [k for k in gu.FIELDS if h[k] == ref[k]]))
    lb.close()


@pytest.mark.parametrize("name", ["bed_dem", "spheres_pbc_dem"])
def test_run_dem_equals_single_cycles_and_restart(name):
    g = gu.Golden(name)
    a, b, c = _gpu(g), _gpu(g), _gpu(g)
    a.runDem(60)
    for _ in range(60):
        b.runDem(1)
    c.runDem(25)
    blob = c.save_state()
    c.close()
    c = _gpu(g).load_state(blob)
    c.runDem(35)
    sa, sb, sc = a.demState(), b.demState(), c.demState()
    for k in ("x0", "x1", "w0"):
        assert np.array_equal(sa[k], sb[k]) and np.array_equal(sa[k], sc[k]), k
    fa, fb, fc = a.fetch(), b.fetch(), c.fetch()
    for k in fa:
        assert np.array_equal(fa[k], fb[k]) and np.array_equal(fa[k], fc[k]), k
    for lb in (a, b, c):
        lb.close()


def test_graph_replay_of_the_dem_coupled_cycle(monkeypatch):
    """lbGpuRunDem(count) replays two captured cycles (DEM sub-steps, coupling step, LB step): same as one by one."""
    g = gu.Golden("bed_dem")
    a = _gpu(g)
    monkeypatch.setenv("LBGPU_GRAPH", "0")
    b = _gpu(g)
    monkeypatch.delenv("LBGPU_GRAPH")
    for n in (30, 41):
        a.runDem(n); b.runDem(n)
    assert a.graph_info()[1] >= 25 and b.graph_info() == (0, 0)
    sa, sb = a.demState(), b.demState()
    for k in ("x0", "x1", "w0"):
        assert np.array_equal(sa[k], sb[k]), k
    assert sa["rebuilds"] == sb["rebuilds"] >= 2
    fa, fb = a.fetch(), b.fetch()
    for k in fa:
        assert np.array_equal(fa[k], fb[k]), k
    a.close(); b.close()


def test_cfg3_at_full_size_with_the_dem_on_the_device():
    """BASELINE's configuration 3 as the reference runs it -- the sphere's motion computed by the DEM every cycle -- for 1000
    cycles on the device (lbGpuRunDem), against the full-size fixture of the unmodified reference: trajectory, forces, the
    cell-type / particle-flag map at the fixture's 102 check points, the fields at steps 1 / 100 / 1000."""
    import cases
    import common
    import golden_full_util as gfu
    from hybird_b200 import LB, dem_init
    g = gfu.GoldenFull("cfg3_full")
    case = cases.catalogue()[g.case_name]
    lb = LB(dict(g.params)).latticeBolzmannInit(*g.init_arrays()).demInit(dem_init.dem_from_case(case))
    worst_x = worst_f = 0.0
    for s in range(1, g.steps + 1):
        lb.runDem(1)
        if s in g.type_sha or s % 50 == 0:
            parts, elmts, comps, flag = g.trace[s - 1]
            st = lb.demState()
            worst_x = max(worst_x, np.abs(st["x0"] - parts["x0"]).max(), np.abs(st["x1"] - elmts["x1"]).max())
            F, M, V, W = lb.forces()
            rF, rM, rV, rW = g.forces[s - 1]
            worst_f = max(worst_f, np.abs(F - rF).max() / np.abs(rF).max(), np.abs(V - rV).max() / np.abs(rV).max())
        if s in g.type_sha:
            t = lb.fetch(("type_flags",))["type_flags"] & 0x1F
            assert gfu.sha(t) == g.type_sha[s], "type / particle-flag map differs from the reference after step %d" % s
        if s in g.dumps:
            st = common.gpu_state(lb)
            idx = np.arange(g.stride // 2, g.N, g.stride, dtype=np.int64)
            act = np.isin(g.z["samp_type_%d" % s] & 0x0F, (0, 3))
            for k in ("n", "u", "visc"):
                a, b = st[k][idx][act], g.z["samp_%s_%d" % (k, s)][act]
                # against the field's scale: on the symmetry planes of this case a velocity component is rounding noise of
                # either sign, which a per-value relative error would compare with itself
                e = float(np.abs(a - b).max() / np.abs(b).max())
                assert e <= (1e-12 if s == 1 else TOL_COUPLED), (s, k, e)
    print("cfg3 at full size, DEM on the device, 1000 cycles: trajectory %.2e, forces %.2e" % (worst_x, worst_f))
    assert worst_x <= TOL_COUPLED and worst_f <= TOL_COUPLED, (worst_x, worst_f)
    lb.close()


@pytest.mark.parametrize("name", ["spheres_pbc_dem", "cfg5_mini_dem"])
def test_periodic_dem_boundaries_on_the_device(name):
    """Ghost particles on the device (pbcShift, createGhosts incl. the corner ghost, contacts across a periodic face, a rescan of
    the particle flags after every rebuild): the particle list handed to the LB side -- count, elements, positions of
    particles and ghosts -- the cell-type / particle-flag map and the forces against the reference, every cycle."""
    g = gu.Golden(name)  # cfg5_mini_dem: configuration 5 in small -- free surface, 14 spheres, 13 ghosts across the periodic y faces
    lb = _gpu(g)
    worst_x = worst_f = 0.0
    counts = set()
    for s in range(1, g.steps + 1):
        parts, elmts, comps, flag = g.trace[s - 1]
        lb.runDem(1)
        worst_x = max(worst_x, _worst(lb, parts, elmts))
        counts.add(len(parts))
        t = lb.fetch(("type_flags", "solidIndex"))
        assert np.array_equal(t["type_flags"] & 0x1F, g.types[s]), "type / particle-flag map differs from the reference after step %d" % s
        F, M, V, W = lb.forces()
        rF, rM, rV, rW = g.forces[s - 1]
        arm = float(parts["r"].max()); fmax = float(np.abs(rF).max())
        for a, b, floor in ((F, rF, 0.0), (M, rM, fmax * arm), (V, rV, 0.0)):
            worst_f = max(worst_f, np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-300))
    if name == "spheres_pbc_dem":
        assert lb.demState()["rebuilds"] >= 4 and len(counts) >= 2
    assert worst_x <= TOL_COUPLED and worst_f <= TOL_COUPLED, (worst_x, worst_f)
    print("%s: trajectories (particles and ghosts) %.2e, forces %.2e" % (name, worst_x, worst_f))
    lb.close()


def _coupled(g, n_slabs, steps):
    from hybird_b200 import LB
    prm = dict(g.params)
    prm["nSlabs"] = n_slabs; prm["nLocalSlabs"] = n_slabs
    lb = LB(prm).latticeBolzmannInit(*g.init_arrays()).demInit(g.dem())
    lb.runDem(steps)
    out = (lb.demState(), lb.fetch(("type_flags", "f")), lb.forces())
    lb.close()
    return out


def test_coupled_cycle_on_slabs_of_one_device():
    """The element forces are summed per slab first, so trajectories agree to rounding rather than bit for bit."""
    g = gu.Golden("bed_dem")
    (s1, f1, F1), (s3, f3, F3) = _coupled(g, 1, 80), _coupled(g, 3, 80)
    assert np.array_equal(f1["type_flags"], f3["type_flags"])
    for k in ("x0", "x1", "w0"):
        assert np.abs(s1[k] - s3[k]).max() <= TOL_COUPLED, k
    assert np.abs(F1[0] - F3[0]).max() <= TOL_COUPLED * np.abs(F1[0]).max()


@pytest.mark.parametrize("name", ["bed_dem", "cfg5_mini_dem"])
def test_coupled_cycle_on_two_gpus(name, tmp_path):
    """Two ranks, each advancing the same elements from the all-reduced forces (cfg5_mini_dem: with a free surface and the ghost
    particles of the periodic y faces, every rank reading its own particle count back)."""
    import os, subprocess, sys
    import torch
    import common
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = tmp_path / "ranks.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29900 + os.getpid() % 90), os.path.join(common.ROOT, "tests", "run_slab_ranks.py"), name, str(out), "--dem"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    z = np.load(out)
    g = gu.Golden(name)
    s1, f1, F1 = _coupled(g, 1, int(z["steps"]))
    assert np.array_equal(z["type_flags"], f1["type_flags"])
    for k in ("x0", "x1", "w0"):
        assert np.abs(z["dem_" + k] - s1[k]).max() <= TOL_COUPLED, k
    # ... and both follow the reference's recorded trajectory (ghost particles included)
    parts, elmts, comps, flag = g.trace[int(z["steps"]) - 1]
    assert len(z["dem_px0"]) == len(parts) and np.array_equal(z["dem_pcluster"], parts["clusterIndex"])
    assert np.abs(z["dem_px0"] - parts["x0"]).max() <= TOL_COUPLED


def _bed(n, seed=4321):
    """A bed of n spheres (radii 3-4, random velocities) between six walls, cfg5's recipe scaled to n."""
    from hybird_b200 import workloads
    g = gu.Golden("spheres_dem")
    dem = g.dem()
    side = (n / 20000.0) ** (1.0 / 3.0)
    hi = (160.0 * side, 255.0 * side, 1023.0 * side)
    bed = workloads._sphere_bed(n, (1.0, 1.0, 1.0), hi, 3.0, 4.0, seed)
    rng = np.random.default_rng(seed)
    rho = 2.5
    dem["elmts"] = [dict(size=1, radius=e["radius"], m=4.0 / 3.0 * rho * np.pi * e["radius"] ** 3, I=[0.4 * 4.0 / 3.0 * rho * np.pi * e["radius"] ** 5] * 3,
                         x0=e["x0"], x1=list(rng.uniform(-0.4, 0.4, 3)), w0=list(rng.uniform(-0.02, 0.02, 3))) for e in bed]
    dem["walls"] = [dict(n=[1 if a == k and s == 0 else (-1 if a == k else 0) for a in range(3)],
                         p=[(0.5 if s == 0 else hi[k] + 0.5) if a == k else 0.0 for a in range(3)],
                         vel=[0.0] * 3, omega=[0.0] * 3, rotCenter=[0.0] * 3, moving=0) for k in range(3) for s in (0, 1)]
    p = dem["params"]
    p["multiStep"] = 1; p["deltat"] = 1.0; p["nebrRange"] = 12.0; p["maxDisp"] = 6.0; p["demF"] = [-1e-3, 0.0, 3e-4]
    return g, dem


def test_grid_rebuild_equals_all_pairs(monkeypatch):
    """The neighbour table through the uniform grid (count - scan - fill, 27 cells, sorted lists) == the all-pairs pass: the
    same partner lists, hence bit-identical trajectories through contacts and a dozen rebuilds."""
    from hybird_b200 import LB
    g, dem = _bed(6000)
    out = []
    for grid in ("0", "1"):
        monkeypatch.setenv("LBGPU_DEM_GRID", grid)
        lb = LB(dict(g.params)).latticeBolzmannInit(*g.init_arrays()).demInit(dem)
        hydro = np.zeros((len(dem["elmts"]), 7))
        lb.demStep(hydro)
        for _ in range(120):
            lb.demStep()
        out.append((lb.demState(), lb.demContacts()))
        lb.close()
    monkeypatch.delenv("LBGPU_DEM_GRID")
    (a, ca), (b, cb) = out
    assert a["rebuilds"] == b["rebuilds"] >= 5 and a["longest_list"] == b["longest_list"] == 0  # (nonzero only on overflow)
    for k in ("x0", "x1", "w0"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(ca["FParticle"], cb["FParticle"]) and np.abs(ca["FParticle"]).max() > 0  # contacts happened


def test_partner_list_overflow_is_an_error():
    from hybird_b200 import LB
    from hybird_b200.abi import LbGpuError
    g = gu.Golden("spheres_dem")
    dem = g.dem()
    e0 = dem["elmts"][0]
    dem["elmts"] = [dict(e0, x0=[10.0 + 0.01 * k, 12.0, 15.0]) for k in range(60)]  # 59 partners each, the lists hold 48
    lb = LB(dict(g.params)).latticeBolzmannInit(*g.init_arrays()).demInit(dem)
    with pytest.raises(LbGpuError):
        lb.demStep()
        lb.synchronize()
    lb.close()
