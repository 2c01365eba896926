"""GPU: the slab code path.  (1) G slabs resident on ONE device (halo planes moved by device copies) must give
the state of the undivided lattice; (2) two processes on two GPUs (halo planes over NCCL send/recv, sums over
ncclAllReduce) must give it too.  Without a free surface the results are bit-identical; with one the global
mass surplus is summed per slab first, so the north_star tolerance (1e-9) applies to the fields while the
cell-type maps stay identical."""
import os
import subprocess
import sys

import numpy as np
import pytest

import common
import golden_util as gu

pytestmark = pytest.mark.gpu


def _run(g, n_slabs, steps):
    from hybird_b200 import LB
    prm = dict(g.params)
    prm["nSlabs"] = n_slabs
    prm["nLocalSlabs"] = n_slabs
    lb = LB(prm)
    lb.latticeBolzmannInit(*g.init_arrays())
    g.configure(lb)
    forces = []
    for s, F, M, V, W in gu.replay(g, lb, None):
        forces.append((F.copy(), M.copy(), V.copy(), W.copy()))
        if s == steps:
            break
    st = lb.fetch()
    lb.close()
    return st, forces


@pytest.mark.parametrize("name", ["cfg2_mini", "cfg3_mini", "cfg4_mini", "cfg5_mini", "periodic_all", "bubble_periodic",
                                  "couette_dyn", "slip_dyn", "two_spheres_kin", "cluster_dem", "drum_mini"])
@pytest.mark.parametrize("n_slabs", [2, 3])
def test_slabs_on_one_device_match_the_undivided_lattice(name, n_slabs):
    g = gu.Golden(name)
    steps = min(g.steps, 30)
    ref, fref = _run(g, 1, steps)
    got, fgot = _run(g, n_slabs, steps)
    assert np.array_equal(got["type_flags"], ref["type_flags"])
    act = np.isin(ref["type_flags"] & 15, (0, 3))
    p = (ref["type_flags"] & 0x10).astype(bool)
    assert np.array_equal(got["solidIndex"][p], ref["solidIndex"][p])
    exact = not g.params["freeSurface"]
    for k in ("f", "n", "u", "mass", "visc", "hydroForce"):
        a, b = got[k][act], ref[k][act]
        if exact:
            assert np.array_equal(a, b), k
        else:
            assert common.max_rel(a, b) <= 1e-9, (k, common.max_rel(a, b))
    for (F, M, V, W), (F0, M0, V0, W0) in zip(fgot, fref):
        # the torque on a sphere at rest is pure cancellation: its error scales with |F| x lever arm
        fmax = float(np.abs(F0).max()) if F0.size else 0.0
        for a, b, floor in ((F, F0, 0.0), (M, M0, 16.0 * fmax), (V, V0, 0.0), (W, W0, 0.0)):
            if a.size:
                scale = max(np.abs(b).max(), floor, 1e-300)
                assert np.abs(a - b).max() / scale <= 1e-9


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", ["peer", "nccl", "peer+device-init"])
@pytest.mark.parametrize("name", ["cfg2_mini", "cfg5_mini", "periodic_all", "cfg4_mini", "cfg3_mini"])
def test_two_processes_two_gpus_match_one_gpu(name, mode, tmp_path):
    """mode: how the per-step halo travels (peer memory over NVLink, or NCCL send/recv) and where the initial state is
    built (host arrays per slab, or lbGpuInitBox on every rank)."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    if mode != "peer" and name in ("periodic_all", "cfg3_mini"):
        pytest.skip("covered by the other cases")
    out = tmp_path / "ranks.npz"
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(common.ROOT, "tests", "run_slab_ranks.py"), name, str(out)] + \
          (["--nccl-halo"] if mode == "nccl" else []) + (["--device-init"] if "device-init" in mode else [])
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    z = np.load(out)
    g = gu.Golden(name)
    steps = int(z["steps"])
    ref, fref = _run(g, 1, steps)
    assert np.array_equal(z["type_flags"], ref["type_flags"])
    act = np.isin(ref["type_flags"] & 15, (0, 3))
    exact = not g.params["freeSurface"]
    for k in ("f", "n", "u", "mass"):
        a, b = z[k][act], ref[k][act]
        if exact:
            assert np.array_equal(a, b), k
        else:
            assert common.max_rel(a, b) <= 1e-9, (k, common.max_rel(a, b))
    if g.params["nElmts"]:
        F0 = np.stack([f[0] for f in fref]); M0 = np.stack([f[1] for f in fref])
        assert np.abs(z["F"] - F0).max() <= 1e-9 * max(np.abs(F0).max(), 1e-300)
        assert np.abs(z["M"] - M0).max() <= 1e-9 * max(np.abs(M0).max(), 16.0 * np.abs(F0).max(), 1e-300)


@pytest.mark.parametrize("name", ["cfg5_mini", "cfg4_mini", "cfg2_mini", "cfg3_mini"])
def test_four_processes_four_gpus_match_one_gpu(name, tmp_path):
    """The same with four ranks (interior ranks have a neighbour on either side: two face launches, two peers)."""
    if _n_gpus() < 4:
        pytest.skip("needs 4 GPUs (gpurun --gpus 4)")
    out = tmp_path / "ranks.npz"
    port = 29400 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "4", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(common.ROOT, "tests", "run_slab_ranks.py"), name, str(out)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    z = np.load(out)
    g = gu.Golden(name)
    steps = int(z["steps"])
    ref, fref = _run(g, 1, steps)
    assert np.array_equal(z["type_flags"], ref["type_flags"])
    act = np.isin(ref["type_flags"] & 15, (0, 3))
    exact = not g.params["freeSurface"]
    for k in ("f", "n", "u", "mass"):
        a, b = z[k][act], ref[k][act]
        if exact:
            assert np.array_equal(a, b), k
        else:
            assert common.max_rel(a, b) <= 1e-9, (k, common.max_rel(a, b))
    if g.params["nElmts"]:
        F0 = np.stack([f[0] for f in fref]); M0 = np.stack([f[1] for f in fref])
        assert np.abs(z["F"] - F0).max() <= 1e-9 * max(np.abs(F0).max(), 1e-300)
        assert np.abs(z["M"] - M0).max() <= 1e-9 * max(np.abs(M0).max(), 16.0 * np.abs(F0).max(), 1e-300)
