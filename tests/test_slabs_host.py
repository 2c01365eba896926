"""Host side of the slab decomposition (CPU, gloo, world_size 2 and 3): every rank builds only its window of
the initial state, the windows agree with the whole-lattice state, the NCCL id reaches every rank."""
import os
import sys

import numpy as np
import pytest

import common  # noqa: F401  (sys.path)
from hybird_b200 import lattice_init as li
from hybird_b200 import slabs

sys.path.insert(0, os.path.join(common.ROOT, "oracle"))
import cases  # noqa: E402


def test_slab_ranges_tile_the_interior():
    for Z in (5, 16, 24, 256, 1018):
        for G in (1, 2, 3, 4, 8):
            if G > Z - 2:
                continue
            planes = []
            for k in range(G):
                b, e = slabs.slab_range(Z, G, k)
                assert e > b
                planes += list(range(b, e))
            assert planes == list(range(1, Z - 1))


@pytest.mark.parametrize("name", ["cfg2_mini", "cfg4_mini", "cfg5_mini", "periodic_all", "bubble_periodic", "dam_newtonian"])
@pytest.mark.parametrize("world", [2, 3])
def test_window_state_equals_global_state(name, world):
    case = cases.catalogue()[name]
    prm = li.params_from_case(case)
    tr = common.KinematicTrace(case, prm)
    parts = tr.initial_particles() if len(tr.parts) else None
    full = li.build_state(case, parts)
    X, Y, Z = prm["size"]
    XY = X * Y
    # the global maximum a real run obtains through the all-reduce
    act = np.isin(full.type_flags & 15, (0, 3))
    x, y, z = li.coords(prm["size"])
    gmax = np.array([x[act].max(), y[act].max(), z[act].max()], dtype=np.float64)
    for rank in range(world):
        lo, hi = slabs.window_of(Z, world, rank)
        st = li.build_state(case, parts, window=(lo, hi), reduce_max=lambda v: np.maximum(v, gmax))
        b, e = slabs.slab_range(Z, world, rank)
        # planes the rank owns, plus true (non-periodic) shell planes, must be identical; cut planes are ghosts
        # filled by the halo exchange, but the window init reproduces them too
        sl = slice(lo * XY, hi * XY)
        assert np.array_equal(st.type_flags, full.type_flags[sl]), (name, rank)
        assert np.array_equal(st.solidIndex, full.solidIndex[sl])
        for k in ("n", "mass", "visc"):
            assert np.array_equal(getattr(st, k), getattr(full, k)[sl]), k
        assert np.array_equal(st.u, full.u[sl])
        assert slabs.owned_active(st, (lo, hi), Z, world, rank) == int(np.count_nonzero(act[b * XY:e * XY]))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = slabs.broadcast_unique_id(dist, rank, lambda: bytes(range(128)))
        case = cases.catalogue()["cfg4_mini"]
        st, win = slabs.build_slab_state(case, rank, world, None, dist)
        Z = st.params["size"][2]
        act = slabs.owned_active(st, win, Z, world, rank)
        red = slabs.max_reducer(dist)(np.array([rank, 10.0 - rank, 3.0]))
        q.put((rank, uid == bytes(range(128)), win, act, red.tolist(), st.type_flags.tobytes(), st.n.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_gloo_ranks_build_their_windows(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    case = cases.catalogue()["cfg4_mini"]
    full = li.build_state(case)
    X, Y, Z = full.params["size"]
    total = 0
    for rank, uid_ok, win, act, red, tf, n in got:
        assert uid_ok
        assert red == [world - 1.0, 10.0, 3.0]
        assert win == slabs.window_of(Z, world, rank)
        sl = slice(win[0] * X * Y, win[1] * X * Y)
        assert tf == full.type_flags[sl].tobytes()
        assert n == full.n[sl].tobytes()  # hydrostatic reference height reduced across the ranks
        total += act
    assert total == int(np.count_nonzero(np.isin(full.type_flags & 15, (0, 3))))
