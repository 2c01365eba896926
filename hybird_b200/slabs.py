"""Slab decomposition of the lattice across GPUs (one process per GPU) -- host side.

The lattice is cut along z (the slowest index of the reference's cell order i = x + X*(y + Y*z),
LB.cpp:2317-2337, so halo planes are contiguous): rank k owns the interior planes
slab_range(Z, G, k) and holds one ghost plane on each cut side.  The reference itself is a single
process; what is mirrored here is its `LB` call surface, per rank.

torch.distributed is the plumbing only: it broadcasts the NCCL unique id and reduces the few host
scalars of the initialisation.  The halo exchange and the force/mass reductions of the time loop are
issued by liblbgpu.so itself (ncclSend/ncclRecv/ncclAllReduce on the engine's streams).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from . import lattice_init as li
from .lb import LB


def slab_range(Z: int, G: int, k: int):
    """Interior planes 1..Z-2 split evenly: slab k owns global planes [begin, end) (== lbGpuSlabRange)."""
    inner = Z - 2
    return 1 + (k * inner) // G, 1 + ((k + 1) * inner) // G


def window_of(Z: int, G: int, k: int):
    """Global planes [lo, hi) of the host arrays of slab k: the owned planes plus one plane below and above."""
    b, e = slab_range(Z, G, k)
    return b - 1, e + 1


def broadcast_unique_id(dist, rank: int, make_id):
    """Rank 0 creates the 128-byte id (make_id), every rank returns it (torch.distributed broadcast)."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.from_numpy(np.frombuffer(make_id(), dtype=np.uint8).copy())
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = buf.to(dev)
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def init_comm(rank: int, world: int, device: int, dist):
    """lbGpuCommUniqueId on rank 0 -> broadcast -> lbGpuCommInit on every rank."""
    lib = abi.load_library()

    def make_id():
        raw = (C.c_uint8 * 128)()
        abi.check(lib.lbGpuCommUniqueId(raw))
        return bytes(raw)

    uid = broadcast_unique_id(dist, rank, make_id) if world > 1 else bytes(128)
    raw = (C.c_uint8 * 128).from_buffer_copy(uid)
    abi.check(lib.lbGpuCommInit(raw, rank, world, device))


def finalize_comm():
    abi.load_library().lbGpuCommFinalize()


def max_reducer(dist):
    """Element-wise maximum of a small float64 vector over all ranks (hydrostatic reference height)."""
    if dist is None:
        return None

    def reduce(v):
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.as_tensor(np.asarray(v, dtype=np.float64)).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()
    return reduce


def build_slab_state(case: dict, rank: int, world: int, parts=None, dist=None):
    """The window of the initial state rank `rank` uploads: built directly, never the whole lattice."""
    prm = li.params_from_case(case)
    Z = prm["size"][2]
    lo, hi = window_of(Z, world, rank)
    st = li.build_state(case, parts, window=(lo, hi), reduce_max=max_reducer(dist) if world > 1 else None)
    st.params["nSlabs"] = world
    st.params["slabIndex"] = rank
    st.params["nLocalSlabs"] = 1
    return st, (lo, hi)


def owned_active(st, window, Z: int, world: int, rank: int) -> int:
    """Active cells in the planes this rank owns."""
    X, Y = st.params["size"][0], st.params["size"][1]
    b, e = slab_range(Z, world, rank)
    t = st.type_flags.reshape(window[1] - window[0], Y, X)[b - window[0]:e - window[0]]
    return int(np.count_nonzero(np.isin(t & 0x0F, (0, 3))))


def build_engine(case: dict, rank: int, world: int, device: int = -1, dist=None, device_init: bool = False):
    """State of this rank's slab built directly (never the whole lattice), uploaded to `device`.
    For world > 1 the communicator must exist already (init_comm).
    device_init (one process): the lattice is initialised on the device (lbGpuInitBox), no host state at all."""
    import time
    prm = li.params_from_case(case)
    Zg = prm["size"][2]
    elements = case.get("elements", [])
    parts, elmts, comps = li.expand_elements(elements, prm["unitLength"])
    if device_init:
        # every rank initialises its own planes on the device: no host state at all
        prm["nWalls"] = len(li.make_walls(prm))
        if world > 1:
            prm.update(nSlabs=world, slabIndex=rank, nLocalSlabs=1)
        lb = LB(prm, device=device)
        t0 = time.perf_counter()
        lb.latticeBolzmannInitBox(case, parts if len(parts) else None)
        lb.synchronize()
        init_s = time.perf_counter() - t0
        c = lb.counts(local=True)
        active = c["fluid"] + c["interface"]
        c = lb.counts()
        active_total = c["fluid"] + c["interface"]
        N = lb.N
        par = "1 GPU" if world == 1 else "%d z-slabs, one rank per GPU, face planes stored into the neighbour's ghost planes over NVLink (5 populations per face)" % world
        info = dict(upload_s=init_s, upload_bytes=0, params=prm, active_local=active, active_total=active_total, global_z=Zg,
                    parallelism=par, parts=parts, elmts=elmts, comps=comps, kernel="k_step (fused pull stream + collide)",
                    bytes_resident=2 * 19 * 8 * N + 60 * N, init="device (lbGpuInitBox)",
                    halo=None if world == 1 else ("peer memory (cudaIpc over NVLink): one put kernel per step" if lb.peer_halo()
                                                  else "NCCL send/recv"))
        return lb, info
    if world == 1:
        st = li.build_state(case, parts if len(parts) else None)
        active = int(np.count_nonzero(np.isin(st.type_flags & 0x0F, (0, 3))))
        active_total = active
        par = "1 GPU"
    else:
        st, win = build_slab_state(case, rank, world, parts if len(parts) else None, dist)
        active = owned_active(st, win, Zg, world, rank)
        import torch
        t = torch.tensor([active], dtype=torch.int64)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = t.to(dev)
        dist.all_reduce(t)
        active_total = int(t.item())
        par = "%d z-slabs, one rank per GPU, NCCL send/recv halo (5 populations per face)" % world
    lb = LB(st.params, device=device)
    t0 = time.perf_counter()
    lb.latticeBolzmannInit(st.type_flags, st.solidIndex, st.n, st.u, st.mass, st.visc)
    lb.synchronize()
    upload_s = time.perf_counter() - t0
    upload_bytes = int(sum(a.nbytes for a in (st.type_flags, st.solidIndex, st.n, st.u, st.mass, st.visc)))
    info = dict(upload_s=upload_s, upload_bytes=upload_bytes, params=st.params, active_local=active, active_total=active_total, global_z=Zg, parallelism=par,
                parts=parts, elmts=elmts, comps=comps, kernel="k_step (fused pull stream + collide)",
                bytes_resident=2 * 19 * 8 * st.type_flags.size + 60 * st.type_flags.size, init="host arrays (lbGpuInit)",
                halo=None if world == 1 else ("peer memory (cudaIpc over NVLink): one put kernel per step" if lb.peer_halo()
                                              else "NCCL send/recv"))
    return lb, info


def gather_fields(lb: LB, rank: int, world: int, dist, fields=("type_flags", "n", "u", "mass", "f")):
    """Global arrays on rank 0 assembled from the owned planes of every rank (tests, output steps)."""
    X, Y, Z = lb.params["size"]
    lo, hi = window_of(Z, world, rank)
    b, e = slab_range(Z, world, rank)
    d = lb.fetch(fields)
    # owned planes, plus the true shell plane at either end of the lattice
    p0 = 0 if rank == 0 else b - lo
    p1 = (hi - lo) if rank == world - 1 else e - lo
    mine = {k: v.reshape((hi - lo, Y * X) + v.shape[1:])[p0:p1] for k, v in d.items()}
    out = [None] * world
    dist.all_gather_object(out, mine)
    if rank != 0:
        return None
    return {k: np.concatenate([o[k] for o in out]).reshape((X * Y * Z,) + d[k].shape[1:]) for k in fields}
