"""Slab decomposition of the lattice across GPUs (one process per GPU) -- host side.

The lattice is cut along z (the slowest index of the reference's cell order i = x + X*(y + Y*z),
LB.cpp:2317-2337, so halo planes are contiguous): rank k owns the interior planes
slab_range(Z, G, k) and holds one ghost plane on each cut side.
"""
from __future__ import annotations

import numpy as np

from . import lattice_init as li
from .lb import LB


def slab_range(Z: int, G: int, k: int):
    """Interior planes 1..Z-2 split evenly: slab k owns global planes [begin, end)."""
    inner = Z - 2
    return 1 + (k * inner) // G, 1 + ((k + 1) * inner) // G


def build_engine(case: dict, rank: int, world: int, device: int = -1, dist=None):
    """State of this rank's slab built directly (never the whole lattice), uploaded to `device`."""
    prm = li.params_from_case(case)
    Zg = prm["size"][2]
    elements = case.get("elements", [])
    parts, elmts, comps = li.expand_elements(elements, prm["unitLength"])
    if world == 1:
        st = li.build_state(case, parts if len(parts) else None)
        lb = LB(st.params, device=device)
        lb.latticeBolzmannInit(st.type_flags, st.solidIndex, st.n, st.u, st.mass, st.visc)
        active = int(np.count_nonzero(np.isin(st.type_flags & 0x0F, (0, 3))))
        info = dict(params=st.params, active_local=active, active_total=active, global_z=Zg, parallelism="1 GPU",
                    parts=parts, elmts=elmts, comps=comps, kernel="k_step (fused pull stream + collide)",
                    bytes_resident=2 * 19 * 8 * st.type_flags.size + 60 * st.type_flags.size)
        return lb, info
    raise NotImplementedError("multi-GPU slabs: see lbGpuInit nSlabs (work in progress)")
