"""One rank of a multi-process slab run (launched by torch.distributed.run from tests/test_gpu_slabs.py):
a golden case cut into WORLD_SIZE z-slabs, one GPU per rank, replayed with the recorded particle inputs;
rank 0 gathers the fields and writes them to OUT.   python run_slab_ranks.py CASE OUT [--device-init] [--nccl-halo]
--device-init: every rank initialises its planes on the device (lbGpuInitBox) instead of uploading host arrays;
--nccl-halo: the step halo travels over NCCL send/recv instead of through peer memory (LBGPU_PEER_HALO=0);
--dem: the DEM runs on the device too (lbGpuDemInit + lbGpuRunDem, every rank advances the same elements) instead of the
recorded particle inputs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

import common  # noqa: F401
import golden_util as gu
from hybird_b200 import LB, slabs

name, out = sys.argv[1], sys.argv[2]
device_init = "--device-init" in sys.argv
if "--nccl-halo" in sys.argv:
    os.environ["LBGPU_PEER_HALO"] = "0"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
slabs.init_comm(rank, world, local, dist)
g = gu.Golden(name)
prm = dict(g.params)
X, Y, Z = prm["size"]
prm.update(nSlabs=world, slabIndex=rank, nLocalSlabs=1)
lo, hi = slabs.window_of(Z, world, rank)
sl = slice(lo * X * Y, hi * X * Y)
tf, si, n, u, mass, visc = g.init_arrays()
lb = LB(prm, device=local)
if device_init:
    import cases
    from hybird_b200 import lattice_init as li
    case = cases.catalogue()[name]
    parts0 = li.expand_elements(case.get("elements", []), prm["unitLength"])[0]
    lb.latticeBolzmannInitBox(case, parts0 if len(parts0) else None)
else:
    lb.latticeBolzmannInit(tf[sl], si[sl], n[sl], u[sl], mass[sl], visc[sl])
steps = min(g.steps, 30)
Fs, Ms = [], []
extra = {}
if "--dem" in sys.argv:
    steps = min(g.steps, 60)
    lb.demInit(g.dem())
    for s in range(steps):
        lb.runDem(1)
        F, M, V, W = lb.forces()
        Fs.append(F.copy()); Ms.append(M.copy())
    st = lb.demState()
    pt = lb.demParticles()
    extra = dict(dem_x0=st["x0"], dem_x1=st["x1"], dem_w0=st["w0"], dem_px0=pt["x0"], dem_pcluster=pt["clusterIndex"])
else:
    for s, F, M, V, W in gu.replay(g, lb, None):
        Fs.append(F.copy()); Ms.append(M.copy())
        if s == steps:
            break
fields = slabs.gather_fields(lb, rank, world, dist, ("type_flags", "n", "u", "mass", "f"))
cnt = lb.counts()
if rank == 0:
    np.savez(out, steps=steps, F=np.stack(Fs), M=np.stack(Ms), fluid=cnt["fluid"], **fields, **extra)
dist.barrier()  # nobody unmaps / frees arrays a neighbour may still be writing into
lb.close()
slabs.finalize_comm()
dist.destroy_process_group()
