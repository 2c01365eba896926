"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

    python tests/golden/make_golden.py [CASE ...]

Needs oracle/_ref/ref_harness (`make -C oracle ref`, only possible where /root/reference is
mounted).  The reference ships no golden vectors, so they are produced by
running the reference itself at OMP_NUM_THREADS=1 on the small cases of oracle/cases.py.

One `<case>.npz` per case:
  meta           JSON: params as the reference parsed them (lbo.read_log), steps, nWalls, nElmts
  init_*         the state LB::latticeBolzmannInit produced (type_flags, solidIndex, n, u, mass, visc)
  trace          the particle inputs the reference's LB calls saw each step (raw PREFIX_parts.bin)
  forces         per step: elmts[].{FHydro,MHydro,fluidVolume} and walls[].FHydro (raw doubles)
  types          (steps+1, N) uint8: nodeType::t | p<<4 after init and after every step
  sha_<field>_<step>   sha256 of the field's bytes over the active cells after `step` in CHECK_STEPS
  final_n / final_mass / final_u  full arrays after the last step (for diagnosing a hash mismatch)
  dem            (cases with the DEM in the loop) text: material constants, time step, tables' range, elements, walls
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cases  # noqa: E402
import lbo  # noqa: E402

FIELDS = ("fs", "n", "u", "mass", "visc", "shearRate", "hydroForce")

# case -> number of LB steps.  Free-surface cases run 100 steps: north_star asks for identical
# cell-type / interface maps over the first 100 steps.
STEPS = {
    "cfg2_mini": 30, "channel_oblique": 30, "box_noforce": 20, "periodic_all": 30,
    "smago_channel": 30, "bingham_channel": 30, "bingham_smago": 30,
    "couette_dyn": 30, "slip_box": 30, "slip_dyn": 30,
    "sphere_fixed": 40, "cfg3_mini": 40, "sphere_kin": 40, "two_spheres_kin": 40, "cluster_dem": 40,
    "cfg4_mini": 100, "dam_newtonian": 100, "droplet": 100, "bubble_periodic": 100,
    "cfg1_mini": 100, "cfg5_mini": 100,
    "drum_mini": 300, "drum_bingham": 200,
    # DEM in the loop (contacts, table rebuilds): the fixtures also carry what DEM::discreteElementInit left (`dem`)
    "spheres_dem": 120, "spheres_hertz": 120, "bed_dem": 200, "spheres_pbc_dem": 160, "clusters_hit": 120, "cfg5_mini_dem": 100,
}


def check_steps(steps):
    return sorted({1, 2, steps // 2, steps})


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def state_hashes(state, step):
    """Hashes over the active cells of one state dict (keys as lbo.read_state / tests.common)."""
    active = np.isin(state["type_flags"] & 0x0F, (0, 3))
    out = {}
    for k in FIELDS:
        # +0.0 canonicalises -0.0 (the sign of a zero is not a result the reference defines)
        out["sha_%s_%d" % (k, step)] = sha(np.asarray(state[k])[active] + 0.0)
    return out


def generate(name, workdir="/tmp/hb_golden"):
    case = cases.catalogue()[name]
    steps = STEPS[name]
    cs = check_steps(steps)
    out, _ = cases.run_reference(case, workdir, steps, dumps=[0] + cs, types_every=True, threads=1)
    hdr = lbo.read_log(out + "_log.txt")
    st0 = lbo.read_state(out + "_state%06d.bin" % 0)
    N = int(np.prod(hdr["size"]))
    d = dict(
        meta=json.dumps(dict(params=hdr, steps=steps, check_steps=cs, case=name, dem_solve=int(case.get("demSolve", 1)))),
        init_type_flags=st0["type_flags"], init_solidIndex=st0["solidIndex"], init_n=st0["n"], init_u=st0["u"],
        init_mass=st0["mass"], init_visc=st0["visc"],
        trace=np.fromfile(out + "_parts.bin", dtype=np.uint8),
        forces=np.fromfile(out + "_forces.bin", dtype="<f8"),
        types=np.fromfile(out + "_types.bin", dtype=np.uint8).reshape(-1, N),
    )
    if case.get("motion") == "dem" and os.path.exists(out + "_dem.txt"):
        d["dem"] = open(out + "_dem.txt").read()  # parameters, elements and walls of the reference's DEM after its init
    if len(st0.get("curve_cells", ())):
        d["curve_cells"] = st0["curve_cells"]; d["curve_delta"] = st0["curve_delta"]
    d["sha_init_f"] = sha(st0["f"][np.isin(st0["type"], (0, 3))] + 0.0)
    for s in cs:
        st = lbo.read_state(out + "_state%06d.bin" % s)
        d.update(state_hashes(st, s))
        if s == steps:
            d["final_n"] = st["n"]; d["final_mass"] = st["mass"]; d["final_u"] = st["u"]
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print("%-18s steps %3d  %7d cells  %6.1f KiB" % (name, steps, N, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    names = sys.argv[1:] or list(STEPS)
    for nm in names:
        generate(nm)
