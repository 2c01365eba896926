"""Loader for the golden fixtures of tests/golden/ (generated from the unmodified reference by
tests/golden/make_golden.py) and the replay loop shared by the oracle and the GPU parity tests."""
from __future__ import annotations

import glob
import hashlib
import json
import os
import tempfile

import numpy as np

import common

GOLDEN_DIR = common.GOLDEN_DIR
FIELDS = ("fs", "n", "u", "mass", "visc", "shearRate", "hydroForce")


def names():
    """The step-by-step fixtures (make_golden.py); the hash-only *_long / *_full ones have their own loader."""
    all_ = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in all_ if not n.endswith(("_long", "_full"))]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class Golden:
    def __init__(self, name):
        import lbo
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.name = name
        meta = json.loads(str(z["meta"]))
        self.params = meta["params"]
        self.steps = int(meta["steps"])
        self.check_steps = [int(s) for s in meta["check_steps"]]
        self.dem_solve = bool(meta.get("dem_solve", 1))  # goCycle runs the coupling step only with demSolve (hybird.cpp:53,61)
        self.N = int(np.prod(self.params["size"]))
        self.types = z["types"]
        with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as fh:
            fh.write(z["trace"].tobytes())
            path = fh.name
        try:
            self.trace = lbo.read_particle_trace(path)
        finally:
            os.unlink(path)
        fpath = path + ".f"
        z["forces"].tofile(fpath)
        try:
            self.forces = lbo.read_forces(fpath, self.params["nElmts"], self.params["nWalls"])
        finally:
            os.unlink(fpath)

    def dem(self):
        """What the reference's DEM held after its initialisation (cases with the DEM in the loop), as dem_port parses it."""
        import dem_port
        return dem_port.parse_dem_text(str(self.z["dem"])) if "dem" in self.z.files else None

    def init_arrays(self):
        z = self.z
        return (z["init_type_flags"], z["init_solidIndex"], z["init_n"], z["init_u"], z["init_mass"], z["init_visc"])

    def configure(self, engine):
        """Curved-wall data and the DRUM mass target, for either engine (oracle or GPU mirror)."""
        if "curve_cells" in self.z.files:
            (engine.setCurves if hasattr(engine, "setCurves") else engine.set_curves)(self.z["curve_cells"], self.z["curve_delta"])
        if self.params.get("enforceMass"):
            (engine.setMassTarget if hasattr(engine, "setMassTarget") else engine.set_mass_target)(self.params["totalMass"])
        return engine

    def hashes(self, step):
        return {k: str(self.z["sha_%s_%d" % (k, step)]) for k in FIELDS}


def state_hashes(state):
    """state: dict with type_flags and FIELDS ('f' of the engines = post-collision = reference fs)."""
    active = np.isin(state["type_flags"] & 0x0F, (0, 3))
    out = {}
    for k in FIELDS:
        src = state["f"] if k == "fs" and "fs" not in state else state[k]
        out[k] = sha(np.asarray(src)[active] + 0.0)
    return out


def replay(g: Golden, engine, get_state, dem_solve=None, on_step=None):
    """Drive `engine` (oracle or GPU LB mirror) with the reference's recorded inputs.
    Yields (step, F, M, V, wallF)."""
    prm = g.params
    if dem_solve is None:
        dem_solve = g.dem_solve
    for s in range(1, g.steps + 1):
        parts, elmts, comps, flag = g.trace[s - 1]
        if prm["freeSurface"]:
            engine.latticeBoltzmannFreeSurfaceStep()
        if dem_solve:
            engine.latticeBoltzmannCouplingStep(flag, elmts, parts, comps)
        yield (s,) + tuple(engine.latticeBolzmannStep(elmts, parts, components=comps))
