"""Shared drivers for the parity tests: run one case through an engine with the reference's call
sequence (hybird.cpp:47-59) and compare states."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from hybird_b200 import lattice_init as li  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class KinematicTrace:
    """Particle inputs per LB step for prescribed motion (the harness' --motion kin/none):
    x0 += x1*unitTime before every step; newNeighborList raised every `rescan_every` steps."""

    def __init__(self, case, params):
        self.parts, self.elmts, self.comps = li.expand_elements(case.get("elements", []))
        self.x0 = np.array([e["x0"] for e in case.get("elements", [])], dtype=np.float64).reshape(-1, 3)
        self.motion = case.get("motion", "none")
        self.rescan_every = int(case.get("rescan_every", 0))
        self.dt = params["unitTime"]
        self.step = 0

    def initial_particles(self):
        return self.parts.copy()

    def next(self):
        self.step += 1
        if self.motion == "kin" and len(self.parts):
            self.x0 = self.x0 + self.elmts["x1"] * self.dt
            self.parts["x0"] = self.x0[self.parts["clusterIndex"]] + self.parts["radiusVec"]
        flag = bool(self.rescan_every and self.step % self.rescan_every == 0)
        return self.parts.copy(), self.elmts.copy(), self.comps.copy(), flag


class RecordedTrace:
    """Particle inputs recorded from the reference run (golden fixture)."""

    def __init__(self, steps):
        self.steps = steps
        self.k = 0

    def next(self):
        s = self.steps[self.k]
        self.k += 1
        return s


def make_oracle(state, threads=1):
    import lbo
    return lbo.Oracle(state.params, state.type_flags, state.solidIndex, state.n, state.u, state.mass, state.visc,
                      threads=threads)


def make_gpu(state, device=-1):
    from hybird_b200 import LB
    lb = LB(state.params, device=device)
    lb.latticeBolzmannInit(state.type_flags, state.solidIndex, state.n, state.u, state.mass, state.visc)
    return lb


def cycle(engine, params, trace, dem_solve=True):
    """One goCycle worth of LB calls; returns (F, M, V, wallF)."""
    if params["freeSurface"]:
        engine.latticeBoltzmannFreeSurfaceStep()
    parts = elmts = None
    if dem_solve:
        parts, elmts, comps, flag = trace.next()
        engine.latticeBoltzmannCouplingStep(flag, elmts, parts, comps)
    return engine.latticeBolzmannStep(elmts, parts)


def gpu_state(lb):
    d = lb.fetch()
    return d


def oracle_state(o):
    return dict(type_flags=np.array(o.type_flags), solidIndex=np.array(o.solid_index), n=np.array(o.n), u=np.array(o.u),
                mass=np.array(o.mass), visc=np.array(o.visc), shearRate=np.array(o.shear_rate),
                hydroForce=np.array(o.hydro_force), f=np.array(o.fs))


def max_rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    d = np.abs(a - b)
    s = np.maximum(np.abs(a), np.abs(b))
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(s > 0, d / s, 0.0)
    return float(np.nanmax(r)) if not np.isnan(d).any() else float("inf")


def compare(mine, ref, fields=("f", "n", "u", "mass", "visc", "shearRate", "hydroForce")):
    """Mismatch report between two state dicts (type map exact; fields on active cells)."""
    t_m, t_r = mine["type_flags"] & 0x3F, ref["type_flags"] & 0x3F
    rep = dict(type_mismatch=int(np.count_nonzero(t_m != t_r)))
    active = np.isin(ref["type_flags"] & 0x0F, (0, 3))
    p = (ref["type_flags"] & 0x10).astype(bool)
    rep["solidIndex_mismatch"] = int(np.count_nonzero(mine["solidIndex"][p] != ref["solidIndex"][p]))
    for k in fields:
        if k not in mine or k not in ref:
            continue
        a, b = mine[k][active], ref[k][active]
        rep[k + "_neq"] = int(np.count_nonzero(a != b))
        rep[k + "_rel"] = max_rel(a, b)
    return rep
