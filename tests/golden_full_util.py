"""Loader + replay check for the hash-only fixtures of tests/golden/make_golden_full.py
(`<cfg>_full.npz`: a BASELINE.json configuration at its full size; `<mini>_long.npz`: 1000 steps of a free-surface
mini).  The same check drives the C restatement (CPU) and the CUDA engine (GPU)."""
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
TOL_STEP1, TOL_END, TOL_FORCE = 1e-12, 1e-9, 1e-9  # north_star


def names(kind):
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*_%s.npz" % kind)))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sha_active(arr, active, chunk=1 << 21):
    h = hashlib.sha256()
    for b in range(0, active.shape[0], chunk):
        m = active[b:b + chunk]
        if m.any():
            h.update(np.ascontiguousarray(arr[b:b + chunk][m] + 0.0).tobytes())
    return h.hexdigest()


class GoldenFull:
    def __init__(self, name):
        import lbo
        self.name = name
        self.z = z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        meta = json.loads(str(z["meta"]))
        self.meta = meta
        self.params = meta["params"]
        self.case_name = meta["case"]
        self.steps = int(meta["steps"])
        self.dumps = [int(s) for s in meta["dumps"]]
        self.stride = int(meta["stride"])
        self.N = int(np.prod(self.params["size"]))
        self.type_sha = {int(s): str(h) for s, h in zip(z["type_steps"], z["type_sha"])}
        self.type_counts = {int(s): c for s, c in zip(z["type_steps"], z["type_counts"])}
        self.trace = self.forces = None
        if self.params["nParts"]:
            with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as fh:
                fh.write(z["trace"].tobytes())
            try:
                self.trace = lbo.read_particle_trace(fh.name)
            finally:
                os.unlink(fh.name)
        if self.params["nElmts"] or self.params["nWalls"]:
            with tempfile.NamedTemporaryFile(suffix=".f", delete=False) as fh:
                fh.write(z["forces"].tobytes())
            try:
                self.forces = lbo.read_forces(fh.name, self.params["nElmts"], self.params["nWalls"])
            finally:
                os.unlink(fh.name)

    def init_arrays(self):
        """Initial state: stored (long fixtures) or rebuilt by the host restatement of the reference's initialisation and
        proven identical to the reference's through the stored hashes (full-size fixtures)."""
        z = self.z
        keys = ("type_flags", "solidIndex", "n", "u", "mass", "visc")
        if "init_type_flags" in z.files:
            return tuple(z["init_" + k] for k in keys)
        import cases
        from hybird_b200 import lattice_init as li
        case = cases.materialise(dict(cases.catalogue()[self.case_name]))
        parts = None
        if case.get("elements"):
            parts = li.expand_elements(case["elements"], self.params["unitLength"])[0]
        st = li.build_state(case, parts if parts is not None and len(parts) else None)
        arrs = (st.type_flags, st.solidIndex, st.n, st.u, st.mass, st.visc)
        for k, a in zip(keys, arrs):
            a = np.ascontiguousarray(a)
            if k == "type_flags":
                a = a & 0x3F
            assert sha(a.reshape(-1) if k != "u" else a.reshape(-1, 3)) == str(z["init_sha_" + k]), \
                "the rebuilt initial %s differs from the reference's" % k
        return arrs

    def configure(self, engine):
        if "curve_cells" in self.z.files:
            (engine.setCurves if hasattr(engine, "setCurves") else engine.set_curves)(self.z["curve_cells"], self.z["curve_delta"])
        if self.params.get("enforceMass"):
            (engine.setMassTarget if hasattr(engine, "setMassTarget") else engine.set_mass_target)(self.params["totalMass"])
        return engine


def check_run(g: GoldenFull, engine, get_types, get_state, exact=None, steps=None, log=print):
    """Replay the reference's inputs into `engine` and check every stored quantity.
    get_types(engine) -> (N,) uint8 type | p<<4 (| node<<5); get_state(engine) -> dict of FIELDS ('f' = post-collision)
    + type_flags.  exact: require the sha256 of every field (default: lattices without a free surface).
    Returns a report dict (worst errors per dump step)."""
    prm = g.params
    fs = bool(prm["freeSurface"])
    exact = (not fs) if exact is None else exact
    steps = steps or g.steps
    dem = g.trace is not None
    report = {}
    t0 = get_types(engine) & 0x1F
    assert sha(t0) == g.type_sha[0], "initial type map differs from the reference"
    for s in range(1, steps + 1):
        if fs:
            engine.latticeBoltzmannFreeSurfaceStep()
        elmts = parts = None
        if dem:
            parts, elmts, comps, flag = g.trace[s - 1]
            engine.latticeBoltzmannCouplingStep(flag, elmts, parts, comps)
        want_forces = g.forces is not None
        out = engine.latticeBolzmannStep(elmts, parts) if want_forces or not hasattr(engine, "lib") else \
            engine.latticeBolzmannStep(elmts, parts, fetch_forces=False)
        if want_forces:
            F, M, V, W = out
            rF, rM, rV, rW = g.forces[s - 1]
            arm = float(parts["r"].max()) if parts is not None and len(parts) else 0.0
            fmax = float(np.abs(rF).max()) if rF.size else 0.0
            for a, b, nm, floor in ((F, rF, "FHydro", 0.0), (M, rM, "MHydro", fmax * arm), (V, rV, "fluidVolume", 0.0),
                                    (W, rW, "wallFHydro", 0.0)):
                if b.size:
                    err = np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-300)
                    report["force_err"] = max(report.get("force_err", 0.0), float(err))
                    assert err <= TOL_FORCE, "%s step %d: rel err %.3g" % (nm, s, err)
        if s in g.type_sha:
            t = get_types(engine) & 0x1F
            if sha(t) != g.type_sha[s]:
                cnt = [int(np.count_nonzero((t & 15) == k)) for k in (0, 3, 2)] + [int(np.count_nonzero(t & 16))]
                raise AssertionError("type map differs from the reference after step %d: counts %s, reference %s"
                                     % (s, cnt, list(g.type_counts[s])))
        if s in g.dumps:
            st = get_state(engine)
            tf = st["type_flags"]
            active = np.isin(tf & 0x0F, (0, 3))
            bad = []
            for k in FIELDS:
                src = st["f"] if k == "fs" and "fs" not in st else st[k]
                if sha_active(src, active) != str(g.z["sha_%s_%d" % (k, s)]):
                    bad.append(k)
            # the regular subsample, within north_star's tolerance
            idx = np.arange(g.stride // 2, g.N, g.stride, dtype=np.int64)
            tol = TOL_STEP1 if s == 1 else TOL_END
            worst = {}
            assert np.array_equal(tf[idx] & 0x1F, g.z["samp_type_%d" % s] & 0x1F)
            act = np.isin(g.z["samp_type_%d" % s] & 0x0F, (0, 3))
            for k in ("n", "u", "mass", "visc"):
                worst[k] = common.max_rel(st[k][idx][act], g.z["samp_%s_%d" % (k, s)][act])
            f = st["f"] if "fs" not in st else st["fs"]
            worst["fs"] = common.max_rel(f[idx[::8]][act[::8]], g.z["samp_fs_%d" % s][act[::8]])
            report[s] = dict(not_bit_identical=bad, **worst)
            log("%s step %d: sample rel err %s; fields not bit-identical: %s" % (g.name, s, {k: "%.2g" % v for k, v in worst.items()}, bad or "none"))
            for k, v in worst.items():
                assert v <= tol, "step %d %s: rel err %.3g > %g" % (s, k, v, tol)
            if exact:
                assert not bad, "step %d: not bit-identical to the reference: %s" % (s, bad)
            if s == g.steps and "final_n" in g.z.files:  # small lattices: the complete arrays
                for k in ("n", "u", "mass", "visc"):
                    e = common.max_rel(st[k][active], g.z["final_" + k][active])
                    assert e <= tol, "final %s: rel err %.3g" % (k, e)
                e = common.max_rel(f[active], g.z["final_fs"][active])
                assert e <= tol, "final fs: rel err %.3g" % e
    return report
