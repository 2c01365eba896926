"""Pin the C restatement against the unmodified reference (TEST INFRASTRUCTURE).

    python oracle/check_vs_ref.py CASE [STEPS]

Runs oracle/_ref/ref_harness on CASE, loads the reference's own initial state into the C
restatement, replays the particle inputs the reference saw, and compares every field bit for bit
after every dumped step.  Used by tests/test_oracle_vs_reference.py and to (re)generate
tests/golden/.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases  # noqa: E402
import lbo  # noqa: E402


def max_rel(a, b):
    d = np.abs(a - b)
    s = np.maximum(np.abs(a), np.abs(b))
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(s > 0, d / s, 0.0)
    return float(r.max()) if r.size else 0.0


def compare_state(o: lbo.Oracle, ref: dict, label=""):
    """Return dict of mismatch measures between the oracle state and one reference dump."""
    res = {}
    tf = o.type_flags
    res["type_mismatch"] = int(np.count_nonzero((tf & 0x3F) != (ref["type_flags"] & 0x3F)))
    pm = (ref["flags"] & 1).astype(bool)
    res["solidIndex_mismatch"] = int(np.count_nonzero(o.solid_index[pm] != ref["solidIndex"][pm]))
    for name, mine, theirs in (("f", o.f, ref["f"]), ("fs", o.fs, ref["fs"]), ("n", o.n, ref["n"]), ("u", o.u, ref["u"]),
                               ("hydroForce", o.hydro_force, ref["hydroForce"]), ("mass", o.mass, ref["mass"]),
                               ("visc", o.visc, ref["visc"]), ("shearRate", o.shear_rate, ref["shearRate"])):
        active = np.isin(ref["type"], (0, 3))
        a, b = np.asarray(mine)[active], np.asarray(theirs)[active]
        res[name + "_bits"] = int(np.count_nonzero(a.view(np.uint64) != b.view(np.uint64)))
        res[name + "_rel"] = max_rel(a, b)
    return res


def run_case(case, steps, workdir="/tmp/hb_cases", dumps=None, verbose=True, types_every=True):
    dumps = sorted(set(dumps if dumps is not None else (0, 1, steps)))
    out, _ = cases.run_reference(case, workdir, steps, dumps=dumps, types_every=types_every)
    hdr = lbo.read_log(out + "_log.txt")
    st0 = lbo.read_state(out + "_state%06d.bin" % 0)
    o = lbo.Oracle(hdr, st0["type_flags"], st0["solidIndex"], st0["n"], st0["u"], st0["mass"], st0["visc"], f=st0["f"])
    if len(st0.get("curve_cells", ())):
        o.set_curves(st0["curve_cells"], st0["curve_delta"])
    if hdr.get("enforceMass"):
        o.set_mass_target(hdr["totalMass"])
    trace = lbo.read_particle_trace(out + "_parts.bin")
    forces = lbo.read_forces(out + "_forces.bin", hdr["nElmts"], hdr["nWalls"])
    N = int(np.prod(hdr["size"]))
    types = np.fromfile(out + "_types.bin", dtype=np.uint8).reshape(-1, N) if types_every else None
    report = dict(case=case["name"], steps=steps, type_map_mismatch_steps=0, force_rel=0.0, wall_rel=0.0, states={})
    demSolve = int(case.get("demSolve", 1))
    for s in range(1, steps + 1):
        parts, elmts, comps, flag = trace[s - 1]
        if hdr["freeSurface"]:
            o.latticeBoltzmannFreeSurfaceStep()
        if demSolve:
            o.latticeBoltzmannCouplingStep(flag, elmts, parts, comps)
        F, M, V, Wf = o.latticeBolzmannStep(elmts, parts)
        if types is not None:
            mine = (o.type_flags & 0x1F)
            if np.count_nonzero(mine != types[s]):
                report["type_map_mismatch_steps"] += 1
        if forces:
            rF, rM, rV, rW = forces[s - 1]
            for a, b in ((F, rF), (M, rM), (V, rV)):
                report["force_rel"] = max(report["force_rel"], max_rel(a, b))
            report["wall_rel"] = max(report["wall_rel"], max_rel(Wf, rW))
        if s in dumps:
            ref = lbo.read_state(out + "_state%06d.bin" % s)
            report["states"][s] = compare_state(o, ref)
    if verbose:
        print(case["name"], "steps", steps, "type-map mismatch steps:", report["type_map_mismatch_steps"],
              "force_rel %.3g wall_rel %.3g" % (report["force_rel"], report["wall_rel"]))
        for s, r in report["states"].items():
            bad = {k: v for k, v in r.items() if v}
            print("  step", s, "OK (bit-identical)" if not bad else bad)
    o.close()
    return report


if __name__ == "__main__":
    cat = cases.catalogue()
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg2_mini"]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    for nm in names:
        run_case(cat[nm], steps)
