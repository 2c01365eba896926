"""Hash-only golden fixtures at CONFIGURATION SIZE and at 1000 steps, from the UNMODIFIED reference.

    python tests/golden/make_golden_full.py [NAME ...]      (needs oracle/_ref/ref_harness)

north_star's success criterion is stated on the five configurations themselves: identical cell-type /
interface maps for the first 100 steps, fields within 1e-12 after one step and 1e-9 after 1000.  A full
state of a 256^3 lattice is 4-6 GB, so these fixtures keep

  type_sha[s]          sha256 of the (type | p<<4) byte map after init (s = 0) and after every step 1..100, and
                       after every later dump step
  type_counts[s]       fluid / interface / gas / flagged cell counts of the same maps (diagnosis)
  sha_<field>_<s>      sha256 over the active cells (index order, -0.0 canonicalised) of fs, n, u, mass, visc,
                       shearRate, hydroForce at the dump steps (1, 100, 1000): the bit-exact check of lattices
                       without a free surface
  samp_*_<s>           a regular subsample of the lattice (every `stride`-th cell: type byte, n, u, mass, visc; fs on
                       every 8th of those) at the dump steps: the tolerance check of free-surface lattices, whose
                       mass surplus is summed in another order on the device
  init_sha_*           sha256 of the initial state, so a test can prove it starts from the reference's own state
  trace / forces       particle inputs and element / wall forces of every step (coupled configurations)

`<name>_full.npz` = a BASELINE.json configuration at its full size (cfg1..cfg4; cfg5 needs ~35 GB and hours per
step on the CPU and stays property-checked).  `<name>_long.npz` = a free-surface mini run for 1000 steps, with the
complete final arrays (they are small).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cases  # noqa: E402
import lbo  # noqa: E402

FIELDS = ("fs", "n", "u", "mass", "visc", "shearRate", "hydroForce")

# name -> (case, steps, dump steps, reference threads).  Multi-threaded reference runs are deterministic for these
# cases: the only order-dependent OpenMP constructs (LB.cpp:1216 reduction, :1897 critical) sum moving-wall mass and
# element forces, and the coupled case replays its own recorded particle trace.
FULL = {
    "cfg1_full": ("cfg1", 1000, (1, 100, 1000), 4),
    "cfg3_full": ("cfg3", 1000, (1, 100, 1000), 4),
    "cfg4_full": ("cfg4", 1000, (1, 100, 1000), 4),
    "cfg2_full": ("cfg2", 1000, (1, 100, 1000), 4),
}
LONG = {nm + "_long": (nm, 1000, (1, 100, 1000), 1)
        for nm in ("cfg4_mini", "dam_newtonian", "droplet", "bubble_periodic", "cfg1_mini", "cfg5_mini", "drum_mini",
                   "drum_bingham")}
TYPES_UNTIL = 100


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sha_active(arr, active, chunk=1 << 21):
    """sha256 over arr[active] + 0.0 without materialising it (full-size fields are GBs)."""
    h = hashlib.sha256()
    n = active.shape[0]
    for b in range(0, n, chunk):
        m = active[b:b + chunk]
        if m.any():
            h.update(np.ascontiguousarray(np.asarray(arr[b:b + chunk])[m] + 0.0).tobytes())
    return h.hexdigest()


def sample_stride(N):
    """A prime-ish stride giving ~16-32 k sampled cells (all cells of a small lattice)."""
    if N <= 40000:
        return 1
    s = max(2, N // 24000) | 1
    while any(s % p == 0 for p in (3, 5, 7, 11, 13)):
        s += 2
    return s


def summarise(st, step, stride):
    tf = np.asarray(st["type_flags"])
    active = np.isin(tf & 0x0F, (0, 3))
    d = {}
    for k in FIELDS:
        d["sha_%s_%d" % (k, step)] = sha_active(st[k], active)
    idx = np.arange(stride // 2, tf.shape[0], stride, dtype=np.int64)
    d["samp_type_%d" % step] = tf[idx] & 0x3F
    for k in ("n", "u", "mass", "visc"):
        d["samp_%s_%d" % (k, step)] = np.asarray(st[k][idx])
    d["samp_fs_%d" % step] = np.asarray(st["fs"][idx[::8]])
    d["agg_%d" % step] = np.array([float(np.asarray(st["n"])[active].sum()), float(np.asarray(st["mass"])[active].sum()),
                                   float(np.abs(np.asarray(st["u"])[active]).max()) if active.any() else 0.0,
                                   float(active.sum())])
    return d


def generate(name, workdir="/tmp/hb_golden_full", keep=False):
    base, steps, dumps, threads = (FULL.get(name) or LONG[name])
    is_long = name in LONG
    case = cases.materialise(dict(cases.catalogue()[base]))
    case["name"] = name
    wd = os.path.join(workdir, name)
    os.makedirs(wd, exist_ok=True)
    t0 = time.time()
    cfg = cases.write_case_files(case, wd)
    out = os.path.join(wd, name)
    import subprocess
    cmd = [cases.REF_HARNESS, "-c", cfg, "--out", out, "--steps", str(steps)] + cases.harness_args(case) + \
          ["--dump", ",".join(str(d) for d in (0,) + tuple(dumps)), "--lite", "--types-until", str(TYPES_UNTIL)]
    res = subprocess.run(cmd, env=dict(os.environ, OMP_NUM_THREADS=str(threads)), stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("ref_harness failed:\n" + res.stdout[-3000:])
    t_ref = time.time() - t0
    hdr = lbo.read_log(out + "_log.txt")
    N = int(np.prod(hdr["size"]))
    stride = sample_stride(N)
    type_steps = list(range(0, min(TYPES_UNTIL, steps) + 1)) + [s for s in dumps if s > TYPES_UNTIL]
    types = np.memmap(out + "_types.bin", dtype=np.uint8, mode="r").reshape(-1, N)
    assert types.shape[0] == len(type_steps), (types.shape, len(type_steps))
    d = dict(type_steps=np.asarray(type_steps, dtype=np.int64),
             type_sha=np.asarray([sha(types[k]) for k in range(len(type_steps))]),
             type_counts=np.asarray([[int(np.count_nonzero((types[k] & 15) == t)) for t in (0, 3, 2)] +
                                     [int(np.count_nonzero(types[k] & 16))] for k in range(len(type_steps))], dtype=np.int64))
    st0 = lbo.read_state(out + "_state%06d.bin" % 0)
    for k in ("type_flags", "solidIndex", "n", "u", "mass", "visc"):
        d["init_sha_" + k] = sha(np.asarray(st0[k]))
    if is_long:  # small lattices carry their initial state like the short fixtures do
        for k in ("type_flags", "solidIndex", "n", "u", "mass", "visc"):
            d["init_" + k] = np.asarray(st0[k])
        if len(st0.get("curve_cells", ())):
            d["curve_cells"] = st0["curve_cells"]; d["curve_delta"] = st0["curve_delta"]
    elif len(st0.get("curve_cells", ())):
        raise RuntimeError("full-size curved lattices are not covered")
    del st0
    for s in dumps:
        st = lbo.read_state(out + "_state%06d.bin" % s)
        d.update(summarise(st, s, stride))
        if is_long and s == steps:
            for k in ("n", "u", "mass", "visc"):
                d["final_" + k] = np.asarray(st[k])
            d["final_fs"] = np.asarray(st["fs"])
        del st
    d["trace"] = np.fromfile(out + "_parts.bin", dtype=np.uint8)
    d["forces"] = np.fromfile(out + "_forces.bin", dtype="<f8")
    d["meta"] = json.dumps(dict(params=hdr, steps=steps, dumps=list(dumps), case=base, stride=stride, threads=threads,
                                reference_seconds=round(t_ref, 1)))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print("%-22s %s steps %4d  %9d cells  stride %5d  ref %.0f s  %7.1f KiB" % (name, base, steps, N, stride, t_ref,
                                                                             os.path.getsize(path) / 1024), flush=True)
    if not keep:
        shutil.rmtree(wd, ignore_errors=True)


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or list(LONG) + list(FULL)
    for nm in names:
        generate(nm, keep="--keep" in sys.argv)
