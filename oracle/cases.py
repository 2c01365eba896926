"""Case catalogue for the oracles (TEST INFRASTRUCTURE, not product code).

Every case is a plain dict of hybird configuration keys (the `key = value` pairs GetPot reads in
the reference: hybird.cpp:132-223, LB.cpp:85-188, DEM.cpp:13-183) plus harness options (initial
gas regions, wall velocities, particle list, prescribed particle motion).  The same dict feeds

  * oracle/_ref/ref_harness   (the unmodified reference, via a generated .cfg + particle file),
  * oracle/lb_oracle.c        (the C restatement, via hybird_b200.lattice_init.build_state),
  * the CUDA engine           (via the C ABI in include/lbgpu.h),

so that all three see identical inputs.  The five BASELINE.json configs are `cfg1`..`cfg5`;
`*_mini` are the scaled-down versions the CPU oracles finish in seconds.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF_HARNESS = os.path.join(HERE, "_ref", "ref_harness")

from hybird_b200.workloads import _BASE, _HARNESS_KEYS, make_case, _sphere_bed, catalogue, materialise  # noqa: E402,F401


def write_case_files(case, workdir):
    """Write <workdir>/<name>.cfg, the particle file and an empty object file; return the cfg path."""
    os.makedirs(workdir, exist_ok=True)
    name = case["name"]
    pfile = os.path.join(workdir, name + "_particles.dat")
    ofile = os.path.join(workdir, name + "_objects.dat")
    with open(pfile, "w") as f:
        els = case.get("elements", [])
        f.write("%d\n" % len(els))
        for i, e in enumerate(els):
            # index size radius x y z vx vy vz wx wy wz q0 q1 q2 q3 q0' q1' q2' q3'  (DEM.cpp:84-122)
            vals = [i, e["size"], repr(float(e["radius"]))] + [repr(float(v)) for v in e["x0"]] + \
                   [repr(float(v)) for v in e["x1"]] + [repr(float(v)) for v in e["w"]] + [1, 0, 0, 0, 0, 0, 0, 0]
            f.write(" ".join(str(v) for v in vals) + "\n")
    with open(ofile, "w") as f:
        f.write("0\n")
    cfg = os.path.join(workdir, name + ".cfg")
    with open(cfg, "w") as f:
        for k, v in case.items():
            if k in _HARNESS_KEYS:
                continue
            f.write("%s = %s\n" % (k, repr(float(v)) if isinstance(v, float) else v))
        f.write("particleFile = %s\nobjectFile = %s\n" % (pfile, ofile))
    return cfg


def harness_args(case):
    a = []
    for key, flag in (("fluid_box", "--fluid-box"), ("gas_box", "--gas-box"), ("fluid_sphere", "--fluid-sphere"),
                      ("gas_sphere", "--gas-sphere")):
        if key in case:
            a += [flag] + [str(v) for v in case[key]]
    for w in case.get("wall_vel", []):
        a += ["--wall-vel"] + [str(v) for v in w]
    a += ["--motion", case.get("motion", "none")]
    if case.get("rescan_every"):
        a += ["--rescan-every", str(case["rescan_every"])]
    return a


def run_reference(case, workdir, steps, dumps=(), types_every=False, time_mode=False, warmup=0, threads=1,
                  dump_neighbors=False, timeout=3600):
    """Run the compiled unmodified reference on `case`. Returns (out_prefix, stdout)."""
    if not os.path.exists(REF_HARNESS):
        raise FileNotFoundError(REF_HARNESS + " (run `make -C oracle ref` where /root/reference exists)")
    cfg = write_case_files(case, workdir)
    out = os.path.join(workdir, case["name"])
    cmd = [REF_HARNESS, "-c", cfg, "--out", out, "--steps", str(steps)] + harness_args(case)
    if dumps:
        cmd += ["--dump", ",".join(str(d) for d in dumps)]
    if types_every:
        cmd += ["--types-every"]
    if dump_neighbors:
        cmd += ["--dump-neighbors"]
    if time_mode:
        cmd += ["--time", "--warmup", str(warmup)]
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, text=True)
    if res.returncode != 0:
        raise RuntimeError("ref_harness failed (%d):\n%s" % (res.returncode, res.stdout[-4000:]))
    return out, res.stdout


if __name__ == "__main__":
    import sys
    cat = catalogue()
    nm = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    out, so = run_reference(cat[nm], "/tmp/hb_cases", steps, dumps=(0, steps), types_every=True)
    print(so[-1500:])
    print(out)
