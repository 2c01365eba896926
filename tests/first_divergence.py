"""Ad-hoc: first step at which the GPU engine and the C restatement differ on a catalogue case.
python tests/first_divergence.py CASE STEPS"""
import sys
import numpy as np
import common
from common import li
sys.path.insert(0, common.ROOT + "/oracle")
import cases  # noqa: E402

name, steps = sys.argv[1], int(sys.argv[2])
case = cases.catalogue()[name]
prm = li.params_from_case(case)
tr_o, tr_g = common.KinematicTrace(case, prm), common.KinematicTrace(case, prm)
state = li.build_state(case, tr_o.initial_particles() if len(tr_o.parts) else None)
o, g = common.make_oracle(state), common.make_gpu(state)
X, Y, Z = prm["size"]
for s in range(1, steps + 1):
    common.cycle(o, state.params, tr_o); common.cycle(g, state.params, tr_g)
    a, b = common.gpu_state(g), common.oracle_state(o)
    rep = common.compare(a, b)
    bad = {k: v for k, v in rep.items() if v and (k.endswith("_rel") and v > 1e-12 or k.endswith("mismatch"))}
    if bad:
        print("step", s, bad)
        act = np.isin(b["type_flags"] & 15, (0, 3))
        for k in ("mass", "n", "f"):
            d = np.abs(a[k] - b[k]); d = d.reshape(len(act), -1).max(axis=1) * act
            idx = np.argsort(-d)[:6]
            for i in idx:
                if d[i] > 0:
                    print("  ", k, "cell", int(i), "xyz", int(i % X), int(i // X % Y), int(i // (X * Y)), "type gpu/ref", int(a["type_flags"][i]), int(b["type_flags"][i]),
                          "gpu", a[k][i] if a[k].ndim == 1 else "-", "ref", b[k][i] if b[k].ndim == 1 else "-", "diff", float(d[i]))
        break
else:
    print("no divergence in", steps, "steps")
if len(sys.argv) > 3:
    i = int(sys.argv[3])
    np.set_printoptions(precision=6, linewidth=200)
    print("cell", i, "f gpu", a["f"][i]); print("cell", i, "f ref", b["f"][i])
    for k in ("n", "mass", "visc", "shearRate"):
        print(k, a[k][i], b[k][i])
    print("u", a["u"][i], b["u"][i])
